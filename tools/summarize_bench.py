#!/usr/bin/env python3
"""gpurun_out/r2_bench_<workload>.json (tools/final_runs.sh) -> profiles/r2_bench_configs.md"""
import json, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = ["# One `bench.py --workload W --steps 2 --warmup 3` line per BASELINE configuration on one B200 (`tools/final_runs.sh`)", "",
        "| workload | solve ms (device-timed) | V-cycles | ms per V-cycle | e2e `SSS_amg_solve` ms (first call) | reference CPU path ms (1 core; sample) | e2e speed-up | dominant kernel | its share | chain-floor fraction | launches per solve | SM MHz / reasons |",
        "|---|---|---|---|---|---|---|---|---|---|---|---|"]
per_level = []
for w in ("p2d256", "p3d128", "p3d256", "v27_192", "aniso256"):
    p = os.path.join(ROOT, "gpurun_out", f"r2_bench_{w}.json")
    if not os.path.exists(p):
        continue
    d = json.load(open(p))
    rf, cpu, e = d["roofline"], d["cpu_baseline"], d["e2e"]
    rows.append(f"| {w} | {d['value']:.1f} | {d['vcycles']} | {d['ms_per_vcycle']:.1f} | {e['value']:.1f} ({e['first_call_ms']:.0f}) | {cpu['value']:.0f} ({cpu['sample'].split(' on 1 of')[0]}) | "
                f"{cpu['value'] / e['value']:.2f} | `{rf['kernel']}` | {rf['share_of_step']:.2f} | {rf['chain_floor_frac'] if rf['chain_floor_frac'] is None else round(rf['chain_floor_frac'], 2)} | "
                f"{d['gpu_launches_per_step']:.0f} | {d['clocks']['sm_mhz']} / {d['clocks']['reasons']} |")
    per_level.append(f"\n## {w}: {d['config']['workload']}\n\n| level | kernel | rows | nnz | wavefronts | ms per solve | share | µs per wavefront | chain-floor fraction |\n|---|---|---|---|---|---|---|---|---|")
    for k in sorted(d["kernels"], key=lambda k: k["level"]):
        cf = k["chain_floor_frac"]
        per_level.append(f"| {k['level']} | `{k['kernel']}` | {k['rows']} | {k['nnz']} | {k['wavefronts']} | {k['ms_per_solve']:.1f} | {k['share']:.3f} | {k['us_per_wavefront']:.2f} | {'' if cf is None else round(cf, 2)} |")
    hk = rf["hbm_kernels"]
    per_level.append("\nHBM-bound level-0 kernels: " + ", ".join(f"{n} {v['gbs']:.0f} GB/s ({v['frac']:.2f})" for n, v in hk.items()))
    ph = d["phase_ms_per_solve"]
    per_level.append("Phases (ms per solve): " + ", ".join(f"{n} {v:.1f}" for n, v in ph.items()))
open(os.path.join(ROOT, "profiles", "r2_bench_configs.md"), "w").write("\n".join(rows + per_level) + "\n")
print("\n".join(rows))
