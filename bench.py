#!/usr/bin/env python3
"""bench.py -- V-cycle solve time to 1e-8 of the AMG solve phase on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload p3d128] [--impl amgb200|reference] [--fast]

A "step" is one complete solve of the workload (x0 = 1, b = 1, tol 1e-8, the reference's defaults
apart from the tolerance) on a hierarchy that is already resident in HBM.  Prints ONE JSON line.

  value / ms_per_step   device-timed (CUDA events on the library's stream) solve time, max over ranks
  e2e                   the same solve through the reference-facing C ABI call SSS_amg_solve(mg, x, b)
                        with HOST buffers: hierarchy analysis + H2D upload + solve + D2H inside the timer
  roofline              the kernel with the largest share of the step (live per-level CUDA-event timing
                        of one profiled solve) against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline          the reference's own CPU path (oracle/_ref, 1 thread -- it has no live OpenMP
                        region) timed on this box's host, same workload, one full solve
--impl reference runs that CPU path as the timed arm (the one place besides cpu_baseline where
bench.py executes anything under oracle/).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {           # name: (generator kind, N, eps_z, description, V-cycles to 1e-8 [reference fixtures / SURVEY.md Appendix C], reference sample cycles)
    "p2d256": ("p2d", 256, 0.0, "2D 5-point Poisson 256x256", 8, None),
    "p3d64": ("p3d", 64, 0.0, "3D 7-point Poisson 64^3", 7, None),
    "p3d128": ("p3d", 128, 0.0, "3D 7-point Poisson 128^3", 10, None),
    "p3d256": ("p3d", 256, 0.0, "3D 7-point Poisson 256^3", 20, 2),
    "aniso64": ("aniso3d", 64, 1e-3, "anisotropic 3D diffusion (1,1,1e-3) 64^3", 7, None),
    "aniso128": ("aniso3d", 128, 1e-3, "anisotropic 3D diffusion (1,1,1e-3) 128^3", None, None),
    "aniso256": ("aniso3d", 256, 1e-3, "anisotropic 3D diffusion (1,1,1e-3) 256^3", 8, 3),
    "v27_32": ("v27", 32, 0.0, "3D 27-point variable-coefficient diffusion 32^3", 9, None),
    "v27_64": ("v27", 64, 0.0, "3D 27-point variable-coefficient diffusion 64^3", 11, None),
    "v27_96": ("v27", 96, 0.0, "3D 27-point variable-coefficient diffusion 96^3", None, None),
    "v27_192": ("v27", 192, 0.0, "3D 27-point variable-coefficient diffusion 192^3", 23, 2),
}
METRIC = "vcycle_solve_time_to_1e-8"
SM_CLOCK_MHZ = 1965.0   # B200 SM clock under these single-/16-SM kernels (the clocks sampler reports the measured one)
TOL = 1e-8


def measured_traffic(workload, kernel, level):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed `ncu --set full` captures
    (profiles/r2_traffic.json, written by tools/profile_r2.sh from the .ncu-rep files); None when there is no capture"""
    p = os.path.join(ROOT, "profiles", "r2_traffic.json")
    try:
        return json.load(open(p)).get(f"{workload}/{kernel}/{level}")
    except Exception:
        return None


def make_config(workload, n_gpus, fast=False):
    """identical in both arms (the driver compares them)"""
    kind, N, eps, desc, _, _ = WORKLOADS[workload]
    big = workload in ("p3d128", "p3d256", "v27_64", "v27_96", "v27_192", "aniso128", "aniso256")
    return {"workload": f"{desc}, V-cycle solve to 1e-8, fp64, b=1, x0=1", "tolerance": TOL,
            "l2_policy": "inputs larger than L2 (hierarchy >> 126 MB); no flush" if big else "hierarchy smaller than L2: L2-resident between steps",
            "parallelism": "single GPU" if n_gpus == 1 else f"level 0 row-block sharded over {n_gpus} GPUs (peer halo exchange); levels >= 1 (one dependency chain per sweep) on rank 0",
            "arithmetic": "FAST" if fast else "EXACT (bit-identical to the reference CPU path)"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_problem(workload):
    from amg_b200 import HostHierarchy, generate
    kind, N, eps = WORKLOADS[workload][:3]
    A = generate(kind, N, eps)
    hier = HostHierarchy(A, tol=TOL)
    return A, hier


def reference_objects_on(hier, n, tol, sample_cycles):
    """cpu_baseline leg of the B200 arm: the reference's own SSS_amg_solve (oracle/_ref, 1 thread) on the hierarchy this run already
    built (byte-identical to the reference's own setup: tests/test_setup_parity.py); falls back to the C restatement (kind "port")
    where oracle/_ref is absent.  sample_cycles: stop after that many V-cycles (bounded sample of a long solve)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import ctypes as C
    import numpy as np
    import oracle_ffi
    from amg_b200 import capi
    x = np.ones(n); bb = np.ones(n)
    if oracle_ffi.have_ref():
        ref = oracle_ffi.Reference("fix")
        hier.mg.pars.max_it = sample_cycles if sample_cycles else 100
        hier.mg.pars.tol = tol
        vx, vb = capi.vec_from_array(x), capi.vec_from_array(bb)
        with oracle_ffi.quiet():
            t0 = time.perf_counter()
            r = ref.L.SSS_amg_solve(C.byref(hier.mg), C.byref(vx), C.byref(vb))
            dt = time.perf_counter() - t0
        hier.mg.cg[0].x = capi.Vec(0, None); hier.mg.cg[0].b = capi.Vec(0, None)
        hier.mg.pars.max_it = 100
        return r, dt, "reference"
    orc = oracle_ffi.Oracle()
    hier.mg.pars.max_it = sample_cycles if sample_cycles else 100
    t0 = time.perf_counter()
    r, _x, _h = orc.solve(hier, x, bb, 0)
    dt = time.perf_counter() - t0
    hier.mg.pars.max_it = 100
    return r, dt, "port"


def reference_arm_problem(workload, tol):
    """--impl reference: the reference's own setup and solve objects on a matrix built by oracle/matgen.c -- libamgb200.so is
    never loaded in this arm.  Returns run(sample_cycles) -> (rtn, seconds), kind, cleanup."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    import oracle_ffi
    from amg_b200 import capi
    from amg_b200.host import CsrMatrix
    kind, N, eps = WORKLOADS[workload][:3]
    rp, ci, va = oracle_ffi.MatGen().generate(kind, N, eps)
    A = CsrMatrix(rp, ci, va)
    n = A.nrows
    if oracle_ffi.have_ref():
        ref = oracle_ffi.Reference("fix")
        mg = ref.setup(A, capi.default_pars(tol))
        def run(sample_cycles):
            r, x, dt = ref.solve_timed(mg, np.ones(n), np.ones(n), tol, max_it=sample_cycles if sample_cycles else 100)
            return r, dt
        return run, "reference", lambda: ref.destroy(mg), n
    # oracle/_ref absent (fresh clone without /root/reference): the C restatement on the product's host setup
    from amg_b200 import HostHierarchy
    orc = oracle_ffi.Oracle()
    hier = HostHierarchy(A, tol=tol)
    def run(sample_cycles):
        hier.mg.pars.max_it = sample_cycles if sample_cycles else 100
        t0 = time.perf_counter()
        r, x, h = orc.solve(hier, np.ones(n), np.ones(n), 0)
        return r, time.perf_counter() - t0
    return run, "port", lambda: None, n


_REAL_STDOUT = None


def quiet_stdout():
    """everything other than the result line (NCCL's version banner, the library's iteration tables, ...) goes to stderr:
    stdout carries exactly ONE JSON line"""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    try:
        C.CDLL(None).fflush(None)
    except Exception:
        pass
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def single_gpu_report(dev, hier, workload, n, hbm, hbm_src):
    """per-level kernel shares of one profiled solve (CUDA events around every phase: adds syncs, not part of any timed number),
    the dominant kernel against its bound, and the HBM-bound level-0 kernels against the measured copy peak"""
    import numpy as np
    dev.set_profile(1)
    r2, _, hist = dev.solve(np.ones(n), np.ones(n))
    dev.set_profile(0)
    phase = dev.phase_ms()
    kernels = []
    sweeps = hier.pars.pre_iter + hier.pars.post_iter
    for l in range(dev.num_levels):
        info = dev.info(l)
        lm = dev.level_ms(l)
        if l < dev.num_levels - 1 and lm[0] > 0:
            kname = dev.gs_kernel(l)
            launches_l = r2.nits * 2 * (1 if kname != "gs_pass_kernel" else 2 * hier.pars.pre_iter)
            per_cycle_bytes = sweeps * dev.bytes(l, 0)
            # latency floor of the ordered sweeps: chain terms per sweep x 8.1 cycles (measured fp64 add latency) at the SM clock
            floor_ms = dev.chain_terms(l) * 8.1 / (SM_CLOCK_MHZ * 1e3) * sweeps * r2.nits
            kernels.append({"kernel": kname, "level": l, "rows": info["rows"], "nnz": info["nnz"],
                            "wavefronts": info["wf_F"] + info["wf_C"], "ms_per_solve": lm[0], "share": lm[0] / phase[6],
                            "gbs": per_cycle_bytes * r2.nits / lm[0] / 1e6, "frac_of_hbm": per_cycle_bytes * r2.nits / lm[0] / 1e6 / hbm,
                            "launches_per_solve": launches_l,
                            "us_per_wavefront": 1e3 * lm[0] / (sweeps * r2.nits * max(1, info["wf_F"] + info["wf_C"])),
                            "chain_terms_per_sweep": dev.chain_terms(l), "chain_floor_ms_per_solve": floor_ms,
                            "chain_floor_frac": (floor_ms / lm[0]) if lm[0] > 0 and floor_ms > 0 else None})
    kernels.sort(key=lambda k: -k["ms_per_solve"])
    top = kernels[0]
    ordered = top["kernel"] != "gs_pass_kernel"
    # the HBM-bound kernels of the path (level 0; the north-star's >= 70 % target applies to these)
    # (resid_restrict = what the cycle runs between pre-smoothing and the next level -- one fused launch where the level has one --
    # against the FUSED byte count of SURVEY.md 8d: S(A) + 16 n + S(R) + 8 n_c + 8 n_c, r not counted)
    ops = {"gs_sweep": 0, "residual": 1, "restrict": 2, "resid_restrict": 6, "prolong": 3, "spmv": 4}
    hbm_kernels = {}
    for name, op in ops.items():
        ms = dev.time_op(0, op, 20)
        if ms > 0:
            gbs = dev.bytes(0, op) / ms / 1e6
            kname = dev.gs_kernel(0) if op == 0 else "spmv_kernel"
            if op == 6:
                kname = "resid_restrict_kernel" if dev.fused(0) else "spmv_kernel x2 + memset"
            hbm_kernels[name] = {"kernel": kname, "gbs": gbs, "frac": gbs / hbm, "ms": ms,
                                 "algorithmic_bytes": dev.bytes(0, op),
                                 "traffic": measured_traffic(workload, name, 0)}
    roofline = {"bound": "latency (fp64 dependency chain of the reference's row order)" if ordered else "hbm",
                "kernel": f"{top['kernel']} (level {top['level']})", "achieved": top["gbs"], "peak": hbm, "unit": "GB/s",
                "frac": top["gbs"] / hbm, "traffic": measured_traffic(workload, top["kernel"], top["level"]),
                "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, profiles/r2_traffic.json)",
                "peak_source": hbm_src, "share_of_step": top["share"],
                "avg_launch_ms": top["ms_per_solve"] / top["launches_per_solve"],
                "algorithmic_bytes_per_launch": top["gbs"] * 1e6 * top["ms_per_solve"] / top["launches_per_solve"],
                "chain_floor_frac": top["chain_floor_frac"],
                "note": ("the dominant kernel is an ordered Gauss-Seidel sweep: its bound is the dependency chain of the reference's row order "
                         "(chain_floor_frac = sum over wavefronts of the longest in-order chain x 8.1 cycles / measured), not HBM; "
                         "hbm_kernels lists the HBM-bound kernels of the path (level 0) against the measured copy peak") if ordered else
                        "hbm_kernels lists the other HBM-bound kernels of the path (level 0)",
                "hbm_kernels": hbm_kernels}
    phases = {"gs": phase[0], "residual": phase[1], "restrict": phase[2], "prolong": phase[3], "coarse_solve": phase[4],
              "outer_residual": phase[5], "total_profiled": phase[6]}
    return kernels, roofline, phases, r2


def parse_library_breakdown(text):
    import re
    ups = re.findall(r"schedule analysis ([0-9.]+) s, layout build ([0-9.]+) s, cudaMalloc\+H2D ([0-9.]+) s, total ([0-9.]+) s", text)
    sol = re.findall(r"AMG solve time: ([0-9.eE+-]+) s", text)
    whole = re.findall(r"upload ([0-9.eE+-]+) s, release ([0-9.eE+-]+) s, whole call ([0-9.eE+-]+) s", text)
    out = []
    for u, so, wh in zip(ups, sol, whole):
        out.append({"analysis_ms": 1e3 * float(u[0]), "layout_ms": 1e3 * float(u[1]), "alloc_h2d_ms": 1e3 * float(u[2]),
                    "upload_total_ms": 1e3 * float(wh[0]), "solve_ms": 1e3 * float(so), "release_ms": 1e3 * float(wh[1]),
                    "inside_library_ms": 1e3 * float(wh[2])})
    return out


def e2e_dropin(hier, n, calls):
    """SSS_amg_solve(mg, x, b) with HOST buffers, `calls` + 1 times: schedule analysis + layout build + H2D of the hierarchy + solve +
    D2H inside the timer.  The FIRST call of the process is reported by itself (module load of the cubin, pinned staging ring,
    cudaMalloc of the device pool: what a one-shot `./amg matrix.mtx` run pays); the following calls reuse pool and ring."""
    import tempfile
    import numpy as np
    from amg_b200 import solve_dropin
    os.environ["AMGB200_VERBOSE"] = "2"                    # (the library's own breakdown of every call is parsed below)
    times = []
    rtn = None
    with tempfile.TemporaryFile(mode="w+") as cap:
        saved = os.dup(1); sys.stdout.flush(); os.dup2(cap.fileno(), 1)
        try:
            for rep in range(1 + calls):
                x_host, b_host = np.ones(n), np.ones(n)
                t0 = time.perf_counter()
                rtn, x_e2e = solve_dropin(hier, x_host, b_host)
                times.append(1e3 * (time.perf_counter() - t0))
        finally:
            C.CDLL(None).fflush(None)                      # the library prints through C stdio: drain it into the capture file
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
        cap.seek(0)
        parts = parse_library_breakdown(cap.read())
    del os.environ["AMGB200_VERBOSE"]
    return times, parts, rtn


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: p3d128 (BASELINE configs[1]) on one GPU, p3d256 (configs[2]) on several")
    ap.add_argument("--impl", default="amgb200", choices=["amgb200", "reference"])
    ap.add_argument("--fast", action="store_true", help="FAST arithmetic mode (not parity-exact)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(args.gpus, world)
    workload = args.workload or ("p3d128" if n_gpus == 1 else "p3d256")
    kind, N, eps, desc, known_cycles, ref_sample = WORKLOADS[workload]
    config = make_config(workload, n_gpus, args.fast)

    # ------------------------------------------------------------------ reference arm (CPU): only rank 0 works
    if args.impl == "reference":
        if rank != 0:
            return 0
        run, how, done, n = reference_arm_problem(workload, TOL)
        sample = ref_sample if (ref_sample and known_cycles) else None       # bounded sample of the long solves
        for _ in range(min(args.warmup, 1)):          # one warm-up pass is enough for a CPU loop (bounded run time)
            run(sample)
        times, rtn = [], None
        for _ in range(args.steps):
            rtn, dt = run(sample)
            times.append(dt)
        done()
        cycles = known_cycles if sample else rtn.nits
        ms = 1e3 * sum(times) / len(times) * (cycles / sample if sample else 1.0)
        what = (f"{sample} V-cycles of the {cycles} the solve needs (tests/golden fixture), scaled by {cycles}/{sample}" if sample
                else f"full solve ({rtn.nits} V-cycles)")
        line = {"impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": n_gpus, "steps": args.steps,
                "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config, "vcycles": cycles, "relres": rtn.rres,
                "cpu_baseline": {"value": ms, "unit": "ms", "cores": 1, "kind": how,
                                 "sample": f"{what}, mean of {args.steps}; the reference has no live OpenMP region: 1 of {os.cpu_count()} host cores",
                                 "host_cores_available": os.cpu_count()},
                "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
        return 0

    # ------------------------------------------------------------------ B200 arm
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the host side of the path (schedule analysis, layout packing: part of the e2e
    # number) is OpenMP code, so the ranks share the box's cores instead (set before the OpenMP runtimes load)
    if world > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // world))
    import numpy as np
    import torch
    from amg_b200 import DeviceHierarchy, capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the amgb200 arm has no CPU fallback (use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    A, hier = build_problem(workload)
    n = A.nrows
    hbm, hbm_src = load_peaks()
    E2E_CALLS = 3 if n <= (1 << 22) else 2

    # ---- e2e FIRST (so that its first call is the first use of the library in this process): the reference-facing call with host buffers
    e2e = None
    if not dist:
        e2e_times, e2e_parts, rtn_e2e = e2e_dropin(hier, n, E2E_CALLS)
        first = e2e_parts[0] if e2e_parts else {}
        warm_ms = sum(e2e_times[1:]) / len(e2e_times[1:])
        e2e = {"value": warm_ms, "unit": "ms", "calls": E2E_CALLS, "warmup_calls": 1, "min_ms": min(e2e_times[1:]), "max_ms": max(e2e_times[1:]),
               "per_call_ms": e2e_times[1:], "per_call_breakdown": e2e_parts[1:], "vcycles": rtn_e2e.nits,
               "first_call_ms": e2e_times[0],
               "first_call_breakdown": dict(first, outside_library_ms=e2e_times[0] - first.get("inside_library_ms", 0.0),
                                            note="first SSS_amg_solve of the process (what a one-shot ./amg run pays): alloc_h2d_ms includes cudaMalloc of the "
                                                 "device pool, cudaMallocHost of the 96 MB staging ring and the lazy load of the kernels' cubin; the CUDA context "
                                                 "itself was created before (torch)"),
               "definition": "SSS_amg_solve(mg, x, b) with host buffers: schedule analysis + layout build + H2D of the hierarchy + solve + D2H; "
                             "value = mean of the calls after the first"}

    dev = DeviceHierarchy(hier, device=local_rank, fast=1 if args.fast else None, level0_worker=bool(dist) and rank > 0)
    if e2e is not None:
        e2e["h2d_bytes_per_step"] = dev.device_bytes() + 2 * 8 * n
        e2e["d2h_bytes_per_step"] = 8 * n + 8 * (e2e["vcycles"] + 1)
    x0 = torch.ones(n, dtype=torch.float64, device="cuda")
    b = torch.ones(n, dtype=torch.float64, device="cuda")
    x = torch.empty_like(x0)
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    extra = {}
    if dist:
        # N > 1: level 0 row-block sharded over the ranks with halo exchange, levels >= 1 on rank 0
        from amg_b200.distributed import GpuBackend, ShardedSolver
        # the same problem on ONE GPU (rank 0), so that the strong-scaling efficiency of this line can be computed from the line itself
        single_ms = None
        if rank == 0:
            single_total, _r = dev.bench_solve(x0.data_ptr(), b.data_ptr(), x.data_ptr(), 1, 1)
            single_ms = single_total
        be = GpuBackend(dev, torch)
        if not be.shape()["shardable"]:
            raise SystemExit("bench.py: level 0 of this workload is not two-colour (every pass is a chain of dependent wavefronts): the path does not "
                             "shard across devices under the reference's row order; run --gpus 1")
        sharded = ShardedSolver(be, A, dist, rank, world, hier.pars.pre_iter, hier.pars.post_iter)
        ones = np.ones(n)
        launches0 = capi.lib().amgb200_launch_count()
        # `value`: inputs resident in HBM when the timed region starts, the solution stays on the device (the N = 1 definition)
        sumb = float(np.sqrt(n))
        for _ in range(args.warmup):
            nits, hist = sharded.solve_resident(x0, b, sumb, TOL, x)
        dist.barrier()
        torch.cuda.synchronize()
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            nits, hist = sharded.solve_resident(x0, b, sumb, TOL, x)
        ev1.record()
        torch.cuda.synchronize()
        dist.barrier()
        t = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
        clocks = sampler.stop()
        launches = capi.lib().amgb200_launch_count() - launches0
        rtn = capi.Rtn(float(hist[-1]), float(hist[-1]) / float(np.sqrt(n)), int(nits))
        # "SpMV GB/s vs HBM peak at N GPUs": every rank times its own row-block of the level-0 residual and of one Gauss-Seidel
        # sweep (halo exchange excluded from numerator and time), aggregate = whole-level algorithmic bytes / max over ranks
        def shard_time(fn, reps=20):
            fn(); torch.cuda.synchronize(); dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record(); torch.cuda.synchronize()
            tt = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())
        fa, fb = sharded.part.f_items[rank]
        ca, cb = sharded.part.c_items[rank]
        nF_items = sharded.sh["itemsF"]
        def own_residual():
            be.residual(fa, fb); be.residual(nF_items + ca, nF_items + cb)
        def own_sweep():
            be.gs_pass(0, fa, fb); be.gs_pass(1, ca, cb)
        shard_ms = {"residual": shard_time(own_residual), "gs_sweep": shard_time(own_sweep)}
        # e2e with the N = 1 definition: every rank analyses + uploads the host hierarchy, the sharded solve runs from host buffers, x comes
        # back to the host of rank 0; max over ranks, one untimed call first
        e2e_t = []
        for rep in range(1 + 1):
            dist.barrier(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            dev2 = DeviceHierarchy(hier, device=local_rank, level0_worker=rank > 0)
            t1 = time.perf_counter()
            be2 = GpuBackend(dev2, torch)
            sh2 = ShardedSolver(be2, A, dist, rank, world, hier.pars.pre_iter, hier.pars.post_iter)
            torch.cuda.synchronize()
            t2 = time.perf_counter()
            nits2, hist2, x2 = sh2.solve(ones, ones, TOL)
            torch.cuda.synchronize()
            t3 = time.perf_counter()
            dt = torch.tensor([1e3 * (t3 - t0), 1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)], dtype=torch.float64, device="cuda")
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            if rep:
                e2e_t.append(float(dt[0].item()))
                e2e_parts = {"upload_ms": float(dt[1].item()), "partition_ghost_lists_peer_setup_ms": float(dt[2].item()), "solve_from_host_buffers_ms": float(dt[3].item())}
            del sh2, be2
            dev2.close()
        e2e = {"value": sum(e2e_t) / len(e2e_t), "unit": "ms", "h2d_bytes_per_step": dev.device_bytes() + 2 * 8 * n, "d2h_bytes_per_step": 8 * n,
               "calls": len(e2e_t), "warmup_calls": 1, "vcycles": int(nits2), "breakdown_max_over_ranks": e2e_parts,
               "definition": "per rank: schedule analysis + layout build + H2D of the hierarchy + partition / ghost lists, then the sharded solve from host "
                             "buffers and x back on the host of rank 0; max over ranks (same content as the N = 1 e2e)"}
        ms_step = ms_total / args.steps
        amdahl_l0_ms = None
        extra = {"single_gpu_ms": single_ms,
                 "strong_scaling_efficiency": (single_ms / ms_step / world) if single_ms else None,
                 "halo_bytes_per_exchange_per_rank": sharded.halo_bytes,
                 "level0_sharded": {"residual_gbs_aggregate": dev.bytes(0, 1) / shard_ms["residual"] / 1e6,
                                    "gs_sweep_gbs_aggregate": dev.bytes(0, 0) / shard_ms["gs_sweep"] / 1e6,
                                    "residual_frac_of_aggregate_hbm": dev.bytes(0, 1) / shard_ms["residual"] / 1e6 / (hbm * world),
                                    "gs_sweep_frac_of_aggregate_hbm": dev.bytes(0, 0) / shard_ms["gs_sweep"] / 1e6 / (hbm * world),
                                    "note": "each rank's own row block of level 0 (halo exchange excluded), whole-level algorithmic bytes / max over ranks"}}
        if rank != 0:
            dist.destroy_process_group()
            return 0
    else:
        launches0 = capi.lib().amgb200_launch_count()
        sampler.start()
        ms_total, rtn = dev.bench_solve(x0.data_ptr(), b.data_ptr(), x.data_ptr(), args.warmup, args.steps)
        torch.cuda.synchronize()
        clocks = sampler.stop()
        launches = capi.lib().amgb200_launch_count() - launches0
        ms_step = ms_total / args.steps
    launches_per_step = launches / (args.warmup + args.steps)

    # ---- one profiled single-GPU solve (rank 0): per-level kernel shares, the dominant kernel, the HBM-bound kernels
    kernels, roofline, phases, r_prof = single_gpu_report(dev, hier, workload, n, hbm, hbm_src)
    if dist:
        l0_ms = sum(k["ms_per_solve"] for k in kernels if k["level"] == 0) + phases["outer_residual"]
        extra["amdahl"] = {"level0_ms_per_solve_single_gpu": l0_ms, "level0_share_of_solve": l0_ms / phases["total_profiled"],
                           "best_possible_speedup_at_this_n_gpus": 1.0 / (1.0 - (l0_ms / phases["total_profiled"]) * (1.0 - 1.0 / world)),
                           "note": "only level 0 shards (two-colour); every other level is one chain of dependent wavefronts per sweep"}
    vcycle_bytes = dev.bytes(0, 5)

    # ---- CPU baseline: the reference's own solve objects on this box (1 thread), bounded sample
    cpu = None
    if not args.no_cpu_baseline:
        sample = ref_sample if (ref_sample and known_cycles) else None
        r_cpu, dt, how = reference_objects_on(hier, n, TOL, sample)
        cycles = known_cycles if sample else r_cpu.nits
        scale = cycles / sample if sample else 1.0
        cpu = {"value": 1e3 * dt * scale, "unit": "ms", "cores": 1, "kind": how,
               "sample": (f"{sample} V-cycles of the reference's SSS_amg_solve, scaled by {cycles}/{sample} (the solve needs {cycles}: tests/golden fixture)" if sample
                          else f"one full solve of the same workload ({r_cpu.nits} V-cycles)") + f" on 1 of {os.cpu_count()} host cores (the reference has no live OpenMP region)",
               "vcycles": cycles, "ares_of_sample": r_cpu.ares}

    line = {"metric": METRIC, "value": ms_step, "unit": "ms", "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config,
            "vcycles": rtn.nits, "relres": rtn.rres, "ares": rtn.ares, "ms_per_vcycle": ms_step / max(1, rtn.nits),
            "vcycle_algorithmic_gb": vcycle_bytes / 1e9, "vcycle_gbs": vcycle_bytes * rtn.nits / ms_step / 1e6,
            "vcycle_frac_of_hbm": vcycle_bytes * rtn.nits / ms_step / 1e6 / (hbm * (world if dist else 1)),
            "clocks": clocks, "e2e": e2e,
            "gpu_launches": int(round(launches_per_step * args.steps)), "gpu_launches_per_step": launches_per_step,
            "roofline": roofline, "kernels": kernels[:8], "phase_ms_per_solve": phases,
            "cpu_baseline": cpu}
    line.update(extra)
    emit(line)
    if dist:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
