#!/bin/bash
# one bench line per BASELINE configuration (N = 1), kept under gpurun_out/ and summarised into profiles/r2_bench_configs.md
for w in p2d256 p3d128 p3d256 v27_192 aniso256; do
  python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/r2_bench_$w.json 2> gpurun_out/r2_bench_$w.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_$w.json"))
print("$w", round(d["value"], 1), "ms", d["vcycles"], "cycles e2e", round(d["e2e"]["value"], 1), "first", round(d["e2e"]["first_call_ms"], 1), "cpu", round(d["cpu_baseline"]["value"], 1), d["roofline"]["kernel"], d["roofline"]["chain_floor_frac"])
PY
done
