python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload p3d128 --steps 1 --warmup 1 > gpurun_out/r2_bench_n2_p3d128.json 2> gpurun_out/r2_bench_n2_p3d128.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n2_p3d128.json"))
print({k: d.get(k) for k in ("value", "n_gpus", "single_gpu_ms", "strong_scaling_efficiency", "vcycles")}, "e2e", d["e2e"]["value"], d["e2e"]["breakdown_max_over_ranks"])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n2.json"))
print({k: d.get(k) for k in ("value", "n_gpus", "single_gpu_ms", "strong_scaling_efficiency", "vcycles")}, "e2e", d["e2e"]["value"], d["e2e"]["breakdown_max_over_ranks"], d["level0_sharded"])
PY
