python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514 tools/run_sharded.py p3d 64 2>&1 | grep -E "GPUs\]|Error|error|assert" | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 4 --workload p3d128 --steps 1 --warmup 1 > gpurun_out/r2_bench_n4_p3d128.json 2> gpurun_out/r2_bench_n4_p3d128.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n4_p3d128.json"))
print({k: d.get(k) for k in ("value", "n_gpus", "single_gpu_ms", "strong_scaling_efficiency", "vcycles")}, "e2e", d["e2e"]["value"], d["level0_sharded"])
PY
tail -3 gpurun_out/r2_bench_n4_p3d128.err
