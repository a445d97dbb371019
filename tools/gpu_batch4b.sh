python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29517 tools/run_sharded.py p3d 64 2>&1 | grep -E "GPUs\]|Error|error|assert|Traceback" | tail -4
