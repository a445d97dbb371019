python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/run_sharded.py p3d 64 2>&1 | grep -E "GPUs\]|Error|error|assert|Traceback" | tail -5
AMGB200_NO_OVERLAP=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29516 tools/run_sharded.py p3d 64 2>&1 | grep -E "GPUs\]" | tail -2
