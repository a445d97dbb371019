./tools/ubench5 | cut -c1-210 | awk 'NR%4==1 || NR%4==0'
(time timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "not full_size and not fixture") > gpurun_out/r2_pytest_bf.log 2>&1; tail -4 gpurun_out/r2_pytest_bf.log
python tools/quick_time.py p3d 128 2>&1 | grep -E "^solve|phases|profiled solve L[0-3]" | cut -c1-150
python tools/quick_time.py p2d 256 2>&1 | grep -E "^solve|phases" | cut -c1-150
python tools/fused_probe.py p3d 128 0 2>&1 | tail -1
