(time timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "coarse or vcycle or history or fixture or 1138") > gpurun_out/r2_pytest_pcg.log 2>&1; tail -5 gpurun_out/r2_pytest_pcg.log
python tools/quick_time.py p3d 128 2>&1 | grep -E "^solve|phases" | cut -c1-150
AMGB200_PCG_CTAS=148 python tools/quick_time.py p3d 128 2>&1 | grep -E "phases" | cut -c1-150
AMGB200_PCG_CTAS=64 python tools/quick_time.py p3d 128 2>&1 | grep -E "phases" | cut -c1-150
python tools/quick_time.py p2d 256 2>&1 | grep -E "^solve|phases" | cut -c1-150
