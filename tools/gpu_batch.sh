(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_d.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_d.log
python tools/quick_time.py p3d 128 2>&1 | grep -E "profiled solve|^solve|phases" | cut -c1-150
AMGB200_NO_LOWER=1 python tools/quick_time.py p3d 128 2>&1 | grep -E "^solve" | tail -1
export AMGB200_TIMEOP_SWEEPS=2
python tools/sweep.py v27 96 0 2>&1 | tail -1; AMGB200_LIB=$PWD/build_tl/libamgb200_minb3.so python tools/sweep.py v27 96 0 2>&1 | tail -1
python tools/sweep.py p3d 128 1 2>&1 | tail -1; AMGB200_LIB=$PWD/build_tl/libamgb200_minb3.so python tools/sweep.py p3d 128 1 2>&1 | tail -1
python tools/sweep.py p3d 256 3 "AMGB200_DFW_MIN_WIDTH=80" 2>&1 | tail -2
python tools/sweep.py v27 192 3 "AMGB200_DFW_MIN_WIDTH=80" 2>&1 | tail -2
