(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_c.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_c.log
export AMGB200_TIMEOP_SWEEPS=2
python tools/sweep.py v27 96 0 2>&1 | tail -1; AMGB200_LIB=$PWD/build_tl/libamgb200_minb3.so python tools/sweep.py v27 96 0 2>&1 | tail -1
python tools/sweep.py p3d 128 1 2>&1 | tail -1; AMGB200_LIB=$PWD/build_tl/libamgb200_minb3.so python tools/sweep.py p3d 128 1 2>&1 | tail -1
python tools/sweep.py aniso3d 128 0 2>&1 | tail -1; AMGB200_LIB=$PWD/build_tl/libamgb200_minb3.so python tools/sweep.py aniso3d 128 0 2>&1 | tail -1
