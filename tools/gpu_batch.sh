timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "dataflow" 2>&1 | tail -2
export AMGB200_TIMEOP_SWEEPS=2
python tools/sweep.py p3d 256 1,2,3 "AMGB200_DFW_MIN_WIDTH=100000" 2>&1 | tail -2
python tools/sweep.py aniso3d 256 1,2,3,4 "AMGB200_DFW_MIN_WIDTH=100000" 2>&1 | tail -2
python tools/sweep.py v27 192 0,1,2 "AMGB200_DFW_MIN_WIDTH=100000 AMGB200_NO_DF=1" "AMGB200_DFW_MIN_WIDTH=100" 2>&1 | tail -3
