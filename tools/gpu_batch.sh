export AMGB200_TIMEOP_SWEEPS=2
python tools/sweep.py p3d 256 6,7,8 "AMGB200_STREAM_RELAX=1" 2>&1 | tail -2
python tools/sweep.py v27 192 6,7,8 "AMGB200_STREAM_RELAX=1" 2>&1 | tail -2
