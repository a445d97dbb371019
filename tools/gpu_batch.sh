export AMGB200_TIMEOP_SWEEPS=2
for w in 8 4 2 1; do AMGB200_SMALL_WARPS=$w python tools/sweep.py p2d 256 1,2,3 2>&1 | tail -1; done
AMGB200_LIB=$PWD/build_tl/libamgb200_sleep100.so python tools/sweep.py p2d 256 1,2,3 2>&1 | tail -1
AMGB200_NO_SMALL=1 python tools/sweep.py p2d 256 1,2,3 2>&1 | tail -1
