export AMGB200_TIMEOP_SWEEPS=2
python tools/sweep.py v27 96 0 2>&1 | tail -1; AMGB200_LIB=$PWD/build_tl/libamgb200_minb3.so python tools/sweep.py v27 96 0 2>&1 | tail -1
python tools/sweep.py p3d 128 1 2>&1 | tail -1; AMGB200_LIB=$PWD/build_tl/libamgb200_minb3.so python tools/sweep.py p3d 128 1 2>&1 | tail -1
python tools/sweep.py aniso3d 128 0 2>&1 | tail -1; AMGB200_LIB=$PWD/build_tl/libamgb200_minb3.so python tools/sweep.py aniso3d 128 0 2>&1 | tail -1
export AMGB200_LIB=$PWD/build_tl/libamgb200_tl.so AMGB200_DEBUG_TIMING=1
unset AMGB200_TIMEOP_SWEEPS
for l in 4 7; do echo "=== p3d 256 level $l"; timeout 300 python tools/prof_level.py p3d 256 $l 2>&1 | grep -A3 "timeline\|ms per" | tail -8; done
