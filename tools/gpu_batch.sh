export AMGB200_TIMEOP_SWEEPS=2
B=$PWD/build_tl/libamgb200_base.so
for r in 1 2 3; do
python tools/sweep.py p3d 128 1,2 2>&1 | tail -1; AMGB200_LIB=$B python tools/sweep.py p3d 128 1,2 2>&1 | tail -1
done
for r in 1 2; do
python tools/sweep.py aniso3d 128 0,1,2,3 2>&1 | tail -1; AMGB200_LIB=$B python tools/sweep.py aniso3d 128 0,1,2,3 2>&1 | tail -1
python tools/sweep.py v27 96 0,1,2 2>&1 | tail -1; AMGB200_LIB=$B python tools/sweep.py v27 96 0,1,2 2>&1 | tail -1
done
