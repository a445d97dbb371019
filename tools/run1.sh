for a in 11 12 15; do AMGB200_LIB=$PWD/build_tl/libamgb200_abl$a.so python tools/sweep.py p3d 128 4,5,6 2>&1 | tail -1; done
