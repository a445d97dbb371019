for a in u2 u4; do AMGB200_LIB=$PWD/build_tl/libamgb200_$a.so python tools/sweep.py p3d 128 4,5,6 "AMGB200_STREAM_G=4" 2>&1 | tail -2; done
python tools/sweep.py p3d 128 4,5,6 2>&1 | tail -1
