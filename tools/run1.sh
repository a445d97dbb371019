for a in 11 12 13 14 15; do AMGB200_LIB=$PWD/build_tl/libamgb200_abl$a.so python tools/sweep.py p3d 128 4,5,6 2>&1 | tail -1; done
python tools/sweep.py p3d 128 4,5,6 "AMGB200_STREAM_S=16" "AMGB200_STREAM_S=8" "AMGB200_STREAM_S=4" "AMGB200_STREAM_G=2" "AMGB200_STREAM_G=8" "AMGB200_STREAM_G=1" 2>&1 | tail -7
python tools/sweep.py p3d 128 2,3 "AMGB200_XC_S=16" "AMGB200_XC_S=8" "AMGB200_XC_F=2" "AMGB200_XC_F=2 AMGB200_XC_S=16" "AMGB200_XC_P=32" "AMGB200_XC_F=4 AMGB200_XC_S=8" 2>&1 | tail -7
