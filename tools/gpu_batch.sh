(time python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or vcycle or history") > gpurun_out/r2_pytest_fused.log 2>&1; tail -5 gpurun_out/r2_pytest_fused.log
python tools/fused_probe.py p3d 256 0,1 "AMGB200_NO_FUSED=1" "AMGB200_RR_ALL=1" "AMGB200_RR_CHUNKS=16" "AMGB200_RR_CHUNKS=64" "AMGB200_RR_LAG=3" "AMGB200_RR_LAG=8" 2>&1 | tail -20
python tools/fused_probe.py p3d 128 0,1 "AMGB200_NO_FUSED=1" "AMGB200_RR_ALL=1" "AMGB200_RR_CHUNKS=16" "AMGB200_RR_LAG=40" 2>&1 | tail -12
python tools/fused_probe.py aniso3d 128 0 "AMGB200_NO_FUSED=1" 2>&1 | tail -12
python tools/fused_probe.py p2d 256 0 "AMGB200_NO_FUSED=1" 2>&1 | tail -12
