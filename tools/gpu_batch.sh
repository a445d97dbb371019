export AMGB200_SETUP_TIMING=1
timeout 600 python tools/setup_probe.py p3d 128 2>&1 | grep -E "===|rap_device" | cut -c1-200
timeout 900 python tools/setup_probe.py p3d 256 2>&1 | grep -E "===|rap_device|\[setup\] level" | cut -c1-200
timeout 900 python tools/setup_probe.py v27 96 2>&1 | grep -E "===|rap_device" | cut -c1-200
