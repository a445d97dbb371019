(time python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_h.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_h.log | head -2
bash tools/final_runs.sh 2>&1 | grep -v "^\s*$" | head -8
bash tools/profile_r2.sh > gpurun_out/r2_profile_script.log 2>&1; tail -8 gpurun_out/r2_profile_script.log
du -sh gpurun_out
