#!/bin/bash
# round-1 ncu evidence (run on the GPU box through gpurun): launch list of the bench command + --set full captures
set -x
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
python tools/prof_level.py p3d 128 6 > gpurun_out/plain_l6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gs_stream_cta_kernel -c 2 -o gpurun_out/r1_stream_cta_l6 \
    python tools/prof_level.py p3d 128 6 > gpurun_out/ncu_l6.log 2>&1
python tools/prof_level.py p3d 128 0 > gpurun_out/plain_l0.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gs_pass_kernel -c 2 -o gpurun_out/r1_gs_pass_l0 \
    python tools/prof_level.py p3d 128 0 > gpurun_out/ncu_l0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:spmv_kernel -c 1 -o gpurun_out/r1_residual_l0 \
    python tools/prof_level.py p3d 128 0 1 > gpurun_out/ncu_l0r.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gs_stream_cluster_kernel -c 1 -o gpurun_out/r1_stream_cluster_l2 \
    python tools/prof_level.py p3d 128 2 > gpurun_out/ncu_l2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gs_ordered_cluster_kernel -c 1 -o gpurun_out/r1_cluster_l1 \
    python tools/prof_level.py p3d 128 1 > gpurun_out/ncu_l1.log 2>&1
ls -la gpurun_out/
