#!/bin/bash
# round-2 ncu evidence (one gpurun call): launch list of the bench command, --set full captures of the HBM-bound level-0 kernels at
# 256^3 (BASELINE configs[2], incl. the fused residual + restriction launch), of the dominant ordered smoothers at 128^3 (configs[1])
# and of the one-launch coarsest-level CG.  Every command runs plainly first.
set -x
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
$B > gpurun_out/r2_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_launches.csv $B > gpurun_out/r2_ncu_bench.log 2>&1
L0="python tools/prof_ops.py p3d 256 0"
$L0 > gpurun_out/r2_plain_l0_256.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k 'regex:gs_pass_kernel|spmv_kernel|resid_restrict_kernel' -c 24 -o gpurun_out/r2_l0_256 -f $L0 > gpurun_out/r2_ncu_l0_256.log 2>&1
for spec in "6 gs_stream_cta_kernel r2_stream_cta_l6" "1 gs_dataflow_kernel r2_dataflow_l1" "2 gs_dataflow_csr_kernel r2_dataflow_csr_l2" "3 gs_stream_cluster_kernel r2_stream_cluster_l3"; do
  set -- $spec
  T="python tools/prof_ops.py p3d 128 $1 0"
  $T > gpurun_out/r2_plain_$3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o gpurun_out/$3 -f $T > gpurun_out/r2_ncu_$3.log 2>&1
done
C="python tools/prof_coarse.py p3d 128"
$C > gpurun_out/r2_plain_coarse_cg.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:coarse_cg_kernel -s 1 -c 1 -o gpurun_out/r2_coarse_cg -f $C > gpurun_out/r2_ncu_coarse_cg.log 2>&1
# the reports of one call exceed gpurun's 64 MiB return limit: keep their raw pages as CSV (what tools/summarize_profiles.py reads)
for r in gpurun_out/r2_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null && rm -f $r; done
ls -la gpurun_out/*.raw.csv
