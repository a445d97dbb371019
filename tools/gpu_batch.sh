python tools/upload_breakdown.py p3d 128 2>&1 | grep -E "^upload|libamgb200:" | tail -4
AMGB200_LOWER_SYNC=1 python tools/upload_breakdown.py p3d 128 2>&1 | grep -E "^upload|libamgb200:" | tail -4
python tools/upload_breakdown.py p3d 256 2>&1 | grep -E "^upload|layout\]|libamgb200:" | tail -70
for spec in "2 gs_dataflow_csr_kernel r2_dataflow_csr_l2" "3 gs_stream_cluster_kernel r2_stream_cluster_l3"; do
  set -- $spec
  T="python tools/prof_ops.py p3d 128 $1 0"
  $T > gpurun_out/r2_plain_$3.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -o gpurun_out/$3 -f $T > gpurun_out/r2_ncu_$3.log 2>&1
done
rm -f gpurun_out/r2_stream_cluster_l2.ncu-rep
