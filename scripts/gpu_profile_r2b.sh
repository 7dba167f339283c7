# ncu evidence of the final kernels of round 2 (one GPU): launch list of one reverse step (time + DRAM bytes + tensor-pipe activity per
# launch) and one --set full capture per kernel that changed since r02_ncu_full_B1024.  Numbers printed under ncu are never bench values.
set -x
TAG=${1:-r02b}
python scripts/one_step.py 1024 > gpurun_out/plain_step.log 2>&1 || { tail -5 gpurun_out/plain_step.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
   --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_per_launch_metrics_B1024.csv python scripts/one_step.py 1024 > gpurun_out/ncu_${TAG}_pl.log 2>&1; echo "per-launch rc=$?"
# conv_gemm2_kernel launch 0 of the step = the 16x128 qkv projection (16 epilogue warps)
for spec in stem_conv7_tc_kernel:0 conv_row64_kernel:0 conv_gemm2_kernel:0 conv_row2_kernel:2 gn_silu_head_kernel:0; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s $skip -c 1 \
      -o gpurun_out/${TAG}_full_$k -f python scripts/one_step.py 1024 > gpurun_out/ncu_${TAG}_$k.log 2>&1; echo "$k rc=$?"
done
du -sh gpurun_out
