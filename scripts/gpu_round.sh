set -x
TAG=${1:-x}
python -m pytest tests/test_datagen_gpu.py -q -x > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_$TAG.log
python scripts/one_step.py 1024 > gpurun_out/plain_step.log 2>&1 || exit 1
timeout 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
   --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_per_launch_metrics.csv python scripts/one_step.py 1024 > gpurun_out/ncu_${TAG}_pl.log 2>&1; echo "per-launch rc=$?"
for spec in channel_layernorm_h:0 stem_im2col:0 head_conv1:0; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s $skip -c 1 \
      -o gpurun_out/${TAG}_$k -f python scripts/one_step.py 1024 > gpurun_out/ncu_${TAG}_$k.log 2>&1; echo "$k rc=$?"
done
du -sh gpurun_out
