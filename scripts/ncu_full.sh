# ncu --set full captures, ONE launch per kernel (about 40 replay passes each): kernels of one reverse step (B = 1024) and the
# rollout kernel (N = 100k).  Whole-step capture is not feasible (150 launches x 40 passes: > 12 minutes, > 140 MB).
set -x
TAG=${1:-x}
python scripts/one_step.py 1024 > gpurun_out/plain_step.log 2>&1 || exit 1
for spec in gn_silu_h8:3 channel_layernorm:0 linattn_context_mma:0 conv_row2:2 conv_gemm2:30 stem_im2col:0 reverse_step:0 linattn_fold:0; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s $skip -c 1 \
      -o gpurun_out/${TAG}_$k -f python scripts/one_step.py 1024 > gpurun_out/ncu_${TAG}_$k.log 2>&1; echo "$k rc=$?"
done
python scripts/solver_once.py 100000 > gpurun_out/plain_solver.log 2>&1 || exit 1
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:burgers_rollout -c 1 \
    -o gpurun_out/${TAG}_burgers_rollout -f python scripts/solver_once.py 100000 > gpurun_out/ncu_${TAG}_solver.log 2>&1; echo "solver rc=$?"
du -sh gpurun_out
