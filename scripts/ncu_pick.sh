# ncu --set full of selected conv_gemm2 launches of one step: usage ncu_pick.sh TAG name:skip ...
set -x
TAG=$1; shift
python scripts/one_step.py 1024 > gpurun_out/plain_step.log 2>&1 || exit 1
for spec in "$@"; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 150 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s $skip -c 1 \
      -o gpurun_out/${TAG}_${k}_$skip -f python scripts/one_step.py 1024 > gpurun_out/ncu_${TAG}_${k}_$skip.log 2>&1; echo "$k:$skip rc=$?"
done
