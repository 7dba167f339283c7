# usage: bash scripts/gpu_multi_r2.sh N   (under gpurun --gpus N): NCCL rank-invariance tests + the driver's bench launch line
set -x
N=${1:-2}
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_runner_nccl_gpu.py -q -s -m gpu > gpurun_out/r02_nccl_ranks_${N}gpu.log 2>&1; echo "nccl tests rc=$?"; grep -E "world|passed|failed|skipped|Error" gpurun_out/r02_nccl_ranks_${N}gpu.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_${N}gpu.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"])
print("config3", d["config3"]["strict"]["rollouts_per_s"], d["config3"]["fast"]["rollouts_per_s"])
print("config4", d["config4"]); print("config5", d["config5"])
PY
