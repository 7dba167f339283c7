# N-GPU pass: bench.py both arms under torchrun (the driver's launch line) + BASELINE configs 3/4/5
set -x
N=${1:-2}; TAG=${2:-x}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_ref_${N}gpu.json 2> gpurun_out/${TAG}_ref_${N}gpu.err; echo "ref rc=$?"
$TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err; echo "bench rc=$?"
for c in 3 5 4; do
  $TR scripts/run_configs.py --config $c > gpurun_out/${TAG}_cfg${c}_${N}gpu.json 2> gpurun_out/${TAG}_cfg${c}_${N}gpu.err; echo "cfg$c rc=$?"
done
cut -c1-330 gpurun_out/${TAG}_*_${N}gpu.json
