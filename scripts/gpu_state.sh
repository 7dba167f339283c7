# current state of the round on one GPU: all GPU tests, executor per-launch profile, the driver's bench line
set -x
TAG=${1:-s3}
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/tall_$TAG.log 2>&1; echo "all tests rc=$?"; tail -4 gpurun_out/tall_$TAG.log
timeout 300 python scripts/exec_profile.py 1024 3 > gpurun_out/exec_profile_$TAG.txt 2>&1; tail -22 gpurun_out/exec_profile_$TAG.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$TAG.err; cut -c1-600 gpurun_out/bench_$TAG.json
