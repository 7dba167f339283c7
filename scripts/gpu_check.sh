set -x
python -m pytest tests -m gpu -q > gpurun_out/r01h_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r01h_tests.log
python bench.py > gpurun_out/r01h_bench.json 2> gpurun_out/r01h_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r01h_bench.json
