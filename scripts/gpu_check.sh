set -x
python -m pytest tests -m gpu -q -s > gpurun_out/r01f_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r01f_tests.log

