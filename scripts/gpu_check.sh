set -x
python -m pytest tests/test_unet_pgrad_gpu.py -q -s > gpurun_out/pgrad_tests.log 2>&1; echo "pgrad tests rc=$?"; tail -40 gpurun_out/pgrad_tests.log
