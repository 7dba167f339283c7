set -x
python -m pytest tests/test_unet_gpu.py tests/test_unet_bwd_gpu.py tests/test_unet_pgrad_gpu.py -q -x > gpurun_out/t.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t.log
python scripts/time_unet.py 1024 > gpurun_out/time_unet.txt 2>&1; cat gpurun_out/time_unet.txt
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01h.csv python scripts/one_step.py 1024 > gpurun_out/ncu_h.log 2>&1; python scripts/agg_launches.py gpurun_out/launches_r01h.csv | head -30
