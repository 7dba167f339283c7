set -x
TAG=${1:-x}
python -m pytest tests/test_unet_gpu.py tests/test_unet_bwd_gpu.py tests/test_unet_pgrad_gpu.py -q -x > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_$TAG.log
python scripts/eps_error.py > gpurun_out/eps_error_$TAG.txt 2>&1; cat gpurun_out/eps_error_$TAG.txt | head -8
python scripts/time_unet.py 1024 > gpurun_out/time_unet_$TAG.txt 2>&1; cat gpurun_out/time_unet_$TAG.txt
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_$TAG.csv python scripts/one_step.py 1024 > gpurun_out/ncu_$TAG.log 2>&1; python scripts/agg_launches.py gpurun_out/launches_$TAG.csv | head -30
