set -x
TAG=${1:-r2a}
timeout 600 python -m pytest tests/test_unet_gpu.py -q -x -k "fused_groupnorm or row_gn_head" > gpurun_out/t_$TAG.log 2>&1; echo "gn tests rc=$?"; tail -15 gpurun_out/t_$TAG.log
timeout 300 python scripts/time_row_gn.py > gpurun_out/time_row_gn_$TAG.txt 2>&1; cat gpurun_out/time_row_gn_$TAG.txt
timeout 300 python scripts/eps_error.py > gpurun_out/eps_error_$TAG.txt 2>&1; cat gpurun_out/eps_error_$TAG.txt | head -8
timeout 300 python scripts/time_unet.py 1024 > gpurun_out/time_unet_$TAG.txt 2>&1; cat gpurun_out/time_unet_$TAG.txt
SDC_FUSE_GN=0 timeout 300 python scripts/time_unet.py 1024 > gpurun_out/time_unet_${TAG}_nofuse.txt 2>&1; cat gpurun_out/time_unet_${TAG}_nofuse.txt
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/tall_$TAG.log 2>&1; echo "all tests rc=$?"; tail -8 gpurun_out/tall_$TAG.log
