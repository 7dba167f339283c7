set -x
python -m pytest tests/test_graph_chain_gpu.py tests/test_unet_pgrad_gpu.py -q -x > gpurun_out/t.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t.log
python scripts/time_finetune.py 50 > gpurun_out/time_finetune.txt 2>&1; tail -3 gpurun_out/time_finetune.txt
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_gemm2 --launch-skip 1 --launch-count 1 -o gpurun_out/r01_qkv_conv_v2 python scripts/one_step.py 1024 > gpurun_out/ncu_qkv.log 2>&1; tail -2 gpurun_out/ncu_qkv.log
