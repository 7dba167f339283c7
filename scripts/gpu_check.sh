set -x
python scripts/eps_error.py > gpurun_out/eps_error.txt 2>&1; cat gpurun_out/eps_error.txt | tail -12
python scripts/time_unet.py 1024 > gpurun_out/time_unet.txt 2>&1; SDC_NO_FUSED_UPSAMPLE=1 python scripts/time_unet.py 1024 >> gpurun_out/time_unet.txt 2>&1; cat gpurun_out/time_unet.txt
python bench.py > gpurun_out/r01g_bench.json 2> gpurun_out/r01g_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r01g_bench.json
