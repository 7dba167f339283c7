set -x
python -m pytest tests/test_graph_chain_gpu.py -x -q > gpurun_out/graph_tests.log 2>&1; echo "graph tests rc=$?"; tail -30 gpurun_out/graph_tests.log
python -m pytest tests -m gpu -x -q > gpurun_out/r01e_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r01e_tests.log
python scripts/time_small_batch.py 8 50 128 250 > gpurun_out/small_batch_graph.txt 2>&1; cat gpurun_out/small_batch_graph.txt
