# compute-sanitizer over the op-level and executor tests (memcheck = out-of-bounds / misaligned accesses, racecheck = shared-memory
# hazards, synccheck = barrier misuse).  Small problem sizes only: the tools slow kernels down 10-100x.  Run on the GPU box:
#   gpurun -- 'bash scripts/sanitize.sh'      -> gpurun_out/r02_sanitize_*.log, copied to profiles/ when clean
set -x
SEL='test_solver_gpu.py::test_solve_free_bit_exact_vs_reference_golden tests/test_solver_gpu.py::test_control_trajectories_and_fused_scoring tests/test_conformal_gpu.py tests/test_plan_gpu.py::test_executor_matches_python_schedule[f16-64] tests/test_plan_gpu.py::test_executor_matches_python_schedule[tf32-32] tests/test_unet_gpu.py::test_fold_on_tensor_cores_matches_fp32 tests/test_chain_gpu.py::test_p_mean_variance_matches_p_sample'
TESTS="tests/$SEL"
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 99 --log-file gpurun_out/r02_sanitize_$tool.log \
      python -m pytest $TESTS -x -q -m gpu -p no:cacheprovider > gpurun_out/r02_sanitize_${tool}_pytest.log 2>&1
  echo "$tool rc=$?"; tail -2 gpurun_out/r02_sanitize_${tool}_pytest.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r02_sanitize_$tool.log | tail -3
done
