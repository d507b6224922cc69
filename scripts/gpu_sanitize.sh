# memcheck of the small GPU tests (every kernel family at least once)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_fd.py tests/test_gpu_kernels.py tests/test_gpu_tc.py -q -x -k "golden or low_rank_root_matches_reference or tc_grouped or fused_quant or dynamic_exponent or all_padding or opt_in or exact_on" 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed|out of bounds" | head -12
echo "fd/ggemm/tc memcheck rc=${PIPESTATUS[0]}"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_root.py -q -x -k "eigh or residual or n1_closed or padding_invariance" 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed|out of bounds" | head -12
echo "root memcheck rc=${PIPESTATUS[0]}"
