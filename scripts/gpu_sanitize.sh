# memcheck of the kernels added this session on small problems
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_fd.py tests/test_gpu_kernels.py -q -x -k "golden or low_rank_root_matches_reference or tc_grouped or dynamic_exponent or all_padding" 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed|out of bounds|=========     at" | head -20
echo "fd/ggemm memcheck rc=${PIPESTATUS[0]}"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_root.py -q -x -k "eigh or residual or n1_closed or padding_invariance" 2>&1 | grep -E "ERROR SUMMARY|Invalid|passed|failed|out of bounds|=========     at" | head -20
echo "root memcheck rc=${PIPESTATUS[0]}"
