# The GPU test suite under compute-sanitizer memcheck (slow: the full-size cases are left out).
cd ${GRAFT_REPO_ROOT:-.}
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -x -k "not full_size and not 1080 and not batch64 and not hybrid" > gpurun_out/san_tests.log 2>&1
echo "rc=$?" >> gpurun_out/san_tests.log
