cd $GRAFT_REPO_ROOT
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/z_san_mem.log 2>&1; echo "rc=$?" >> gpurun_out/z_san_mem.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/z_san_race.log 2>&1; echo "rc=$?" >> gpurun_out/z_san_race.log
