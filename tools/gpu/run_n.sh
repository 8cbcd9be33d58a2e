#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 200 python tools/sanitize_small.py > gpurun_out/n_plain.log 2>&1; echo "rc=$?" >> gpurun_out/n_plain.log
timeout 800 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/n_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/n_memcheck.log
timeout 800 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/n_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/n_racecheck.log
