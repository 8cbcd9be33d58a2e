#!/bin/bash
cd $GRAFT_REPO_ROOT
for v in s21r s20; do
  export SLAMKLT_LIB=$PWD/slam.jl_b200/csrc/variants/libslamklt_$v.so
  echo "== $v" >> gpurun_out/q2_track.log
  for r in 1 2; do timeout 120 python tools/stage_bench.py track 30 >> gpurun_out/q2_track.log 2>&1; done
  SLAMKLT_LK_SLOTS=20 timeout 120 python tools/stage_bench.py track 30 >> gpurun_out/q2_track.log 2>&1
  timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/q2_pytest_$v.log 2>&1; echo "rc=$?" >> gpurun_out/q2_pytest_$v.log
done
