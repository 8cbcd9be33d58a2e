#!/bin/bash
cd $GRAFT_REPO_ROOT
for v in main s21 s22 s22c; do
  if [ $v = main ]; then unset SLAMKLT_LIB; else export SLAMKLT_LIB=$PWD/slam.jl_b200/csrc/variants/libslamklt_$v.so; fi
  echo "== $v" >> gpurun_out/q_track.log
  for r in 1 2; do timeout 120 python tools/stage_bench.py track 30 >> gpurun_out/q_track.log 2>&1; done
  timeout 120 python tools/stage_bench.py step 30 >> gpurun_out/q_track.log 2>&1
  if [ $v != main ]; then timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/q_pytest_$v.log 2>&1; echo "rc=$?" >> gpurun_out/q_pytest_$v.log; fi
done
