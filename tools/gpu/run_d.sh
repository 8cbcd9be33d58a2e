#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest.log
for v in main m18 m19; do
  if [ $v = main ]; then unset SLAMKLT_LIB; else export SLAMKLT_LIB=$PWD/slam.jl_b200/csrc/variants/libslamklt_$v.so; fi
  for r in 1 2; do timeout 120 python tools/stage_bench.py track 30 >> gpurun_out/d_track_$v.log 2>&1; done
done
unset SLAMKLT_LIB
timeout 120 python tools/stage_bench.py step 30 >> gpurun_out/d_step_main.log 2>&1
