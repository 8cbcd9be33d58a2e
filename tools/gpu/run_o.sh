#!/bin/bash
cd $GRAFT_REPO_ROOT
for u in "" 1; do
  echo "STAGE_U8=$u" >> gpurun_out/o_stage.log
  STAGE_U8=$u timeout 120 python tools/stage_bench.py build 30 >> gpurun_out/o_stage.log 2>&1
  STAGE_U8=$u timeout 120 python tools/stage_bench.py step 30 >> gpurun_out/o_stage.log 2>&1
done
