#!/bin/bash
cd $GRAFT_REPO_ROOT
for s in 16 18 19 20; do
  echo "slots $s" >> gpurun_out/e_track.log
  SLAMKLT_LK_SLOTS=$s timeout 120 python tools/stage_bench.py track 30 >> gpurun_out/e_track.log 2>&1
done
timeout 600 python bench.py > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
