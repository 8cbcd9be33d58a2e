#!/bin/bash
cd $GRAFT_REPO_ROOT
for lag in 0 1 2; do
  SLAMKLT_FUSED=$lag timeout 600 python -m pytest tests/test_gpu_batch.py tests/test_gpu_configs.py -x -q > gpurun_out/c_pytest_$lag.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/c_pytest_$lag.log
  SLAMKLT_FUSED=$lag timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/c_b$lag.json 2> gpurun_out/c_b$lag.err
done
SLAMKLT_FUSED=1 SLAMKLT_FUSED_RING=5 timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/c_b1r5.json 2> gpurun_out/c_b1r5.err
SLAMKLT_FUSED=1 SLAMKLT_NO_L2_WINDOW=1 timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/c_b1nw.json 2> gpurun_out/c_b1nw.err
