#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 200 python tools/det_bench.py > gpurun_out/f_det_v2.log 2>&1
SLAMKLT_DETECT_V1=1 timeout 200 python tools/det_bench.py > gpurun_out/f_det_v1.log 2>&1
