#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j_pytest.log
timeout 200 python tools/det_bench.py > gpurun_out/j_det_main.log 2>&1
