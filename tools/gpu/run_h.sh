#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
timeout 200 python tools/det_bench.py > gpurun_out/h_det_main.log 2>&1
SLAMKLT_LIB=$PWD/slam.jl_b200/csrc/variants/libslamklt_d3.so timeout 200 python tools/det_bench.py > gpurun_out/h_det_d3.log 2>&1
