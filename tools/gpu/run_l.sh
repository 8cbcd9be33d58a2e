#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/l_pytest.log
timeout 200 python tools/det_bench.py > gpurun_out/l_det_main.log 2>&1
timeout 300 python bench.py --config c5 --no-cpu --steps 50 > gpurun_out/l_c5.json 2> gpurun_out/l_c5.err
