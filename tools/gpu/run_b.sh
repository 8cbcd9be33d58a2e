#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_batch.py -x -q > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/b_b1.json 2> gpurun_out/b_b1.err
SLAMKLT_TP_SPLIT=2 timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/b_b2.json 2> gpurun_out/b_b2.err
SLAMKLT_NO_HYBRID=1 timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/b_b3.json 2> gpurun_out/b_b3.err
SLAMKLT_TP_SPLIT=3 timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/b_b4.json 2> gpurun_out/b_b4.err
