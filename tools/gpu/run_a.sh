#!/bin/bash
cd $GRAFT_REPO_ROOT
nproc > gpurun_out/a_nproc.txt
timeout 600 python -m pytest tests/test_gpu_batch.py -x -q > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/a_b1.json 2> gpurun_out/a_b1.err
SLAMKLT_NO_HYBRID=1 timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/a_b2.json 2> gpurun_out/a_b2.err
SLAMKLT_STEP_CHUNKS=4 timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/a_b3.json 2> gpurun_out/a_b3.err
SLAMKLT_RAW_LOOKAHEAD=8 timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/a_b4.json 2> gpurun_out/a_b4.err
SLAMKLT_STEP_CHUNKS=1 timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/a_b5.json 2> gpurun_out/a_b5.err
