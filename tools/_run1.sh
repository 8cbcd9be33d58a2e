cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_large.py -q -m gpu -x 2>&1 | tail -30 > gpurun_out/z_large.log
timeout 300 python tools/large_probe.py >> gpurun_out/z_large.log 2>&1
