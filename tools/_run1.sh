cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_large.py -q -m gpu -x -k 1080p 2>&1 | tail -3 > gpurun_out/z_sweep4.log
SLAMKLT_SWEEP_SEEDS=120 timeout 1500 python -m pytest tests/test_gpu_configs.py -q -m gpu -k "pyramid_parameter_sweep" 2>&1 | tail -15 >> gpurun_out/z_sweep4.log
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:k_cols_all<20, 3" -s 3 -c 1 -f -o gpurun_out/cols_pair_r2c python bench.py --config c5 --steps 4 --warmup 3 --no-cpu > gpurun_out/z_ncu_pair.log 2>&1
