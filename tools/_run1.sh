cd $GRAFT_REPO_ROOT
for v in base ps5 ps6; do
  if [ $v = base ]; then unset SLAMKLT_LIB; else export SLAMKLT_LIB=$GRAFT_REPO_ROOT/slam.jl_b200/csrc/variants/libslamklt_$v.so; fi
  for i in 1 2; do echo "$v: $(timeout 120 python tools/stage_bench.py build 20 2>&1 | tail -1)"; done
  echo "$v u8: $(STAGE_U8=1 timeout 120 python tools/stage_bench.py build 20 2>&1 | tail -1)"
done > gpurun_out/z_ps.log 2>&1
unset SLAMKLT_LIB
timeout 300 python bench.py --config c5 --steps 50 --no-cpu > gpurun_out/z_c5_k20.json 2> gpurun_out/z_c5_k20.err
timeout 900 python -m pytest tests/test_gpu_large.py tests/test_gpu_configs.py tests/test_gpu_parity.py -q -m gpu -x -k "two_warps or 1080 or prefix_planes or large_frames" 2>&1 | tail -3 >> gpurun_out/z_ps.log
