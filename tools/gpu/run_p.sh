#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest.log
for c in c3 c5; do timeout 600 python bench.py --config $c --steps 100 --no-cpu > gpurun_out/p_$c.json 2> gpurun_out/p_$c.err; done
for v in main cm5; do
  if [ $v = main ]; then unset SLAMKLT_LIB; else export SLAMKLT_LIB=$PWD/slam.jl_b200/csrc/variants/libslamklt_$v.so; fi
  echo $v >> gpurun_out/p_build.log
  timeout 120 python tools/stage_bench.py build 30 >> gpurun_out/p_build.log 2>&1
done
