cd $GRAFT_REPO_ROOT
timeout 600 python bench.py > gpurun_out/round2_bench.json 2> gpurun_out/round2_bench.err
for c in c1 c3 c5; do timeout 600 python bench.py --config $c --steps 100 > gpurun_out/round2_bench_$c.json 2> gpurun_out/round2_bench_$c.err; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/round2_bench_reference_arm.json 2> gpurun_out/round2_ref.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
