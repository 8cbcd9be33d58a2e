# Final evidence of round 2 (after the pair column kernel and the general kernels): tests, smoke, bench lines, c5 ncu captures.
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_smoke.log
timeout 600 python bench.py > gpurun_out/round2_bench.json 2> gpurun_out/round2_bench.err
for c in c1 c3 c5; do timeout 600 python bench.py --config $c --steps 100 > gpurun_out/round2_bench_$c.json 2> gpurun_out/round2_bench_$c.err; done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/round2_bench_reference_arm.json 2> gpurun_out/round2_ref.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 60 --csv --log-file gpurun_out/launches_r2c_c5.csv python bench.py --config c5 --steps 6 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_cols_all -s 12 -c 2 -f -o gpurun_out/cols_pair_r2c python bench.py --config c5 --steps 4 --warmup 3 --no-cpu > /dev/null 2>&1
ls -la gpurun_out | tail -12
