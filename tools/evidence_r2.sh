set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py > gpurun_out/round2_bench.json 2> gpurun_out/round2_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/round2_bench_reference_arm.json 2> gpurun_out/round2_ref.err
for c in c1 c3 c5; do timeout 600 python bench.py --config $c --steps 100 > gpurun_out/round2_bench_$c.json 2> gpurun_out/round2_bench_$c.err; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 72 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 10 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none -s 40 -c 18 -f -o gpurun_out/step_r2 python tools/stage_bench.py step 3 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_lk_tma -s 1 -c 1 -f -o gpurun_out/lk_r2 python tools/stage_bench.py track 2 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_detect_cells2 -s 8 -c 1 -f -o gpurun_out/det_r2m python tools/det_bench.py > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_detect_cells2 -s 2 -c 1 -f -o gpurun_out/det_r2p python tools/det_bench.py > /dev/null 2>&1
ls -la gpurun_out | tail -12
