cd $GRAFT_REPO_ROOT
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_cols_all<\(int\)20, \(int\)3' -s 3 -c 1 -f -o gpurun_out/cols_pair_r2c python bench.py --config c5 --steps 4 --warmup 3 --no-cpu > gpurun_out/z_ncu_pair.log 2>&1
