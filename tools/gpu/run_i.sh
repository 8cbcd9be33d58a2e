#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_detect_cells2 -s 8 -c 1 -o gpurun_out/det_v4m -f python tools/det_bench.py > gpurun_out/i_ncu.log 2>&1
