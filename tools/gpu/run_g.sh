#!/bin/bash
cd $GRAFT_REPO_ROOT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_detect_cells2 -s 4 -c 2 -o gpurun_out/det_v3 -f python tools/det_bench.py > gpurun_out/g_ncu.log 2>&1
