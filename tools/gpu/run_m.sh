#!/bin/bash
cd $GRAFT_REPO_ROOT
nproc > gpurun_out/m_nproc.txt; free -g | head -2 >> gpurun_out/m_nproc.txt; nvidia-smi -L >> gpurun_out/m_nproc.txt
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 100 --warmup 3 > gpurun_out/m_n$n.json 2> gpurun_out/m_n$n.err
done
timeout 300 python bench.py --no-cpu --steps 100 > gpurun_out/m_n1.json 2> gpurun_out/m_n1.err
