cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 200 --warmup 5 > gpurun_out/round2c_scale_n4.json 2> gpurun_out/round2c_scale_n4.err
