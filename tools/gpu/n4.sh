#!/bin/bash
export NSA_OPT_FUSED=1
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 tools/train_ddp_bench.py --graph --B 8 --steps 10 --warmup 3 > gpurun_out/train_c5_n4_graph.json 2> gpurun_out/train_c5_n4_graph.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n4.log 2> gpurun_out/bench_n4.err
exit 0
