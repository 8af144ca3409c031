#!/bin/bash
export NSA_OPT_FUSED=1
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tools/train_ddp_bench.py --graph --B 8 --steps 10 --warmup 3 > gpurun_out/train_c5_n8_graph.json 2> gpurun_out/train_c5_n8_graph.err
exit 0
