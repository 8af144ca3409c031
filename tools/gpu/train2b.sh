#!/bin/bash
export NSA_OPT_FUSED=1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/train_ddp_bench.py --graph --B 8 --steps 10 --warmup 3 > gpurun_out/train_c5_n2_graph2.json 2> gpurun_out/train_c5_n2_graph2.err
exit 0
