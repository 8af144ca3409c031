#!/bin/bash
# 2-GPU C5 training step: CUDA graph with the flat bf16 gradient all_reduce captured, and eager DDP + bf16_compress_hook beside it
export NSA_OPT_FUSED=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/train_ddp_bench.py --graph --B 8 --steps 10 --warmup 3 > gpurun_out/train_c5_n2_graph.json 2> gpurun_out/train_c5_n2_graph.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/train_ddp_bench.py --B 8 --steps 10 --warmup 3 > gpurun_out/train_c5_n2_ddp.json 2> gpurun_out/train_c5_n2_ddp.err
exit 0
