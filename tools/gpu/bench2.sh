#!/bin/bash
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_n2_v2.log 2> gpurun_out/bench_n2_v2.err
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/bench_n2_ref.log 2> gpurun_out/bench_n2_ref.err
exit 0
