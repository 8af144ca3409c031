#!/bin/bash
# one gpurun batch: block-level tests, microbench, C5 training step (foreach / fused AdamW), training profile
python -m pytest tests/test_block_ops_gpu.py tests/test_producers_gpu.py tests/test_module_gpu.py -x -q -m gpu 2>&1 | tail -12 > gpurun_out/pytest_block3.log
python tools/bench_block_ops.py > gpurun_out/bench_block_ops3.json 2> gpurun_out/bench_block_ops.err
python tools/train_ddp_bench.py --graph --B 8 --steps 10 --warmup 3 > gpurun_out/train_c5_proj2.json 2> gpurun_out/train_c5_proj2.err
NSA_OPT_FUSED=1 python tools/train_ddp_bench.py --graph --B 8 --steps 10 --warmup 3 > gpurun_out/train_c5_proj3.json 2> gpurun_out/train_c5_proj3.err
python tools/prof_train.py 8 > gpurun_out/prof_train4.log 2>&1
python tools/prof_module_decode.py 4096 592 > gpurun_out/prof_mdec3.log 2>&1
exit 0
