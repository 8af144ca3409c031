#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_gpu22.log
NSA_OPT_FUSED=1 python tools/train_ddp_bench.py --graph --B 8 --steps 10 --warmup 3 > gpurun_out/train_c5_v4.json 2> gpurun_out/train_c5_v4.err
python tools/prof_train.py 8 > gpurun_out/prof_train5.log 2>&1
exit 0
