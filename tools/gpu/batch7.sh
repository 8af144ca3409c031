#!/bin/bash
python -m pytest tests/test_bwd_tc_gpu.py tests/test_attention_gpu.py tests/test_module_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_bwd7.log
NSA_OPT_FUSED=1 python tools/train_ddp_bench.py --graph --B 8 --steps 10 --warmup 3 > gpurun_out/train_c5_v6.json 2> gpurun_out/train_c5_v6.err
python tools/prof_train.py 8 > gpurun_out/prof_train7.log 2>&1
exit 0
