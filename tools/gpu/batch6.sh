#!/bin/bash
python -m pytest tests/test_producers_gpu.py tests/test_module_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_mod7.log
NSA_OPT_FUSED=1 python tools/train_ddp_bench.py --graph --B 8 --steps 10 --warmup 3 > gpurun_out/train_c5_v5.json 2> gpurun_out/train_c5_v5.err
python tools/prof_train.py 8 > gpurun_out/prof_train6.log 2>&1
python tools/prof_module.py 65536 1 > gpurun_out/prof_module5.log 2>&1
python tools/prof_module_decode.py 4096 592 > gpurun_out/prof_mdec6.log 2>&1
exit 0
