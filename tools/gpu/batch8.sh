#!/bin/bash
python -m pytest tests/test_attention_gpu.py tests/test_tc_gpu.py tests/test_module_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_cmb8.log
python tools/prof_train.py 8 > gpurun_out/prof_train8.log 2>&1
python tools/prof_module.py 65536 1 > gpurun_out/prof_module6.log 2>&1
exit 0
