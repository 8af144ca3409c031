#!/bin/bash
python -m pytest tests/test_attention_gpu.py tests/test_tc_gpu.py tests/test_module_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_cmb9.log
python tools/prof_module.py 65536 1 > gpurun_out/prof_module7.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:combine_fast -s 1 -c 1 -o gpurun_out/prof_combine python tools/prof_module.py 65536 1 > gpurun_out/ncu_combine.log 2>&1
ncu -i gpurun_out/prof_combine.ncu-rep --page raw --csv > gpurun_out/prof_combine_raw.csv 2>/dev/null
exit 0
