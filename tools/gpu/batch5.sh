#!/bin/bash
python tools/sweep.py > gpurun_out/sweep3.jsonl 2> gpurun_out/sweep3.err
ncu --set full --clock-control none --import-source on -k regex:rmsnorm_.*_reg -s 20 -c 3 -o gpurun_out/prof_rmsnorm python tools/bench_block_ops.py > gpurun_out/ncu_rmsnorm.log 2>&1
ncu -i gpurun_out/prof_rmsnorm.ncu-rep --page raw --csv > gpurun_out/prof_rmsnorm_raw.csv 2>/dev/null
exit 0
