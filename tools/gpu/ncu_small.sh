#!/bin/bash
# ncu --set full pages of the small streaming kernels of the training step (combine, delta, gate backward)
ncu --set full --clock-control none --import-source on -k regex:"combine_fast|bwd_delta16|gate_bwd_fast" -s 3 -c 3 -o gpurun_out/prof_small python tools/prof_bwd.py 2048 8 > gpurun_out/ncu_small.log 2>&1
ncu -i gpurun_out/prof_small.ncu-rep --page raw --csv > gpurun_out/prof_small_raw.csv 2>/dev/null
exit 0
