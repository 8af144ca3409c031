#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:decode_produce -s 1 -c 1 -o gpurun_out/prof_produce python tools/prof_module.py 65536 1 > gpurun_out/ncu_produce.log 2>&1
ncu -i gpurun_out/prof_produce.ncu-rep --page raw --csv > gpurun_out/prof_produce_raw.csv 2>/dev/null
exit 0
