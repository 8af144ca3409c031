#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_gpu23.log
python tools/prof_module_decode.py 4096 592 > gpurun_out/prof_mdec5.log 2>&1
python tools/prof_module_decode.py 4096 1 > gpurun_out/prof_mdec5_b1.log 2>&1
python bench.py --no-cpu --no-train > gpurun_out/bench22.log 2> gpurun_out/bench22.err
exit 0
