#!/bin/bash
# final evidence batch: smoke, default bench, reference arm, ncu launch list of the bench command, module profiles
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1
python bench.py > gpurun_out/bench21.log 2> gpurun_out/bench21.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench21_ref.log 2> gpurun_out/bench21_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_v5.csv python bench.py --steps 5 --no-cpu > gpurun_out/ncu_v5.log 2>&1
python tools/prof_module.py 65536 1 > gpurun_out/prof_module4.log 2>&1
python tools/prof_module_decode.py 4096 592 > gpurun_out/prof_mdec4.log 2>&1
exit 0
