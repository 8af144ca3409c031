#!/bin/bash
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 > gpurun_out/pytest_gpu24.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke4.log 2>&1
python bench.py > gpurun_out/bench23.log 2> gpurun_out/bench23.err
exit 0
