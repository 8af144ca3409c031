#!/bin/bash
# One parametrised gpurun batch (replaces the per-session scripts): every step writes under gpurun_out/<tag>_*.
#   gpurun --timeout 1500 -- 'bash tools/gpu/run.sh <tag> [steps...]'
# steps: tests | tests:<pytest args> | smoke | bench | bench:<bench args> | ref | launches | ncu:<kernel regex>:<skip>:<count>[:<script>]
tag=$1; shift
mkdir -p gpurun_out
for step in "$@"; do
  case "$step" in
    tests) python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log ;;
    tests:*) python -m pytest ${step#tests:} -x -q -m gpu 2>&1 | tail -25 > gpurun_out/${tag}_pytest_sel.log ;;
    smoke) python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1 ;;
    bench) python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ;;
    bench:*) python bench.py ${step#bench:} > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err ;;
    ref) python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref.json 2> gpurun_out/${tag}_ref.err ;;
    launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
                python bench.py --steps 2 --warmup 3 --no-cpu --no-decode --no-train --no-train-ddp > gpurun_out/${tag}_launches.log 2>&1 ;;
    ncu:*) IFS=: read -r _ kre skip cnt script <<< "$step"
           script=${script:-"bench.py --steps 2 --warmup 3 --no-cpu --no-decode --no-train --no-train-ddp"}
           name=$(echo "$kre" | tr -c 'A-Za-z0-9_' '_')
           ncu --set full --clock-control none --import-source on -k regex:$kre -s ${skip:-0} -c ${cnt:-1} -o gpurun_out/${tag}_ncu_$name \
             python $script > gpurun_out/${tag}_ncu_$name.log 2>&1
           ncu -i gpurun_out/${tag}_ncu_$name.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_${name}_raw.csv 2>/dev/null ;;
    py:*) python ${step#py:} > gpurun_out/${tag}_py.log 2>&1 ;;
    *) echo "unknown step $step" ;;
  esac
done
exit 0
