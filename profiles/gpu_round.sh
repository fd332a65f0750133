#!/bin/bash
# One GPU-box session: parity tests, a short bench, then the ncu captures of the same commands.
# usage: bash profiles/gpu_round.sh <tag> [what...]   (what: test bench ncu_list ncu_full sanitize)
tag=$1; shift
what=${*:-test bench}
mkdir -p gpurun_out
for w in $what; do
  case $w in
    test) timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_test.log 2>&1; echo "test rc=$?"; tail -5 gpurun_out/${tag}_test.log;;
    bench) timeout 900 python bench.py --steps 30 --warmup 3 --components none --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/${tag}_bench.json;;
    benchfull) timeout 1500 python bench.py > gpurun_out/${tag}_benchfull.json 2> gpurun_out/${tag}_benchfull.err; echo "benchfull rc=$?"; tail -c 6000 gpurun_out/${tag}_benchfull.json;;
    ncu_list) B="python bench.py --steps 2 --warmup 3 --components none --no-cpu-baseline"
       timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches_bench.csv $B > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu_list rc=$?";;
    ncu_full) timeout 300 python profiles/prof_driver.py pool8 > gpurun_out/${tag}_plain_driver.log 2>&1 && \
       timeout 900 ncu --set full --clock-control none --import-source on -k regex:"pool_enum|pool_select" -c 16 -o gpurun_out/${tag}_prof -f \
         python profiles/prof_driver.py pool8 > gpurun_out/${tag}_ncu_driver.log 2>&1; echo "ncu_full rc=$?";;
    ts) timeout 300 python profiles/pool_ts.py > gpurun_out/${tag}_ts.log 2>&1; echo "ts rc=$?"; cat gpurun_out/${tag}_ts.log;;
  esac
done
