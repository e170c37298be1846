#!/bin/bash
# Round-2 ncu evidence (run under gpurun, one GPU).  Writes into gpurun_out/; summaries are copied to profiles/ by hand.
#   1. plain run (must exit 0), 2. launch list of one steady-state bench (device time per launch),
#   3. ncu --set full of the two hot kernels of the bucketed path (one launch each, steady state).
set -e
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extra"
$CMD > gpurun_out/r02_plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:bkt_(build|partition)' -s 6 -c 2 -o gpurun_out/r02_prof_bkt_final $CMD > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_plain.log | cut -c1-400
