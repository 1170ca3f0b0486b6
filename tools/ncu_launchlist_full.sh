#!/bin/bash
# Launch list (device time of every kernel, serialised, cold caches) of ONE full step of the bench workload
# (32 streams x 30 s, beam 5, ~110 decode steps, word timing).  Compare SHARES with the bench line, not absolutes.
TAG=${1:-r02}
export B200W_NO_GRAPHS=1
CMD="python tools/profile_step.py --max-new 0"
$CMD > gpurun_out/ncu_plain_full_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_full_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_full_$TAG.csv $CMD > /dev/null 2>&1
ls -la gpurun_out/launches_full_$TAG.csv
