#!/bin/bash
# launch list of one batch with the final build (after the plain run of the same command exited 0)
set -u
O=gpurun_out/ncu_r2
mkdir -p $O
CMD="python tools/profile_shape.py 296 4.0 4096 12 100"
$CMD > $O/plain_G.log 2>&1 && ncu --clock-control none --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches2.csv $CMD > $O/ncu_G.log 2>&1
tail -1 $O/plain_G.log
