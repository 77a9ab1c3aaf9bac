#!/bin/bash
# second pass: the two captures whose kernel filter did not match (ncu matches the base name), after the plain runs
set -u
O=gpurun_out/ncu_r2
mkdir -p $O
NCU="ncu --clock-control none"
raw() { ncu -i $1 --page raw --csv > ${1%.ncu-rep}_raw.csv 2>/dev/null; }
lines() { ncu -i $1 --page source --print-source cuda,sass --csv > /tmp/src.csv 2>/dev/null && python tools/ncu_lines.py /tmp/src.csv 60 > ${1%.ncu-rep}_lines.txt; }
export GSC_STREAMS=1
CMD="python tools/profile_shape.py 148 4.0 4096 12 24"
$CMD > $O/plain_B.log 2>&1 && $NCU --set full --import-source on -k regex:'^k_online$' -s 2 -c 1 -f -o $O/k_online $CMD > $O/ncu_B.log 2>&1
raw $O/k_online.ncu-rep; lines $O/k_online.ncu-rep
CMD="python tools/profile_shape.py 148 4.0 4096 12 1 1"
$CMD > $O/plain_D.log 2>&1 && $NCU --set full --import-source on -k regex:'^k_assign$' -c 1 -f -o $O/k_assign $CMD > $O/ncu_D.log 2>&1
raw $O/k_assign.ncu-rep; lines $O/k_assign.ncu-rep
rm -f $O/k_assign.ncu-rep
ls -la $O | head -40
