#!/bin/bash
# fourth pass: k_online at the bench shape again (after the packed exact distances), with the per-line and per-SASS
# stall samples kept in full -- the serial section of warp 0 is spread over many lines that a top-60 list hides.
set -u
O=gpurun_out/ncu_r2
mkdir -p $O
NCU="ncu --clock-control none"
export GSC_STREAMS=1
CMD="python tools/profile_shape.py 148 4.0 4096 12 24"
$CMD > $O/plain_F.log 2>&1 && $NCU --set full --import-source on -k regex:'^k_online$' -s 2 -c 1 -f -o $O/k_online2 $CMD > $O/ncu_F.log 2>&1
ncu -i $O/k_online2.ncu-rep --page raw --csv > $O/k_online2_raw.csv 2>/dev/null
ncu -i $O/k_online2.ncu-rep --page source --print-source cuda,sass --csv > /tmp/src.csv 2>/dev/null && python tools/ncu_lines.py /tmp/src.csv 400 > $O/k_online2_lines.txt
ncu -i $O/k_online2.ncu-rep --page source --print-source sass --csv > $O/k_online2_sass.csv 2>/dev/null
rm -f $O/k_online2.ncu-rep
ls -la $O | tail -8
