#!/bin/bash
# third pass (after the K <= 256 rewrite): k_online_warp at the bench's occupancy (1184 frames = 8 warps per SM), 3 passes,
# and k_assign at a split shard's size.  Each capture only after the plain run of the same command exited 0.
set -u
O=gpurun_out/ncu_r2
mkdir -p $O
NCU="ncu --clock-control none"
raw() { ncu -i $1 --page raw --csv > ${1%.ncu-rep}_raw.csv 2>/dev/null; }
lines() { ncu -i $1 --page source --print-source cuda,sass --csv > /tmp/src.csv 2>/dev/null && python tools/ncu_lines.py /tmp/src.csv 70 > ${1%.ncu-rep}_lines.txt; }
export GSC_STREAMS=1
CMD="python tools/profile_shape.py 1184 4.0 256 8 3"
$CMD > $O/plain_E.log 2>&1 && $NCU --set full --import-source on -k regex:'^k_online_warp$' -c 1 -f -o $O/k_online_warp $CMD > $O/ncu_E.log 2>&1
raw $O/k_online_warp.ncu-rep; lines $O/k_online_warp.ncu-rep
ncu -i $O/k_online_warp.ncu-rep --page source --print-source sass --csv > $O/k_online_warp_sass.csv 2>/dev/null
rm -f $O/k_online_warp.ncu-rep
ls -la $O | head -40
