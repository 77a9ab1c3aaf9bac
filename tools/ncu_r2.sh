#!/bin/bash
# Round-2 ncu captures at the bench shape (4 s stereo 44.1 kHz frames, one CTA-per-SM wave = 148 frames).
# Run on the GPU box: bash tools/ncu_r2.sh -> gpurun_out/ncu_r2/*.csv / *.txt (summaries are copied to profiles/ by hand)
set -u
O=gpurun_out/ncu_r2
mkdir -p $O
NCU="ncu --clock-control none"
raw() { ncu -i $1 --page raw --csv > ${1%.ncu-rep}_raw.csv 2>/dev/null; }
lines() { ncu -i $1 --page source --print-source cuda,sass --csv > /tmp/src.csv 2>/dev/null && python tools/ncu_lines.py /tmp/src.csv 45 > ${1%.ncu-rep}_lines.txt; }
# A. launch list of a whole batch (two lanes, sliced k_online)
CMD="python tools/profile_shape.py 296 4.0 4096 12 100"
$CMD > $O/plain_A.log 2>&1 && $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches.csv $CMD > $O/ncu_A.log 2>&1
# B. k_online<8,16,256>: third slice (passes 16..23) of a one-wave batch (GSC_STREAMS=1: one lane, 148 CTAs per launch)
export GSC_STREAMS=1
CMD="python tools/profile_shape.py 148 4.0 4096 12 24"
$CMD > $O/plain_B.log 2>&1 && $NCU --set full --import-source on -k regex:k_onlineILi8ELi16 -s 2 -c 1 -f -o $O/k_online $CMD > $O/ncu_B.log 2>&1
raw $O/k_online.ncu-rep; lines $O/k_online.ncu-rep
# C. k_seed2 (whole seeding of 148 frames)
$NCU --set full --import-source on -k regex:k_seed2 -c 1 -f -o $O/k_seed2 $CMD > $O/ncu_C.log 2>&1
raw $O/k_seed2.ncu-rep; lines $O/k_seed2.ncu-rep
# D. k_assign (Lloyd mode)
CMD="python tools/profile_shape.py 148 4.0 4096 12 1 1"
$CMD > $O/plain_D.log 2>&1 && $NCU --set full --import-source on -k regex:k_assignILi8 -c 1 -f -o $O/k_assign $CMD > $O/ncu_D.log 2>&1
raw $O/k_assign.ncu-rep; lines $O/k_assign.ncu-rep
# E. k_online_warp (K = 256, 8-bit, 3 passes)
CMD="python tools/profile_shape.py 592 4.0 256 8 3"
$CMD > $O/plain_E.log 2>&1 && $NCU --set full --import-source on -k regex:k_online_warp -c 1 -f -o $O/k_online_warp $CMD > $O/ncu_E.log 2>&1
raw $O/k_online_warp.ncu-rep; lines $O/k_online_warp.ncu-rep
# F. the memory-bound / small kernels
CMD="python tools/profile_shape.py 148 4.0 4096 12 8"
$CMD > $O/plain_F.log 2>&1 && $NCU --set full -k regex:"k_knnfit_win|k_find_divider2|k_class_means_members|k_member_sums_f|k_group_labels|k_seed_prep|k_make_chunks|k_pack_frames|k_dictionary|k_finalize|k_knn_prep|k_compact_stream" -c 14 -f -o $O/small $CMD > $O/ncu_F.log 2>&1
raw $O/small.ncu-rep
rm -f $O/k_seed2.ncu-rep $O/k_assign.ncu-rep $O/small.ncu-rep $O/k_online_warp.ncu-rep
ls -la $O
