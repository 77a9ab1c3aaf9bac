#!/bin/bash
# compute-sanitizer over every dispatched k_online shape (gsc_api.cu stage_online), the seeding kernel and the whole
# frame pipeline.  Run on the GPU box:  bash tools/sanitize.sh  -> gpurun_out/sanitize/*.log + summary.txt
set -u
OUT=gpurun_out/sanitize
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # name tool args...
    name=$1; tool=$2; shift 2
    timeout 900 $CS --tool $tool --print-limit 20 "$@" > $OUT/${name}_${tool}.log 2>&1
    rc=$?
    echo "$name $tool rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|RESULT' $OUT/${name}_${tool}.log | tr '\n' ' ')" >> $OUT/summary.txt
}
: > $OUT/summary.txt
# K -> shape: 200 <8,4,64>  400 <8,4,128>  1000 <8,8,128>  2000 <8,8,256>  4096 <8,16,256>
for tool in racecheck synccheck; do
    for cfg in "200 0.1 4" "400 0.2 4" "1000 0.3 3" "2000 0.4 3" "4096 0.6 3"; do
        set -- $cfg
        run online_K$1 $tool python tools/mini_stage.py online $1 $2 $3
    done
    run seed_K300 $tool python tools/mini_stage.py seed 300 0.2
    run frame_K256 $tool python tools/mini_stage.py frame 256 0.3 8
    run frame_K1024 $tool python tools/mini_stage.py frame 1024 0.5 12
done
for tool in initcheck memcheck; do
    run online_K4096 $tool python tools/mini_stage.py online 4096 0.6 3
    run seed_K300 $tool python tools/mini_stage.py seed 300 0.2
    run frame_K1024 $tool python tools/mini_stage.py frame 1024 0.5 12
done
cat $OUT/summary.txt
