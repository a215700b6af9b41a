#!/bin/bash
# compute-sanitizer over every kernel family at a small size (run on the GPU box): memcheck + racecheck
out=gpurun_out/sanitize.log; : > $out
run() { echo "=== $*" >> $out; timeout 300 compute-sanitizer --tool $1 --print-limit 5 python tools/prof_run.py ${@:2} >> $out 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $out | tail -1; }
for cfg in "fast 3 160 90 2 1" "fast 5 160 90 2 1" "fast 6 160 90 2 1" "fast 1 160 90 1 1" "fast 0 96 54 1 1" "exact 32 96 54 2 1" "exact 8 96 54 2 1" "exact 1 96 54 2 1"; do run memcheck $cfg; done
for cfg in "fast 3 160 90 2 1" "fast 5 160 90 2 1" "fast 6 160 90 2 1" "fast 1 160 90 1 1" "exact 8 96 54 2 1"; do run racecheck $cfg; done
