#!/bin/bash
# compute-sanitizer over small invocations of every kernel; logs -> gpurun_out/sanitizer_<tag>/ (copied to profiles/ afterwards)
TAG=${1:-r2a}
OUT=gpurun_out/sanitizer_$TAG
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool case [env]
  local tool=$1 case=$2; shift 2
  echo "=== $tool $case $*" | tee -a $OUT/summary.txt
  env "$@" timeout 900 $CS --tool $tool --print-limit 20 python profiles/sanitize_case.py $case > $OUT/${tool}_${case}${1:+_sb}.log 2>&1
  echo "rc=$?" | tee -a $OUT/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $OUT/${tool}_${case}${1:+_sb}.log | tail -3 | tee -a $OUT/summary.txt
}
for c in tc tc_short tc_vdt fast generic nuts diag; do run memcheck $c; done
run memcheck tc HMC_B200_TC_SUBBLOCKS=2
for c in tc tc_short fast nuts diag; do run racecheck $c; done
run racecheck tc HMC_B200_TC_SUBBLOCKS=2
for c in tc fast; do run synccheck $c; done
