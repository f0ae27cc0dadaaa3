#!/bin/bash
# development aid: time alternative builds of the library (TINSEL_B200_LIB) x scheduling modes
#   SCHEDS="free" bash tools/variant_bench.sh libA.so libB.so
for lib in "$@"; do
  for sched in ${SCHEDS:-hard free}; do
    echo "== $lib sched=$sched"
    for sc in "cornell 1024 1024 32" "veach 1024 1024 16" "ajax 1024 1024 16"; do
      TINSEL_B200_SCHED=$sched TINSEL_B200_LIB=$PWD/tinsel_b200/$lib timeout 60 python tools/profile_run.py $sc 5 2>&1 | tail -1
    done
  done
done
