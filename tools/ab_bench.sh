#!/bin/bash
# development aid: A/B two builds of the library with the per-scene heuristics in effect
#   bash tools/ab_bench.sh libA.so libB.so   (each timed twice, interleaved)
for lib in "$@" "$@"; do
  echo "== $lib"
  for sc in "cornell 1024 1024 32" "veach 1024 1024 16" "ajax 1024 1024 16" "env 1024 1024 16" "meshlight 1024 1024 16"; do
    [ -f scenes/${sc%% *}.tsnap ] || continue
    TINSEL_B200_LIB=$PWD/tinsel_b200/$lib timeout 60 python tools/profile_run.py $sc 5 2>&1 | tail -1
  done
done
