#!/bin/bash
# development aid: lane-queue variants on the on-chip free-running scenes
run() { echo -n "$1 $2: "; env $2 TINSEL_B200_LIB=$PWD/tinsel_b200/$1 timeout 90 python tools/profile_run.py $3 1024 1024 $4 5 2>&1 | tail -1; }
for sc in "cornell 32" "glass 16"; do
  run libtinsel_b200.so TINSEL_B200_QUEUES=ring $sc
  for lib in libtinsel_b200.so $LIBS; do
    run $lib X=1 $sc
  done
done
