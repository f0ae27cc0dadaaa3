#!/bin/bash
# development aid: one pass over the scenes for each library given (TINSEL_B200_LIB), default scheduling,
# plus veach under both schedulers
for lib in "$@"; do
  echo "== $lib"
  for sc in "cornell 1024 1024 32" "veach 1024 1024 16" "ajax 1024 1024 16" "env 1024 1024 16" "many 1024 1024 16"; do
    [ -f scenes/${sc%% *}.tsnap ] || continue
    TINSEL_B200_LIB=$PWD/tinsel_b200/$lib timeout 90 python tools/profile_run.py $sc 5 2>&1 | tail -1
  done
  for sched in hard free; do
    echo -n "sched=$sched: "; TINSEL_B200_SCHED=$sched TINSEL_B200_LIB=$PWD/tinsel_b200/$lib timeout 90 python tools/profile_run.py veach 1024 1024 16 5 2>&1 | tail -1
    echo -n "sched=$sched: "; TINSEL_B200_SCHED=$sched TINSEL_B200_LIB=$PWD/tinsel_b200/$lib timeout 90 python tools/profile_run.py cornell 1024 1024 32 5 2>&1 | tail -1
  done
done
