#!/bin/bash
# development aid: time alternative builds of the library (TINSEL_B200_LIB) on the headline workload
for lib in "$@"; do
  echo "== $lib"
  TINSEL_B200_LIB=$PWD/tinsel_b200/$lib python tools/profile_run.py cornell 1024 1024 32 2>&1 | tail -1
  TINSEL_B200_LIB=$PWD/tinsel_b200/$lib python tools/profile_run.py veach 1024 1024 16 2>&1 | tail -1
  TINSEL_B200_LIB=$PWD/tinsel_b200/$lib python tools/profile_run.py ajax 1024 1024 16 2>&1 | tail -1
done
