#!/bin/bash
# compute-sanitizer memcheck over every kernel on small workloads (run under gpurun); prints the error summaries
for c in "cornell 96 64 2" "veach 96 64 2" "ajax 96 64 1" "envmini 64 48 2" "table 64 40 1"; do
  [ -f scenes/${c%% *}.tsnap ] || continue
  echo "== memcheck $c"
  timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/profile_run.py $c 1 2>&1 | grep -E "ERROR SUMMARY|Invalid|error|Msamples" | head -5
done
echo "== memcheck finish/nlm/render (tests/test_finish.py, streamed read-back)"
timeout 280 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_finish.py -m gpu -x -q 2>&1 | grep -E "ERROR SUMMARY|passed|failed" | head -4
