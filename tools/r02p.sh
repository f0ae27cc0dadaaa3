mkdir -p gpurun_out
L=gpurun_out/r02p.log; : > $L
for lib in tinsel_b200 tb_nodrain; do
  echo "== lib $lib" >> $L
  TINSEL_B200_LIB=$PWD/tinsel_b200/lib$lib.so timeout 200 python tools/e2e_probe.py >> $L 2>&1
  for sc in "cornell 1024 1024" "veach 1920 1080" "ajax 1024 1024" "env 2048 2048"; do
    TINSEL_B200_LIB=$PWD/tinsel_b200/lib$lib.so timeout 120 python tools/profile_run.py $sc 16 5 >> $L 2>&1
  done
done
(timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "per_sample or scheduling or ragged or streamed" 2>&1 | tail -3) >> $L
cat $L
