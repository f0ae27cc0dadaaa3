mkdir -p gpurun_out
L=gpurun_out/r02d.log; : > $L
export TINSEL_B200_CTA=512
for v in tinsel_b200 tb_v1 tb_v2 tb_v4 tb_v5; do
  for w in 44 56; do
    echo "lib $v walkers $w" >> $L
    TINSEL_B200_LIB=$PWD/tinsel_b200/lib$v.so TINSEL_B200_WALKERS=$w timeout 120 python tools/profile_run.py ajax 1024 1024 16 5 >> $L 2>&1
  done
done
unset TINSEL_B200_CTA
echo "parity, fast-slab variant" >> $L
(TINSEL_B200_LIB=$PWD/tinsel_b200/libtb_v1.so timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "scheduling_modes and offload" 2>&1 | tail -3) >> $L
echo "bvh build tests" >> $L
(timeout 300 python -m pytest tests/test_bvh_build.py -m gpu -x -q 2>&1 | tail -15) >> $L
cat gpurun_out/bvh_build.txt >> $L 2>/dev/null
echo "multi gpu tests (1 GPU)" >> $L
(timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5) >> $L
cat $L
