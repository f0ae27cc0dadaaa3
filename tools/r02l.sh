mkdir -p gpurun_out
L=gpurun_out/r02l.log; : > $L
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_bvh_build.py tests/test_scene_cache.py -m gpu -x -q 2>&1 | tail -5) >> $L 2>&1
run() { echo "scene=$1 treelet=$2 cta=$3" >> $L; TINSEL_B200_TREELET=$2 TINSEL_B200_CTA=$3 timeout 120 python tools/profile_run.py $1 $4 $5 16 5 >> $L 2>&1; }
for t in 1 0; do
  run ajax $t 768 1024 1024
  run ajax $t 512 1024 1024
  run env $t 768 2048 2048
  run meshlight $t 768 1024 1024
  run table $t 768 1024 1024
done
cat $L
