mkdir -p gpurun_out
L=gpurun_out/r02m.log; : > $L
(timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_bvh_build.py tests/test_scene_cache.py tests/test_fuzz.py -m gpu -x -q 2>&1 | tail -5) >> $L 2>&1
run() { echo "scene=$1 bvh4=$2 cta=$3" >> $L; TINSEL_B200_BVH4=$2 TINSEL_B200_CTA=$3 timeout 120 python tools/profile_run.py $1 $4 $5 16 5 >> $L 2>&1; }
for b in 1 0; do
  run ajax $b 768 1024 1024
  run env $b 768 2048 2048
  run meshlight $b 768 1024 1024
  run table $b 768 1024 1024
  run cornell $b 512 1024 1024
done
run ajax 1 512 1024 1024
cat $L
