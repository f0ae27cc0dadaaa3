mkdir -p gpurun_out
L=gpurun_out/r02f.log; : > $L
run() { echo "$1 scene=$2 offload=$3" >> $L; TINSEL_B200_LIB=$PWD/tinsel_b200/lib$1.so TINSEL_B200_OFFLOAD=$3 timeout 120 python tools/profile_run.py $2 $4 $5 16 5 >> $L 2>&1; }
for lib in tinsel_b200 tb_p512; do
  run $lib cornell 0 1024 1024
  run $lib veach 0 1920 1080
  run $lib env 0 2048 2048
  run $lib ajax 0 1024 1024
done
echo "scene cache + multi gpu + finish tests (red.v4 splat)" >> $L
(timeout 600 python -m pytest tests/test_scene_cache.py tests/test_multi_gpu.py tests/test_finish.py tests/test_plugin_gpu.py -m gpu -x -q 2>&1 | tail -5) >> $L
(timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "render_matches or streamed or box_filter" 2>&1 | tail -3) >> $L
cat $L
