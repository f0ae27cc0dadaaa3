mkdir -p gpurun_out
L=gpurun_out/r02g.log; : > $L
echo "parity (two-tier slot state, 2048 slots)" >> $L
(timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -6) >> $L
run() { echo "scene=$1 offload=$2 walkers=$3 cta=$4" >> $L; TINSEL_B200_OFFLOAD=$2 TINSEL_B200_WALKERS=$3 TINSEL_B200_CTA=$4 timeout 120 python tools/profile_run.py $1 $5 $6 16 5 >> $L 2>&1; }
run cornell 0 0 512 1024 1024
run veach 0 0 512 1920 1080
run env 0 0 768 2048 2048
run env 0 0 512 2048 2048
run ajax 0 0 768 1024 1024
run ajax 0 0 512 1024 1024
for w in 40 56 72; do run ajax 1 $w 512 1024 1024; done
cat $L
