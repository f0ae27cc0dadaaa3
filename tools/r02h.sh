mkdir -p gpurun_out
L=gpurun_out/r02h.log; : > $L
run() { echo "lib=$1 scene=$2 offload=$3 walkers=$4 cta=$5" >> $L; TINSEL_B200_LIB=$PWD/tinsel_b200/lib$1.so TINSEL_B200_OFFLOAD=$3 TINSEL_B200_WALKERS=$4 TINSEL_B200_CTA=$5 timeout 120 python tools/profile_run.py $2 $6 $7 16 5 >> $L 2>&1; }
run tb_cs cornell 0 0 512 1024 1024
run tb_cs veach 0 0 512 1920 1080
run tb_cs env 0 0 768 2048 2048
run tb_cs ajax 0 0 768 1024 1024
run tb_cs ajax 1 56 512 1024 1024
echo "parity of the SMEM-AoS variant" >> $L
(TINSEL_B200_LIB=$PWD/tinsel_b200/libtb_cs.so timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "per_sample or scheduling" 2>&1 | tail -4) >> $L
cat $L
