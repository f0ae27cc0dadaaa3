mkdir -p gpurun_out
L=gpurun_out/r02e.log; : > $L
export TINSEL_B200_CTA=512
run() { echo "$1 walkers=$2 offload=$3" >> $L; TINSEL_B200_LIB=$PWD/tinsel_b200/lib$1.so TINSEL_B200_WALKERS=$2 TINSEL_B200_OFFLOAD=$3 timeout 120 python tools/profile_run.py ajax 1024 1024 16 5 >> $L 2>&1; }
for w in 56 64 72; do run tinsel_b200 $w 1; done
for w in 48 56 64; do run tb_sw8 $w 1; done
for w in 48 56 64; do run tb_sw12 $w 1; done
run tb_p512 56 1
unset TINSEL_B200_CTA
run tb_p512 0 0
run tinsel_b200 0 0
cat $L
