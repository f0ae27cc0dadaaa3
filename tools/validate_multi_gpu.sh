# multi-GPU validation (run with gpurun --gpus N): multi-device tests, then bench.py under torchrun at 2..N
mkdir -p gpurun_out
N=${1:-2}
L=gpurun_out/multi_n$N.log; : > $L
nvidia-smi -L >> $L 2>&1
(timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -5) >> $L 2>&1
for n in 2 4 8; do
  [ $n -le $N ] || continue
  echo "== bench N=$n" >> $L
  (timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 6 --warmup 3 > gpurun_out/multi_bench_n$n.json 2>> gpurun_out/multi_bench.err)
  tail -c 2500 gpurun_out/multi_bench_n$n.json >> $L
  echo >> $L
done
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/multi_bench_reference_n$N.json 2>> gpurun_out/multi_bench.err)
tail -5 gpurun_out/multi_bench.err >> $L
cat $L
