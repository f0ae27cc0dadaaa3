# 8-GPU validation: bench.py under torchrun at N=8 and N=4 (run with gpurun --gpus 8)
mkdir -p gpurun_out
L=gpurun_out/r02n.log; : > $L
nvidia-smi -L | wc -l >> $L
for n in 8 4; do
  echo "== bench N=$n" >> $L
  (timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 6 --warmup 3 > gpurun_out/r02n_bench_n$n.json 2>> gpurun_out/r02n_bench.err)
  python - <<PY >> $L 2>&1
import json
try:
    d=json.loads(open('gpurun_out/r02n_bench_n$n.json').read().strip().splitlines()[-1])
    print('N=$n value %.1f ms/step %.3f e2e %.1f ms/call %.4f reduce_ms %.3f cpu %.1f (%d threads)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_call'], d['reduce_ms'], d['cpu_baseline']['value'], d['cpu_baseline']['cores']))
except Exception as e:
    print('parse failed', e)
PY
done
tail -5 gpurun_out/r02n_bench.err | grep -v "OMP_NUM_THREADS\|\*\*\*\*" >> $L
cat $L
