# bench.py under torchrun at N=8, 4, 2 on one 8-GPU box (gpurun --gpus 8)
mkdir -p gpurun_out
L=gpurun_out/scale.log; : > $L
for n in 8 4 2; do
  (timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 6 --warmup 3 > gpurun_out/scale_bench_n$n.json 2>> gpurun_out/scale_bench.err)
  python - <<PY >> $L 2>&1
import json
try:
    d=json.loads(open('gpurun_out/scale_bench_n$n.json').read().strip().splitlines()[-1])
    print('N=$n value %.1f ms/step %.3f e2e %.1f ms/call %.4f reduce_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_call'], d['reduce_ms']))
    print('   per-device kernel ms of the last Render():', [x['kernel_ms_last_call'] for x in d['e2e']['per_device']])
except Exception as e:
    print('N=$n parse failed', e)
PY
done
tail -5 gpurun_out/scale_bench.err | grep -v "OMP_NUM_THREADS\|\*\*\*\*" >> $L
cat $L
