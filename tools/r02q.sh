mkdir -p gpurun_out
N=${1:-4}
L=gpurun_out/r02q.log; : > $L
(timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -x -q 2>&1 | tail -2) >> $L
(timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/r02q_bench_n$N.json 2>> gpurun_out/r02q_bench.err)
python - <<PY >> $L 2>&1
import json
d=json.loads(open('gpurun_out/r02q_bench_n$N.json').read().strip().splitlines()[-1])
print('N=$N value %.1f ms/step %.3f e2e %.1f ms/call %.4f reduce_ms %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_call'], d['reduce_ms']))
for x in d['e2e']['per_device']: print(x)
PY
tail -3 gpurun_out/r02q_bench.err | grep -v "OMP_NUM\|\*\*\*" >> $L
cat $L
