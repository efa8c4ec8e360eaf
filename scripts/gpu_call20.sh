mkdir -p gpurun_out
for N in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale2_n$N.log 2>&1
  echo "n=$N rc=$?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/scale2_n$N.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],4),'kernel ms/rank',[round(x,3) for x in d['kernel_ms_per_rank']['all']],'parity',d['parity_check']['ok'],d['parity_check']['max_rel_H'],'checksum',d['slab_checksum'], d['clocks'], d['config']['partition_feedback_kernel_ms'], 'build', d['config']['build_seconds'])
else:
    print(open('gpurun_out/scale2_n$N.log').read()[-2000:])
PY
done
