mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/deferred2_n4.log 2>&1
echo "rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/deferred2_n4.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],4),'kernel ms/rank',[round(x,4) for x in d['kernel_ms_per_rank']['all']],'parity',d['parity_check']['ok'],d['parity_check']['max_rel_H'],'checksum',d['slab_checksum'], d['config']['partition_feedback_kernel_ms'])
else:
    print(open('gpurun_out/deferred2_n4.log').read()[-2000:])
PY
