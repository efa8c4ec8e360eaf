mkdir -p gpurun_out
for N in 8 4 2 1; do
  if [ $N -eq 1 ]; then timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 --no-other-workloads --no-cpu-baseline > gpurun_out/scale_n$N.log 2>&1; else
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/scale_n$N.log 2>&1; fi
  echo "n=$N rc=$?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/scale_n$N.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],4),'kernel ms/rank',[round(x,3) for x in d['kernel_ms_per_rank']['all']],'parity',d['parity_check']['ok'],d['parity_check']['max_rel_H'],'checksum',d['slab_checksum'], d['clocks'])
else:
    print(open('gpurun_out/scale_n$N.log').read()[-2000:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 20 --warmup 5 --collective nccl --verify 0 > gpurun_out/scale_n8_nccl.log 2>&1; python -c "
import json
d=json.loads([x for x in open('gpurun_out/scale_n8_nccl.log') if x.startswith('{')][-1]); print('N8 nccl value',round(d['value']),'ms',round(d['ms_per_step'],4),'checksum',d['slab_checksum'])"
