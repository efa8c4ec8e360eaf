mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peer or M4 or global" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-other-workloads --no-cpu-baseline > gpurun_out/deferred_n1.log 2>&1; echo rc=$?
python - <<'PY'
import json
l=[x for x in open('gpurun_out/deferred_n1.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],4),'kernel',d['kernel_ms_per_rank']['all'],'parity',d['parity_check']['ok'],d['parity_check']['max_rel_H'],'checksum',d['slab_checksum'])
else: print(open('gpurun_out/deferred_n1.log').read()[-2000:])
PY
