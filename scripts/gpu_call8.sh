mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1; echo "bench n=$N rc=$?"
python - <<PY
import json
l=[x for x in open('gpurun_out/bench_n$N.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],4),'kernel ms/rank',d['kernel_ms_per_rank'],'parity',d['parity_check'],'checksum',d['slab_checksum'])
else:
    print(open('gpurun_out/bench_n$N.log').read()[-3000:])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref rc=$?"; grep -o '"cores": [0-9]*, "kind": "port", "host_threads_available": [0-9]*' gpurun_out/bench_ref_n$N.log | head -1; grep -o '"value": [0-9.]*' gpurun_out/bench_ref_n$N.log | head -1
timeout 300 python -m pytest tests/test_multi_gpu_box.py -m gpu -q 2>&1 | tail -3
