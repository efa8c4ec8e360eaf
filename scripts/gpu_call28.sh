mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_final.log
python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python bench.py --steps 20 --warmup 5 --no-other-workloads --no-cpu-baseline 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'parity',d['parity_check']['ok'],d['config']['parallelism'][:120])"
