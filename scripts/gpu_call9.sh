mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 400 python scripts/bench_setup.py > gpurun_out/bench_setup.json 2> gpurun_out/bench_setup.err; echo "setup rc=$?"; python -c "
import json
d=json.load(open('gpurun_out/bench_setup.json'))
for k,v in d.items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})" || tail -5 gpurun_out/bench_setup.err
timeout 300 python bench.py --workload preprocess --steps 50 > gpurun_out/bench_preprocess.log 2>&1; echo "pre rc=$?"; python -c "
import json
d=json.loads([x for x in open('gpurun_out/bench_preprocess.log') if x.startswith('{')][-1]); print('pre ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'cpu',d['cpu_baseline']['ms_by_threads'],'x',d['speedup_vs_cpu'],d['speedup_vs_cpu_e2e'],d['parity_check'])"
