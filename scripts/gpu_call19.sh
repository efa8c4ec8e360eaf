mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_default.log') if x.startswith('{')]
if not l: print(open('gpurun_out/bench_default.log').read()[-2000:])
else:
    d=json.loads(l[-1]); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',d['ms_per_step'],'roof',d['roofline']['frac'],d['roofline'].get('dram_frac'),'cpu',d['cpu_baseline'],'parity',d['parity_check']['ok'],d['clocks'],'launches',d['gpu_launches'])
    for k,v in (d.get('roofline_by_workload') or {}).items(): print(k, v)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref rc=$?"; grep '^{' gpurun_out/bench_reference.log | tail -1 | cut -c1-600
timeout 300 python bench.py --workload preprocess --steps 50 > gpurun_out/bench_preprocess.log 2>&1; echo "pre rc=$?"; python -c "
import json
d=json.loads([x for x in open('gpurun_out/bench_preprocess.log') if x.startswith('{')][-1]); print('pre ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'cpu',d['cpu_baseline']['ms_by_threads'],'x',d['speedup_vs_cpu'],d['speedup_vs_cpu_e2e'],d['parity_check'])"
GB_PROFILE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --verify 0 --no-other-workloads > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python -c "
import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
