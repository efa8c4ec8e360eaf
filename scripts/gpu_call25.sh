mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_default.log') if x.startswith('{')]
if not l: print(open('gpurun_out/bench_default.log').read()[-2000:])
else:
    d=json.loads(l[-1]); print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',d['ms_per_step'],'roof',d['roofline']['frac'],d['roofline'].get('dram_frac'),'parity',d['parity_check']['ok'],d['clocks'],'launches',d['gpu_launches'])
    for k,v in (d.get('roofline_by_workload') or {}).items(): print(k, round(v['value']), v['launch_ms'], round(v['frac'],3), round(v['e2e']))
PY
for W in sub_mapping_gpu global_mapping_gpu; do
GB_PROFILE=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_vgicp_sweep -c 1 -f -o gpurun_out/prof_r02c_$W python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --verify 0 --no-other-workloads > gpurun_out/ncu_r02c_$W.log 2>&1; echo "ncu $W rc=$?"
done
python -c "
import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
