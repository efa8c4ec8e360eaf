mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; python - <<'PY'
import json
l=[x for x in open('gpurun_out/bench_default.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1])
    print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ms',round(d['ms_per_step'],3),'parity',d['parity_check']['ok'],'cpu',round(d['cpu_baseline']['value']),d['cpu_baseline']['cores'])
    for k,v in d['roofline_by_workload'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
else:
    print(open('gpurun_out/bench_default.log').read()[-2000:])
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref.log
timeout 300 python bench.py --workload preprocess --steps 50 > gpurun_out/bench_preprocess.log 2>&1; echo "pre rc=$?"; python -c "
import json
d=json.loads([x for x in open('gpurun_out/bench_preprocess.log') if x.startswith('{')][-1]); print('pre ms',d['ms_per_step'],'e2e ms',d['e2e']['ms_per_step'],'cpu',d['cpu_baseline']['ms_by_threads'],'x',d['speedup_vs_cpu'],d['speedup_vs_cpu_e2e'],d['parity_check'])"
for cfg in "odometry_gpu v5_odometry" "livox_stress v3_livox" "global_mapping_gpu v3_global"; do
  set -- $cfg
  GB_PROFILE=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_vgicp_sweep -c 1 -f -o gpurun_out/prof_r02b_$2 python bench.py --workload $1 --steps 1 --warmup 3 --no-cpu-baseline --verify 0 --no-other-workloads > gpurun_out/ncu_r02b_$2.log 2>&1; echo "ncu $2 rc=$?"
done
GB_PROFILE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --verify 0 --no-other-workloads > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
