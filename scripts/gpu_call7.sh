mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for w in odometry_gpu single_pair livox_stress; do
timeout 300 python bench.py --workload $w --steps 30 --no-cpu-baseline > gpurun_out/bench_$w.log 2>&1; echo "$w rc=$?"; python - <<PY
import json
l=[x for x in open('gpurun_out/bench_$w.log') if x.startswith('{')]
if l:
    d=json.loads(l[-1]); print('$w value',round(d['value']),'e2e',round(d['e2e']['value']),'us/launch',round(d['roofline']['launch_ms']*1e3,2),'frac',round(d['roofline']['frac'],3),'parity',d['parity_check'] and d['parity_check']['ok'], d['parity_check'] and d['parity_check']['max_rel_H'])
else:
    print(open('gpurun_out/bench_$w.log').read()[-1500:])
PY
done
GB_GRAPH=0 timeout 300 python bench.py --workload odometry_gpu --steps 30 --no-cpu-baseline --verify 0 > gpurun_out/bench_odo_nograph.log 2>&1; python -c "
import json
d=json.loads([x for x in open('gpurun_out/bench_odo_nograph.log') if x.startswith('{')][-1]); print('odometry GB_GRAPH=0 value',round(d['value']),'e2e',round(d['e2e']['value']))"
