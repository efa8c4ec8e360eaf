mkdir -p gpurun_out
for w in odometry_gpu sub_mapping_gpu livox_stress global_mapping_gpu single_pair; do
 for r in 0 1; do
   GB_NO_REORDER=$r timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 30 > gpurun_out/sw.log 2>&1
   python -c "
import json
d=json.loads(open('gpurun_out/sw.log').read().strip().splitlines()[-1]); print('$w no_reorder=$r', round(d['value']), round(d['roofline']['launch_ms']*1000,1),'us', 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']))" 2>/dev/null || (echo "$w $r FAILED"; tail -3 gpurun_out/sw.log)
 done
done
