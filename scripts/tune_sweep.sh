mkdir -p gpurun_out
for w in odometry_gpu sub_mapping_gpu livox_stress global_mapping_gpu; do
 for r in 1 2 4 8; do
   GB_TABLE_MULT=$r timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 30 > gpurun_out/sw.log 2>&1
   python -c "
import json
d=json.loads(open('gpurun_out/sw.log').read().strip().splitlines()[-1]); print('$w table_mult=$r', round(d['value']), round(d['roofline']['launch_ms']*1000,1),'us')" 2>/dev/null || (echo "$w $r FAILED"; tail -3 gpurun_out/sw.log)
 done
done
