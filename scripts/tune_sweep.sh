mkdir -p gpurun_out
for w in odometry_gpu sub_mapping_gpu livox_stress global_mapping_gpu; do
 for b in 2 3 4; do
   GB_MIN_BLOCKS=$b timeout 300 python bench.py --workload $w --no-cpu-baseline --steps 30 > gpurun_out/sw.log 2>&1
   python -c "
import json
d=json.loads(open('gpurun_out/sw.log').read().strip().splitlines()[-1]); print('$w min_blocks=$b', round(d['value']), round(d['roofline']['launch_ms']*1000,1),'us')" 2>/dev/null || (echo "$w $b FAILED"; tail -3 gpurun_out/sw.log)
 done
done
