mkdir -p gpurun_out
AB_CONFIGS="auto,auto ipw10,auto ipw16,auto ipw24" timeout 600 python scripts/ab_sweep.py sub_mapping_gpu livox_stress > gpurun_out/ab_r02f.txt 2> gpurun_out/ab_r02f.err; echo "ab rc=$?"; grep -v "^#" gpurun_out/ab_r02f.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['workload'][:12].ljust(12), d['config'].ljust(14), str(d['M_pf_s']).rjust(8), str(d['us_per_launch']).rjust(9), d['frac'], d['items_grid'], d['same_as_first'])"
tail -3 gpurun_out/ab_r02f.err
