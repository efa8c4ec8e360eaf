mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
AB_CONFIGS="auto,auto notaper" timeout 600 python scripts/ab_sweep.py sub_mapping_gpu livox_stress global_mapping_gpu odometry_gpu > gpurun_out/ab_r02e.txt 2> gpurun_out/ab_r02e.err; echo "ab rc=$?"; grep -v "^#" gpurun_out/ab_r02e.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['workload'][:12].ljust(12), d['config'].ljust(14), str(d['M_pf_s']).rjust(8), str(d['us_per_launch']).rjust(9), d['frac'], d['items_grid'], d['same_as_first'])"
tail -3 gpurun_out/ab_r02e.err
