# round 2, call 1: parity of the new kernel, A/B of its configurations, the default bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/nvsmi.txt; nproc > gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/ab_sweep.py > gpurun_out/ab_r02a.txt 2> gpurun_out/ab_r02a.err; echo "ab rc=$?"; grep -v "^#" gpurun_out/ab_r02a.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['workload'][:12].ljust(12), d['config'].ljust(14), str(d['M_pf_s']).rjust(8), str(d['us_per_launch']).rjust(9), d['frac'], d['items_grid'], d['same_as_first'])"
tail -3 gpurun_out/ab_r02a.err
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_default.log
