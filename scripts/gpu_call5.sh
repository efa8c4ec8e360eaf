mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
AB_CONFIGS="v3,v5,v5 pipe1,v5 pipe2,v5 pipe3" timeout 600 python scripts/ab_sweep.py > gpurun_out/ab_r02d.txt 2> gpurun_out/ab_r02d.err; echo "ab rc=$?"; grep -v "^#" gpurun_out/ab_r02d.txt | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['workload'][:12].ljust(12), d['config'].ljust(14), str(d['M_pf_s']).rjust(8), str(d['us_per_launch']).rjust(9), d['frac'], d['items_grid'], d['same_as_first'])"
tail -3 gpurun_out/ab_r02d.err
timeout 300 python bench.py --workload preprocess --steps 50 > gpurun_out/bench_preprocess.log 2>&1; echo "pre rc=$?"; tail -c 2500 gpurun_out/bench_preprocess.log
