mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
AB_CONFIGS="pre=,nopre=GB_PREFETCH:0,pre1024=GB_TILE:1024,pre1536=GB_TILE:1536" timeout 900 python scripts/shard_emulate.py 1 8 > gpurun_out/shard_emulate4.txt 2> gpurun_out/shard_emulate4.err; echo rc=$?
cat gpurun_out/shard_emulate4.txt; tail -3 gpurun_out/shard_emulate4.err
for P in 1 0; do GB_PREFETCH=$P AB_CONFIGS="auto" timeout 600 python scripts/ab_sweep.py sub_mapping_gpu livox_stress 2>/dev/null | grep -v "^#" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('prefetch $P', d['workload'][:12].ljust(12), d['config'].ljust(14), str(d['M_pf_s']).rjust(8), str(d['us_per_launch']).rjust(9), d['frac'], d['items_grid'], d['same_as_first'])"; done
