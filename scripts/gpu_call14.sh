mkdir -p gpurun_out
GB_LAZY=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
AB_CONFIGS="auto=,lazy=GB_LAZY:1,lazy1024=GB_LAZY:1;GB_TILE:1024,lazyipw12=GB_LAZY:1;GB_ITEMS_PER_WARP:12,lazywide=GB_LAZY:1;GB_TAPER_A:300;GB_TAPER_B:100" timeout 900 python scripts/shard_emulate.py 1 8 > gpurun_out/shard_emulate2.txt 2> gpurun_out/shard_emulate2.err; echo rc=$?
cat gpurun_out/shard_emulate2.txt; tail -3 gpurun_out/shard_emulate2.err
