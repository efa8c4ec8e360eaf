mkdir -p gpurun_out
AB_CONFIGS="auto=,notaper=GB_TAPER:0,ipw12=GB_ITEMS_PER_WARP:12,taperwide=GB_TAPER_A:300;GB_TAPER_B:100,tile1024=GB_TILE:1024" timeout 900 python scripts/shard_emulate.py 1 8 > gpurun_out/shard_emulate.txt 2> gpurun_out/shard_emulate.err; echo rc=$?
cat gpurun_out/shard_emulate.txt; tail -3 gpurun_out/shard_emulate.err
