mkdir -p gpurun_out
M=sm__cycles_active.avg,sm__cycles_active.min,sm__cycles_active.max,sm__cycles_elapsed.max,gpu__time_duration.sum,smsp__warps_active.avg.per_cycle_active,smsp__cycles_active.avg,smsp__cycles_active.min,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,smsp__inst_executed.sum
for W in "8 1" "1 0"; do
timeout 600 ncu --metrics $M --clock-control none --cache-control none -k regex:k_vgicp_sweep3 -s 4 -c 2 --csv --log-file gpurun_out/ncu_shard_$(echo $W | tr ' ' _).csv python scripts/shard_emulate.py --one $W > gpurun_out/ncu_shard.log 2>&1; echo rc=$?
python - <<PY
import csv
rows=list(csv.reader(l for l in open('gpurun_out/ncu_shard_$(echo $W | tr ' ' _).csv') if l.startswith('"')))
h=rows[0]; 
for r in rows[1:]:
    d=dict(zip(h,r)); print(d['ID'], d['Metric Name'], d['Metric Value'])
PY
done
