# one batched GPU call: A/B of the first-item policy, the five workloads with CPU baselines, the reference arm, ncu captures
mkdir -p gpurun_out
for sf in 1 0; do
  for w in odometry_gpu single_pair; do
    GB_STATIC_FIRST=$sf timeout 200 python bench.py --workload $w --no-cpu-baseline --steps 60 > gpurun_out/ab.log 2>&1
    python -c "
import json
d=json.loads([l for l in open('gpurun_out/ab.log').read().splitlines() if l.startswith('{')][-1]); print('AB static_first=$sf $w', round(d['value']), round(d['roofline']['launch_ms']*1000,1),'us')" || tail -3 gpurun_out/ab.log
  done
done | tee gpurun_out/ab_static_first.txt
for w in global_mapping_gpu sub_mapping_gpu livox_stress odometry_gpu single_pair; do
  timeout 400 python bench.py --workload $w --steps 100 > gpurun_out/bench_$w.log 2>&1; echo "$w rc=$?"
done
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
GB_PROFILE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1; echo "ncu1 rc=$?"
GB_PROFILE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_vgicp_sweep -c 1 -f -o gpurun_out/prof_r01_final python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
GB_PROFILE=1 timeout 400 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_vgicp_sweep -c 1 -f -o gpurun_out/prof_r01_final_odometry python bench.py --workload odometry_gpu --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_odo.log 2>&1; echo "ncu3 rc=$?"
timeout 200 python scripts/bench_setup.py > gpurun_out/bench_setup.json 2> gpurun_out/bench_setup.err; echo "setup rc=$?"
