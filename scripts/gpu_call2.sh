# round 2, call 2: ncu captures of the v4 (bulk-async staged) and v3 kernels on sub-mapping and odometry
mkdir -p gpurun_out
for cfg in "4 sub_mapping_gpu" "3 sub_mapping_gpu" "4 odometry_gpu"; do
  set -- $cfg
  GB_KERNEL=$1 GB_PROFILE=1 timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_vgicp_sweep -c 1 -f -o gpurun_out/prof_r02a_v$1_$2 python bench.py --workload $2 --steps 1 --warmup 3 --no-cpu-baseline --verify 0 > gpurun_out/ncu_r02a_v$1_$2.log 2>&1; echo "ncu v$1 $2 rc=$?"
done
ls -la gpurun_out/*.ncu-rep | tail -5
