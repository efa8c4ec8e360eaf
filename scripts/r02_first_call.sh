# First GPU call of round 2: things written after round 1's GPU budget was spent and therefore never executed.
#   gpurun --timeout 1200 -- 'bash scripts/r02_first_call.sh'
mkdir -p gpurun_out
# 1. the whole GPU suite; -rxX lists xfail / XPASS (k_deskew's first execution is tests/test_deskew.py::test_gpu_deskew_matches_oracle)
timeout 600 python -m pytest tests -m gpu -q -rxX --timeout 200 2>&1 | tail -15 | tee gpurun_out/r02_pytest.txt
# 2. smoke + default bench line (what the driver runs)
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/r02_bench.log 2>&1; tail -c 1500 gpurun_out/r02_bench.log
# 3. frame-setup and preprocess timings with the N-scaled k-NN ring budget (not re-measured in round 1)
timeout 300 python scripts/bench_setup.py > gpurun_out/r02_bench_setup.json 2>/dev/null; grep -E "find_neighbors|points\"" gpurun_out/r02_bench_setup.json
