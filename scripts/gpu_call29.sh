# last call of round 2 (1.2 GPU-minutes left): the only GPU-facing things this session touched are the C++ shim headers
# (Pose conversions, CloudDeskewing) -- run the C++ shim tests and smoke(); kernels are SASS-identical to the verified build
mkdir -p gpurun_out
timeout 55 python -m pytest tests/test_cpp_shim.py -m gpu -x -q > gpurun_out/pytest_cpp_shim_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_cpp_shim_final.log
timeout 40 python __graft_entry__.py smoke 2>&1 | tail -1
