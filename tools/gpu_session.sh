mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "eight_parts or traversal_tables" > gpurun_out/s18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s18_pytest.log; tail -25 gpurun_out/s18_pytest.log | cut -c1-300
