mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_host_tools.py -m gpu -x -q -k "compact or createdb" > gpurun_out/s19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s19_pytest.log; tail -30 gpurun_out/s19_pytest.log | cut -c1-300
