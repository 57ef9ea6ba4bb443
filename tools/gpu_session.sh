mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/o1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/o1_pytest.log; tail -12 gpurun_out/o1_pytest.log | cut -c1-300
timeout 900 python tools/ab_rerank.py --steps 5 > gpurun_out/o1_ab_1b.log 2>&1; tail -4 gpurun_out/o1_ab_1b.log
