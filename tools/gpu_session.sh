mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s6_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s6_pytest.log; tail -4 gpurun_out/s6_pytest.log | cut -c1-300
timeout 600 python tools/phase_probe.py --dbsize 100000000 > gpurun_out/s6_phase_100m.log 2>&1; head -16 gpurun_out/s6_phase_100m.log
timeout 1200 python bench.py --no-cpu-version > gpurun_out/s6_bench_1b.log 2>&1; tail -c 3500 gpurun_out/s6_bench_1b.log | cut -c1-2200
