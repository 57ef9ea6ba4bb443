mkdir -p gpurun_out
REPRO_N=20000000 REPRO_QN=2500 timeout 900 compute-sanitizer --tool synccheck --print-limit 3 python tools/repro_k100.py > gpurun_out/k1_synccheck.log 2>&1; grep -v "^$" gpurun_out/k1_synccheck.log | grep -v "Host Frame" | tail -12 | cut -c1-200
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/k1_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/k1_pytest.log; tail -3 gpurun_out/k1_pytest.log
( time timeout 1700 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/k1_ref_1b.log 2>&1 ) 2>&1 | grep real; tail -c 900 gpurun_out/k1_ref_1b.log; free -g | head -2; df -h /dev/shm | tail -1
( time timeout 1700 python bench.py --steps 20 --warmup 5 > gpurun_out/k1_bench_1b.log 2>&1 ) 2>&1 | grep real; tail -c 400 gpurun_out/k1_bench_1b.log; python - <<'P'
import json
try:
    j=json.loads(open('gpurun_out/k1_bench_1b.log').read().strip().splitlines()[-1])
    print('1B', j['value'], j['roofline']['frac'], j['roofline']['stage_ms_per_step'], 'e2e', j['e2e']['value'], 'k100', j['e2e_k100']['value'], 'parity', j['parity_vs_oracle_on_cpu_sample'], 'cpu', j['cpu_baseline'], 'big', j['big_variant']['value'], j['big_variant']['roofline']['stage_ms_per_step'], j['big_variant']['parity_vs_oracle_on_cpu_sample'], 'cpuv', j['cpu_version_baseline'], j['setup'])
except Exception as e: print('1B parse failed', e)
P
