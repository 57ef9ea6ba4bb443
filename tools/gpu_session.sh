mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log; tail -4 gpurun_out/s2_pytest.log | cut -c1-300
for m in bitonic bucket; do
PQT_RANK_SORT=$m timeout 600 python tools/phase_probe.py --dbsize 100000000 > gpurun_out/s2_phase_100m_$m.log 2>&1; echo "== phase $m"; head -12 gpurun_out/s2_phase_100m_$m.log
PQT_RANK_SORT=$m timeout 600 python bench.py --dbsize 100000000 --no-cpu-baseline --variants knn --steps 10 > gpurun_out/s2_bench_100m_$m.log 2>&1
python - <<P
import json
try:
    j=json.loads([l for l in open('gpurun_out/s2_bench_100m_$m.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('100M $m', j['value'], j['ms_per_step'], j['roofline']['stage_ms_per_step'], 'recall', j['recall_at_1'], j['recall_at_100'], 'e2e', j['e2e']['value'], 'exact', j['exact_rank_queries_per_step'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/s2_bench_100m_$m.log').read()[-1500:])
P
done
timeout 1200 python bench.py --no-cpu-version > gpurun_out/s2_bench_1b.log 2>&1; tail -c 3500 gpurun_out/s2_bench_1b.log
