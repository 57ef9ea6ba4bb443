mkdir -p gpurun_out
PQT_SCAN_PREFETCH=2 PQT_RANK_CTAS=5 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s10_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s10_pytest.log; tail -3 gpurun_out/s10_pytest.log | cut -c1-300
run() {
  tag=$1; shift
  env "$@" timeout 600 python bench.py --dbsize 100000000 --no-cpu-baseline --variants knn --steps 10 > gpurun_out/s10_bench_100m_$tag.log 2>&1
  python - <<P
import json
try:
    j=json.loads([l for l in open('gpurun_out/s10_bench_100m_$tag.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('100M $tag', round(j['value']), round(j['ms_per_step'],3), {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()}, 'recall', j['recall_at_1'], j['recall_at_100'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/s10_bench_100m_$tag.log').read()[-1500:])
P
}
run base PQT_RANK_CTAS=4 PQT_SCAN_PREFETCH=0
run ctas5 PQT_RANK_CTAS=5 PQT_SCAN_PREFETCH=0
run pf2 PQT_RANK_CTAS=4 PQT_SCAN_PREFETCH=2
run pf4 PQT_RANK_CTAS=4 PQT_SCAN_PREFETCH=4
run pf1 PQT_RANK_CTAS=4 PQT_SCAN_PREFETCH=1
