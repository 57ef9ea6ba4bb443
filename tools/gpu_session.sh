mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s15_pytest.log; tail -3 gpurun_out/s15_pytest.log | cut -c1-300
run() {
  tag=$1; shift
  env "$@" timeout 900 python bench.py --no-cpu-baseline --variants knn --steps 20 > gpurun_out/s15_bench_1b_$tag.log 2>&1
  python - <<P
import json
try:
    j=json.loads([l for l in open('gpurun_out/s15_bench_1b_$tag.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('1B $tag', round(j['value']), round(j['ms_per_step'],3), {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()}, 'frac', round(j['roofline']['frac'],3), 'recall', j['recall_at_1'], j['recall_at_100'], 'e2e', round(j['e2e']['value']), 'k100', round(j['e2e_k100']['value']))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/s15_bench_1b_$tag.log').read()[-1500:])
P
}
run dyn X=1
