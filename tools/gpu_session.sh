mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/s16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s16_pytest.log; tail -3 gpurun_out/s16_pytest.log | cut -c1-300
timeout 900 python bench.py --no-cpu-baseline --steps 20 > gpurun_out/s16_bench_1b.log 2>&1
python - <<P
import json
j=json.loads([l for l in open('gpurun_out/s16_bench_1b.log').read().strip().splitlines() if l.startswith('{')][-1])
b=j['big_variant']
print('1B', round(j['value']), round(j['ms_per_step'],3), {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()}, 'frac', round(j['roofline']['frac'],3), 'recall', j['recall_at_1'], j['recall_at_100'], 'e2e', round(j['e2e']['value']), 'k100', round(j['e2e_k100']['value']))
print('BIG', round(b['value']), round(b['ms_per_step'],3), {k:round(v,3) for k,v in b['roofline']['stage_ms_per_step'].items()}, 'recall', b['recall_at_1'], b['recall_at_100'])
P
