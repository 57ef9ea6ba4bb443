# final verification of a round: the driver's own sequence (GPU tests, smoke, both bench arms)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.log; tail -3 gpurun_out/g_pytest.log | cut -c1-300
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/g_smoke.log 2>&1; tail -2 gpurun_out/g_smoke.log | cut -c1-300
timeout 900 python bench.py --impl reference > gpurun_out/g_bench_ref.log 2>&1; tail -c 400 gpurun_out/g_bench_ref.log
timeout 1500 python bench.py > gpurun_out/g_bench_1b.log 2>&1
python - <<P
import json
j=json.loads([l for l in open('gpurun_out/g_bench_1b.log').read().strip().splitlines() if l.startswith('{')][-1])
b=j['big_variant']
print('1B', round(j['value']), round(j['ms_per_step'],3), {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()}, 'frac', round(j['roofline']['frac'],3), 'recall', j['recall_at_1'], j['recall_at_100'], 'parity', j['parity_vs_oracle_on_cpu_sample'], 'e2e', round(j['e2e']['value']), round(j['e2e']['ms_per_step'],3), 'k100', round(j['e2e_k100']['value']), 'cpu', j['cpu_baseline']['value'], 'clocks', j['clocks'])
print('BIG', round(b['value']), round(b['ms_per_step'],3), b['roofline']['stage_ms_per_step'], 'parity', b['parity_vs_oracle_on_cpu_sample'], 'e2e', round(b['e2e']['value']), 'recall', b['recall_at_1'], b['recall_at_100'])
P
