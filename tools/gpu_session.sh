mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/s8_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s8_pytest.log; tail -4 gpurun_out/s8_pytest.log | cut -c1-300
for m in cta warp; do
PQT_INBOX_MODE=$m timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/s8_bench_n2_1b_$m.log 2>&1
python - <<P
import json
try:
    j=json.loads([l for l in open('gpurun_out/s8_bench_n2_1b_$m.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('N2 1B inbox=$m', j['value'], j['ms_per_step'], j['roofline']['stage_ms_per_step'], 'recall', j['recall_at_1'], j['recall_at_100'], 'e2e', j['e2e']['value'], j['clocks'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/s8_bench_n2_1b_$m.log').read()[-1500:])
P
done
