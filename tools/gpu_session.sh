mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/d2h_probe.py > gpurun_out/s9_d2h_probe_n8.log 2>&1; tail -1 gpurun_out/s9_d2h_probe_n8.log | cut -c1-600
for m in warp cta; do
PQT_INBOX_MODE=$m timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s9_bench_n8_1b_$m.log 2>&1
python - <<P
import json
try:
    j=json.loads([l for l in open('gpurun_out/s9_bench_n8_1b_$m.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('N8 1B inbox=$m', j['value'], j['ms_per_step'], j['roofline']['stage_ms_per_step'], 'recall', j['recall_at_1'], j['recall_at_100'], 'e2e', j['e2e']['value'], j['clocks'], j['setup'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/s9_bench_n8_1b_$m.log').read()[-1500:])
P
done
