mkdir -p gpurun_out
for dd in 1 0; do
PQT_BINS_DEDUPE=$dd timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --dbsize 100000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_n2_100m_dd$dd.log 2>&1
python - <<P
import json
try:
    j=json.loads([l for l in open('gpurun_out/r1_bench_n2_100m_dd$dd.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('N2 100M dedupe $dd', j['value'], j['ms_per_step'], j['roofline']['stage_ms_per_step'], 'recall', j['recall_at_1'], j['recall_at_100'], 'e2e', j['e2e']['value'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/r1_bench_n2_100m_dd$dd.log').read()[-1500:])
P
done
