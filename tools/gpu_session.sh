mkdir -p gpurun_out
for n in 2 4; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/s17_bench_n${n}_1b.log 2>&1
python - <<P
import json
try:
    j=json.loads([l for l in open('gpurun_out/s17_bench_n${n}_1b.log').read().strip().splitlines() if l.startswith('{')][-1])
    print('N$n 1B', round(j['value']), round(j['ms_per_step'],3), {k:round(v,3) for k,v in j['roofline']['stage_ms_per_step'].items()}, 'recall', j['recall_at_1'], j['recall_at_100'], 'e2e', round(j['e2e']['value']), j['limiter']['largest_stage'], j['setup']['index_build_s'])
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/s17_bench_n${n}_1b.log').read()[-1500:])
P
done
