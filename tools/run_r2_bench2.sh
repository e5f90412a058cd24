mkdir -p gpurun_out/r2
nvidia-smi -L | head -4
for w in c1 c3; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 2 --warmup 2 --workload $w > gpurun_out/r2/bench_n2_$w.json 2> gpurun_out/r2/bench_n2_$w.err; echo "$w rc=$?"; tail -3 gpurun_out/r2/bench_n2_$w.err | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/r2/bench_n2_$w.json').read().strip().splitlines()[-1])
print('$w N=2 img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms/step', round(d['ms_per_step'],1), 'shard_invariance', d.get('shard_invariance'), d['config']['parallelism'], 'launches', d['gpu_launches'])
PY
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 1 --warmup 0 --impl reference --cpu-batch 8 2>&1 | tail -1 | cut -c1-400
