mkdir -p gpurun_out/r2s
for w in c1 c3; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --workload $w > gpurun_out/r2s/bench_scale_n2_$w.json 2> gpurun_out/r2s/bench_scale_n2_$w.err; echo "$w rc=$?"
tail -1 gpurun_out/r2s/bench_scale_n2_$w.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],d['config']['name'],'img/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],1),'shard_inv',d.get('shard_invariance'), d['e2e'])" || tail -5 gpurun_out/r2s/bench_scale_n2_$w.err
done
