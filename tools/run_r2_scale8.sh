mkdir -p gpurun_out/r2s
nvidia-smi -L | wc -l
for spec in "8 c1" "4 c1" "8 c3"; do
set -- $spec
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $1 --steps 3 --warmup 3 --workload $2 > gpurun_out/r2s/bench_scale_n$1_$2.json 2> gpurun_out/r2s/bench_scale_n$1_$2.err; echo "N=$1 $2 rc=$?"
tail -1 gpurun_out/r2s/bench_scale_n$1_$2.json | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],d['config']['name'],'img/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],1),'shard_inv',d.get('shard_invariance'), d['clocks']['sm_mhz'])" || tail -5 gpurun_out/r2s/bench_scale_n$1_$2.err
done
