mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 8 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/bench_n$n.log 2>&1; echo "rc=$?"
tail -1 gpurun_out/bench_n$n.log | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('N',d['n_gpus'],'img/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'ms/step',round(d['ms_per_step'],1), d['clocks'])" || tail -5 gpurun_out/bench_n$n.log
done
