mkdir -p gpurun_out
for v in 0 24 16 8 7 0 24; do
SRD_L2_HINT_MASK=$v timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_m$v.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench_m$v.log').read().strip().splitlines()[-1]);
print('mask=$v: img/s',round(d['value'],1),'pass ms',round(d['roofline']['ms_per_pass'],3),'frac',round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
" || tail -3 gpurun_out/bench_m$v.log
done
