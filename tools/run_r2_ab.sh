for i in 1 2 3; do
for v in "SRD_S2_HALO=0" "SRD_S2_HALO=1" "SRD_S2_HALO=1 SRD_S2_MT1=1"; do
  env $v timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-cudnn --layers 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v img/s', round(d['value'],2), 'ms_pass', round(d['roofline']['ms_per_pass'],3), d['clocks']['sm_mhz'])"
  python -c "
import json
L=json.load(open('gpurun_out/layers.json')); print('   stride-2 layers 4/9/15 us:', [round(L[i]['us'],1) for i in (4,9,15)], 'sum all', round(sum(l['us'] for l in L),1))"
done; done
