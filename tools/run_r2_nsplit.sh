timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for n in 1 0 1 0; do
SRD_NSPLIT=$n timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-cudnn --layers 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('nsplit=$n img/s', round(d['value'],1), 'ms_pass', round(d['roofline']['ms_per_pass'],3), 'frac', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])"
python - <<PY
import json
L=json.load(open('gpurun_out/layers.json'))
print('   layers 16-24 us:', [round(l['us'],1) for l in L[16:25]], 'sum', round(sum(l['us'] for l in L[16:25]),1))
PY
done
