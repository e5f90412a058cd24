mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 900 python tools/tc_probe.py > gpurun_out/probe.log 2>&1; grep -E "FAIL|TIMEOUT|probe:" gpurun_out/probe.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --layers > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print('bench',d['value'],d['roofline']['ms_per_pass'],d['roofline']['frac'],d['e2e']['value'])
a=json.load(open('gpurun_out/layers.json'))
print(sum(x['us'] for x in a))
for x in a:
  if x['stride']==2 or x['up']==2: print('%4d %4d %4d s%d up%d t%d | %6.1f us %5.0f TF' % (x['cin'],x['N'],x['H_out'],x['stride'],x['up'],x['taps'],x['us'],x['tflops']))
"
