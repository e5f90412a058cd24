mkdir -p gpurun_out
timeout 900 python tools/tc_probe.py > gpurun_out/probe.log 2>&1; echo "probe rc=$?"
grep -E "RESULT|TIMEOUT|rc=|probe:|watchdog: [1-9]" gpurun_out/probe.log
timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu --layers > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print('cg2',d['value'],d['roofline']['ms_per_pass'],d['roofline']['frac'])
a=json.load(open('gpurun_out/layers.json'))
for x in a: print('%4d %4d %4d s%d up%d t%d | %6.1f us %5.0f TF' % (x['cin'],x['N'],x['H_out'],x['stride'],x['up'],x['taps'],x['us'],x['tflops']))
"
