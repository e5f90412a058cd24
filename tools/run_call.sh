mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python tools/accuracy_report.py bf16 2>&1 | grep -v Warning | tee gpurun_out/accuracy.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --layers > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print('bench',d['value'],d['roofline']['ms_per_pass'],d['roofline']['frac'],d['e2e']['value'])
a=json.load(open('gpurun_out/layers.json'))
for x in a[:5]+a[-6:]: print('%4d %4d %4d s%d up%d t%d | %6.1f us %5.0f TF' % (x['cin'],x['N'],x['H_out'],x['stride'],x['up'],x['taps'],x['us'],x['tflops']))
"
