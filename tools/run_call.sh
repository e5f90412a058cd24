mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
for m in 1 33; do
SRD_TC_A_MODE=$m timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --layers > gpurun_out/bench_m$m.log 2>&1; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench_m$m.log').read().strip().splitlines()[-1]);
a=json.load(open('gpurun_out/layers.json'))
print('mode $m: img/s',round(d['value'],1),'pass ms',round(d['roofline']['ms_per_pass'],3),'sum layers us',round(sum(x['us'] for x in a)))
for x in a:
  if x['H_out']==20 and x['taps']==9: print('   %4d %4d %4d s%d up%d t%d | %6.1f us %5.0f TF' % (x['cin'],x['N'],x['H_out'],x['stride'],x['up'],x['taps'],x['us'],x['tflops']))
"
done
