mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --layers > gpurun_out/bench.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);
a=json.load(open('gpurun_out/layers.json'))
print('img/s',round(d['value'],1),'pass ms',round(d['roofline']['ms_per_pass'],3),'frac',round(d['roofline']['frac'],4),'sum layers us',round(sum(x['us'] for x in a)))
x=a[-1]; print('final conv %6.1f us %5.0f TF' % (x['us'],x['tflops']))
"
