mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
for v in 1 65 1; do
SRD_TC_A_MODE=$v timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_$v.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench_$v.log').read().strip().splitlines()[-1]);
print('a_mode $v: img/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'pass ms',round(d['roofline']['ms_per_pass'],3),'frac',round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
" || tail -5 gpurun_out/bench_$v.log
done
