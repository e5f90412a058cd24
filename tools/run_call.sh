mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log

for v in 1 0; do
SRD_NO_FOLD_COND=$v timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_f$v.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench_f$v.log').read().strip().splitlines()[-1]);
print('no_fold=$v: img/s',round(d['value'],1),'pass ms',round(d['roofline']['ms_per_pass'],3),'frac',round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
" || tail -3 gpurun_out/bench_f$v.log
done
