mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
for v in 0 1 0 1; do
SRD_NO_FUSED_RES=$v timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_v$v.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench_v$v.log').read().strip().splitlines()[-1]);
print('no_fused_res=$v: img/s',round(d['value'],1),'pass ms',round(d['roofline']['ms_per_pass'],3),'frac',round(d['roofline']['frac'],4))
"
done
