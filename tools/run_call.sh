mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest.log
for v in prev default prev default; do
if [ $v = default ]; then unset SRD_LIB_PATH; else export SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_$v.so; fi
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_$v.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench_$v.log').read().strip().splitlines()[-1]);
print('$v: img/s',round(d['value'],1),'pass ms',round(d['roofline']['ms_per_pass'],3),'frac',round(d['roofline']['frac'],4), d['clocks']['sm_mhz'])
" || tail -3 gpurun_out/bench_$v.log
done
