mkdir -p gpurun_out
export SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_di.so
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest(di) rc=$?"; tail -3 gpurun_out/pytest.log
for v in di default di default; do
if [ $v = default ]; then unset SRD_LIB_PATH; else export SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_$v.so; fi
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --layers > gpurun_out/bench_$v.log 2>&1
python -c "
import json;d=json.loads(open('gpurun_out/bench_$v.log').read().strip().splitlines()[-1]);
a=json.load(open('gpurun_out/layers.json'))
print('$v: img/s',round(d['value'],1),'pass ms',round(d['roofline']['ms_per_pass'],3),'frac',round(d['roofline']['frac'],4), 'N=64 layers:', [round(x['us'],1) for x in a if x['N']==64 and x['taps']==9 and x['stride']==1 and x['up']==1])
" || tail -3 gpurun_out/bench_$v.log
done
