mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --layers > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print('bench',d['value'],d['roofline']['ms_per_pass'],d['roofline']['frac'],d['e2e']['value'])
"
export SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_dbg.so
timeout 120 python tools/timeline_conv.py 64 64 160 64 res > gpurun_out/tl_b.log 2>&1
timeout 120 python tools/timeline_conv.py 128 128 80 64 res > gpurun_out/tl_d.log 2>&1
grep -h "kernel\|steady\|mean mma" gpurun_out/tl_b.log | head -3; grep -h "kernel\|steady\|mean mma" gpurun_out/tl_d.log | head -3
