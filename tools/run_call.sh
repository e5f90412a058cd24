set -x
mkdir -p gpurun_out
timeout 120 tools/bin/umma_bw > gpurun_out/umma_bw.log 2>&1; echo "umma rc=$?"
for v in e8lo e8hi e16lo e16hi; do
  SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_$v.so timeout 300 python bench.py --steps 2 --warmup 2 --no-cpu --layers > gpurun_out/bench_$v.log 2>&1; echo "$v rc=$?"
  cp gpurun_out/layers.json gpurun_out/layers_$v.json
  python -c "
import json;d=json.loads(open('gpurun_out/bench_$v.log').read().strip().splitlines()[-1]);print('$v',d['value'],d['roofline']['ms_per_pass'])"
done
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
cat gpurun_out/umma_bw.log
