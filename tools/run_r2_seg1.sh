mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_chain_gpu.py -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --layers > gpurun_out/r2/bench_seg.json 2> gpurun_out/r2/bench_seg.err; echo "bench rc=$?"; tail -2 gpurun_out/r2/bench_seg.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_seg.json').read().strip().splitlines()[-1])
print('FUSED  img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms_pass', round(d['roofline']['ms_per_pass'],3), 'frac', round(d['roofline']['frac'],4), d['clocks'], d['roofline']['kernel'])
PY
SRD_FUSE=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2/bench_nofuse.json 2> gpurun_out/r2/bench_nofuse.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_nofuse.json').read().strip().splitlines()[-1])
print('LAYERS img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms_pass', round(d['roofline']['ms_per_pass'],3), 'frac', round(d['roofline']['frac'],4), d['clocks'])
PY
