mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_chain_gpu.py -x -q > gpurun_out/r2/chain_tests.log 2>&1; echo "chain pytest rc=$?"; tail -5 gpurun_out/r2/chain_tests.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2/bench_chain.json 2> gpurun_out/r2/bench_chain.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_chain.json').read().strip().splitlines()[-1])
print('CHAIN  img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms_pass', round(d['roofline']['ms_per_pass'],3), 'frac', round(d['roofline']['frac'],4), d['clocks'])
PY
SRD_CHAIN=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2/bench_nochain.json 2> gpurun_out/r2/bench_nochain.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_nochain.json').read().strip().splitlines()[-1])
print('LAYERS img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'ms_pass', round(d['roofline']['ms_per_pass'],3), 'frac', round(d['roofline']['frac'],4), d['clocks'])
PY
tail -3 gpurun_out/r2/bench_chain.err
