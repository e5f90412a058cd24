mkdir -p gpurun_out/r2
( time timeout 900 python bench.py ) > gpurun_out/r2/bench_c1.json 2> gpurun_out/r2/bench_c1.err; echo "c1 rc=$?"; tail -4 gpurun_out/r2/bench_c1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2/bench_c1.json').read().strip().splitlines()[-1])
print('c1 img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), 'ps', d['roofline_p_sample']['frac'], d['clocks'])
print('cpu', d['cpu_baseline'])
print('cudnn', d['cudnn_baseline'])
PY
for w in c2_whole c2_tiled c4_T25; do
timeout 900 python bench.py --workload $w --steps 2 --warmup 1 --no-cpu --no-cudnn > gpurun_out/r2/bench_$w.json 2> gpurun_out/r2/bench_$w.err; echo "$w rc=$?"; tail -2 gpurun_out/r2/bench_$w.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2/bench_$w.json').read().strip().splitlines()[-1])
print('$w img/s', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), 'ms/step', round(d['ms_per_step'],1), 'frac', round(d['roofline']['frac'],4))
PY
done
