mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench.log 2>&1; echo "bench rc=$?"
python -c "
import json;d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1]);print('bench',d['value'],d['roofline']['ms_per_pass'],d['roofline']['frac'],d['e2e']['value'],d['roofline_p_sample'])
"
