timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('p_sample us', round(d['roofline_p_sample']['us_per_launch'],2), 'img/s', round(d['value'],1))"
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_next_rows.py -x -q 2>&1 | tail -2
