mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for n in 256 128; do
SRD_FUSE_RES_MAX_N=$n timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-cudnn 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('res_max_n=$n img/s', round(d['value'],1), 'ms_pass', round(d['roofline']['ms_per_pass'],3), 'frac', round(d['roofline']['frac'],4), d['roofline']['kernel'], d['clocks']['sm_mhz'])"
done
