for i in 1 2 3; do
for v in default late; do
  if [ $v = default ]; then unset SRD_LIB_PATH; else export SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-cudnn 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v img/s', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms_pass', round(d['roofline']['ms_per_pass'],3), d['clocks']['sm_mhz'])"
done; done
