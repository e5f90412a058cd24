mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_chain_gpu.py -x -q 2>&1 | tail -3
timeout 900 python tools/chain_time.py 64 40 > gpurun_out/r2/chain_time.txt 2>&1; echo rc=$?; grep -E "^ ?(0|1|2|6|12|18|30|35|37|40|41) |sum|ops" gpurun_out/r2/chain_time.txt
echo "---- forward order"
SRD_NO_REVERSE=1 timeout 600 python tools/chain_time.py 64 40 0:42 1:4 18:25 2>&1 | tail -3
