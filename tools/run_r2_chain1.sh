mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_chain_gpu.py -x -q > gpurun_out/r2/chain1.log 2>&1; echo "chain pytest rc=$?"; tail -30 gpurun_out/r2/chain1.log
