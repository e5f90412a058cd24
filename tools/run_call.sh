mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_configs_gpu.py -m gpu -x -q -k "full_bench" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
