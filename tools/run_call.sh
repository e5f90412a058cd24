mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "mode_switches" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest.log
