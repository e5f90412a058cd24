mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_configs_gpu.py -m gpu -x -q --durations=8 > gpurun_out/pytest_cfg.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_cfg.log
