mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_variants.py -m gpu -q > gpurun_out/pytest_variants.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_variants.log
