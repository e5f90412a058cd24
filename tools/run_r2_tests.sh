mkdir -p gpurun_out/r2
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2/pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
