mkdir -p gpurun_out
timeout 600 python tools/dual_stream_probe.py 64 2 2>&1 | tail -4
timeout 600 python tools/dual_stream_probe.py 64 4 2>&1 | tail -2
timeout 600 python tools/dual_stream_probe.py 128 2 2>&1 | tail -2
