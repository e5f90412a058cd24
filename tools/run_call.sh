mkdir -p gpurun_out
timeout 600 python tools/accuracy_report.py bf16 fp32 > gpurun_out/accuracy_report.txt 2>&1; cat gpurun_out/accuracy_report.txt | tail -12
