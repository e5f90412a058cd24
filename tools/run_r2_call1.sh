mkdir -p gpurun_out/r2
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
timeout 600 python bench.py --steps 5 --warmup 3 --layers --no-cpu > gpurun_out/r2/bench_base.json 2> gpurun_out/r2/bench_base.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2/bench_base.json
cp gpurun_out/layers.json gpurun_out/r2/layers_base.json
SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_dbg3.so timeout 600 python tools/timeline_pass.py 64 40 40 > gpurun_out/r2/timeline_pass.txt 2>&1; echo "timeline rc=$?"
cp gpurun_out/timeline_pass.json gpurun_out/r2/
cat gpurun_out/r2/timeline_pass.txt
