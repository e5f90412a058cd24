# round-2 evidence run: GPU tests, smoke, bench lines, ncu launch lists and full captures -> gpurun_out/r2p (summaries only)
O=gpurun_out/r2p; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest_gpu_final.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 --layers > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"
cp gpurun_out/layers.json $O/layers_final.json
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_arm_final.json 2> $O/bench_ref.err; echo "ref rc=$?"
# driver-style launch list of smoke(): the first 1000 launches must contain our kernels
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file $O/ncu_launch_list_smoke.csv python -c "import __graft_entry__ as g; g.smoke()" > $O/ncu_smoke.log 2>&1; echo "ncu smoke rc=$?"
python tools/ncu_summary.py $O/ncu_launch_list_smoke.csv > $O/ncu_launch_list_smoke_summary.txt 2>&1; head -12 $O/ncu_launch_list_smoke_summary.txt
# launch list of the bench command, L2 not flushed between launches, with DRAM bytes
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"conv_tc|p_sample" --launch-skip 260 -c 330 --csv --log-file $O/ncu_launch_list_bench_warmL2.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-cudnn > $O/ncu_bench.log 2>&1; echo "ncu bench rc=$?"
python tools/ncu_summary.py $O/ncu_launch_list_bench_warmL2.csv > $O/ncu_launch_list_warmL2_summary.txt 2>&1
python tools/ncu_summary.py $O/ncu_launch_list_bench_warmL2.csv --pass-json $O/unet_pass_traffic.json
# full captures of the two roofline kernels (reports stay on the box; summaries + source hot spots come back)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 2 -c 1 -o $O/full_conv64 -f python tools/profile_conv.py 64 64 160 64 conv3 3 > $O/ncu_full_conv64.log 2>&1
python tools/ncu_summary.py $O/full_conv64.ncu-rep > $O/ncu_full_conv64x64_160_B64_summary.txt 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 2 -c 1 -o $O/full_conv256 -f python tools/profile_conv.py 256 256 20 64 conv3 3 > $O/ncu_full_conv256.log 2>&1
python tools/ncu_summary.py $O/full_conv256.ncu-rep > $O/ncu_full_conv256x256_20_B64_summary.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:p_sample -s 20 -c 1 -o $O/full_ps -f python bench.py --steps 1 --warmup 1 --no-cpu --no-cudnn > $O/ncu_full_ps.log 2>&1
python tools/ncu_summary.py $O/full_ps.ncu-rep > $O/ncu_full_p_sample_B64_summary.txt 2>&1
rm -f $O/*.ncu-rep
tail -c 600 $O/bench_final.json; echo; tail -c 500 $O/bench_reference_arm_final.json; echo
cat $O/ncu_launch_list_warmL2_summary.txt | head; tail -1 $O/unet_pass_traffic.json | head -c 300
