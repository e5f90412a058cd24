mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_final.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --impl reference > $O/bench_reference_arm_final.json 2> $O/bench_reference_arm.err; echo "ref rc=$?"; tail -c 600 $O/bench_reference_arm_final.json
timeout 900 python bench.py > $O/bench_final.json 2> $O/bench.err; echo "bench rc=$?"; tail -c 1500 $O/bench_final.json
timeout 600 python tools/accuracy_report.py > $O/accuracy_report.txt 2>&1; tail -12 $O/accuracy_report.txt
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_tc|p_sample|nchw_to_act|upsample2x" --launch-skip 260 -c 330 --csv --log-file $O/ncu_launch_list_bench.csv python bench.py --steps 1 --warmup 1 --no-cpu > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $O/ncu_launch_list_bench.csv > $O/ncu_launch_list_summary.txt 2>&1; cat $O/ncu_launch_list_summary.txt
python tools/ncu_summary.py $O/ncu_launch_list_bench.csv --pass-json $O/unet_pass_traffic.json
