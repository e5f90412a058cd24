mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_final.log 2>gpurun_out/bench_final.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_final.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_final.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_final.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_tc|p_sample|nchw_to_act|act_to_nchw|upsample|randn|time_table|conv_direct" --launch-skip 260 -c 330 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launch rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof7_c64 -f python tools/profile_conv.py 64 64 160 64 conv3 3 > gpurun_out/ncu7a.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof7_c128 -f python tools/profile_conv.py 128 128 80 64 conv3 3 > gpurun_out/ncu7b.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof7_c3 -f python tools/profile_conv.py 3 64 160 64 conv3 3 > gpurun_out/ncu7c.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:p_sample -s 20 -c 1 -o gpurun_out/prof7_psample -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu7d.log 2>&1; echo "ncu rc=$?"
timeout 300 python tools/accuracy_report.py bf16 fp32 2>&1 | grep -v Warning > gpurun_out/accuracy.log; cat gpurun_out/accuracy.log
