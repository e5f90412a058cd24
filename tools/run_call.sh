mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"conv_tc|p_sample|nchw_to_act|act_to_nchw|upsample|randn|time_table|conv_direct" --launch-skip 260 -c 330 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launch rc=$?"
tail -2 gpurun_out/ncu_launch.log | cut -c1-200
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof6_c64 -f python tools/profile_conv.py 64 64 160 64 conv3 3 > gpurun_out/ncu6a.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof6_c128 -f python tools/profile_conv.py 128 128 80 64 conv3 3 > gpurun_out/ncu6b.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --set full --clock-control none -k regex:p_sample -s 20 -c 1 -o gpurun_out/prof6_psample -f python bench.py --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu6c.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r01.csv
