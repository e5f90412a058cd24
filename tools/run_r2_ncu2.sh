mkdir -p gpurun_out/r2
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_seg -s 1 -c 1 -o gpurun_out/r2/ncu_fused_l2 -f python tools/profile_chain.py 2 fused > gpurun_out/r2/ncu_fused_l2.log 2>&1; echo rc=$?
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/r2/ncu_plain_l2 -f python tools/profile_chain.py 2 plain > gpurun_out/r2/ncu_plain_l2.log 2>&1; echo rc=$?
cd gpurun_out/r2
for f in ncu_plain_l2 ncu_fused_l2; do ncu -i $f.ncu-rep --page source --csv > $f.src.csv 2>/dev/null; ncu -i $f.ncu-rep --page raw --csv > $f.raw.csv 2>/dev/null; rm -f $f.ncu-rep; done
