mkdir -p gpurun_out/r2
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_tc_kernel -s 1 -c 1 -o gpurun_out/r2/ncu_plain_l1 -f python tools/profile_chain.py 1 plain > gpurun_out/r2/ncu_plain_l1.log 2>&1; echo rc=$?
