mkdir -p gpurun_out
python tools/profile_conv.py 3 64 160 64 conv3 5
NORES=1 python tools/profile_conv.py 64 64 160 64 conv3 5
python tools/profile_conv.py 64 64 160 64 conv3 5
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 2 -c 1 -o gpurun_out/prof5_c3 -f python tools/profile_conv.py 3 64 160 64 conv3 3 > gpurun_out/ncu5.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu5.log
