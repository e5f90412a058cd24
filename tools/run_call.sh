mkdir -p gpurun_out/final
O=gpurun_out/final
for v in 0 1; do
SRD_NO_REVERSE=$v timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"conv_tc|p_sample|nchw_to_act|upsample2x" --launch-skip 260 -c 330 --csv --log-file $O/ncu_launch_list_warm_norev$v.csv python bench.py --steps 1 --warmup 1 --no-cpu > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $O/ncu_launch_list_warm_norev$v.csv --pass-json $O/unet_pass_traffic_warm_norev$v.json
done
