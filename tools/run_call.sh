mkdir -p gpurun_out/final
O=gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -q > $O/pytest_gpu_final.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -k regex:"conv_tc|p_sample|nchw_to_act|upsample2x" --launch-skip 260 -c 330 --csv --log-file $O/ncu_launch_list_bench_warmL2.csv python bench.py --steps 1 --warmup 1 --no-cpu > $O/ncu_bench.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $O/ncu_launch_list_bench_warmL2.csv > $O/ncu_launch_list_summary.txt 2>&1; cat $O/ncu_launch_list_summary.txt
python tools/ncu_summary.py $O/ncu_launch_list_bench_warmL2.csv --pass-json profiles/r01/unet_pass_traffic.json
cp profiles/r01/unet_pass_traffic.json $O/
timeout 600 python tools/accuracy_report.py bf16 fp32 > $O/accuracy_report.txt 2>&1
timeout 900 python bench.py --impl reference > $O/bench_reference_arm_final.json 2> $O/bench_reference_arm.err; echo "ref rc=$?"
timeout 900 python bench.py > $O/bench_final.json 2> $O/bench.err; echo "bench rc=$?"; python -c "
import json;d=json.loads(open('$O/bench_final.json').read().strip().splitlines()[-1]);
print('img/s',round(d['value'],1),'e2e',round(d['e2e']['value'],1),'pass ms',round(d['roofline']['ms_per_pass'],3),'frac',round(d['roofline']['frac'],4),'traffic GB',round(d['roofline']['traffic']/1e9,3), d['clocks'], 'cpu', d['cpu_baseline']['value'])"
