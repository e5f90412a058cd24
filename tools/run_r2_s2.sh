O=gpurun_out/s2; mkdir -p $O
for spec in "64 64 160 64 conv3s2" "128 128 80 64 conv3s2"; do
  name=$(echo $spec | tr ' ' '_')
  NORES=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:conv_tc -s 2 -c 1 -o $O/rep_$name -f python tools/profile_conv.py $spec 3 > $O/ncu_$name.log 2>&1
  python tools/ncu_summary.py $O/rep_$name.ncu-rep > $O/summary_$name.txt 2>&1
  ncu -i $O/rep_$name.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > $O/srcl_$name.csv.gz
  tail -1 $O/ncu_$name.log
done
rm -f $O/*.ncu-rep
