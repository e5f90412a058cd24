export SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_dbg2s.so
for spec in "64 64 160 64 conv3" ; do
echo "=== $spec (stamp 2 before the buffer wait: 'wait_buf' = geometry/setup only, 'chunks' includes the wait)"; timeout 300 python tools/timeline_epi.py $spec 2>&1 | grep -E "kernel|mean|spacing" | head -4
done
