for v in p0 p15 p8 p7; do
echo "== variant $v"; SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_$v.so timeout 300 python tools/chain_time.py 64 40 one 2 6 18 37 2>&1 | tail -4
done
