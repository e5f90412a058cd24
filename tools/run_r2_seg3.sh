echo "K=1 min-epilogue variant"; SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_k1min.so timeout 300 python tools/chain_time.py 64 40 one 2>&1 | tail -5
