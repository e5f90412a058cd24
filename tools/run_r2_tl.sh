export SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_k3.so
SRD_NO_REVERSE=1 timeout 600 python tools/chain_time.py 64 40 fused 2>&1 | tail -11
