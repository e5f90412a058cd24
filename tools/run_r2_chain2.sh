export SRD_LIB_PATH=$PWD/dgm_img_super_resolution_b200/variants/lib_wd2.so CUDA_LAUNCH_BLOCKING=1
timeout 200 python tools/chain_debug.py 2 8 8 0:2 0:5 0:12 0:42 2>&1 | grep -v "^[0-9]* cin" | cut -c1-300
