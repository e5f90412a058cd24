#!/bin/bash
# Developer tool: build A/B variants of libsrdiff_b200.so (different -D flags for conv_tc.cu / conv_chain.cu) into
# dgm_img_super_resolution_b200/variants/lib_<name>.so; select one at run time with SRD_LIB_PATH.
#   tools/build_variants.sh name1="-DX=1 -DY=2" name2="..."
set -e
cd "$(dirname "$0")/../dgm_img_super_resolution_b200/csrc"
make -j4 >/dev/null
mkdir -p ../variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  for tu in conv_tc conv_chain; do
    $NVCC $flags -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC --expt-relaxed-constexpr -c $tu.cu -o ../variants/${tu}_$name.o &
  done
  wait
  $NVCC $ARCH -shared -o ../variants/lib_$name.so api.o conv_direct.o elementwise.o image_ops.o variant_ops.o ../variants/conv_tc_$name.o ../variants/conv_chain_$name.o -cudart static
  echo "built variants/lib_$name.so ($flags)"
done
