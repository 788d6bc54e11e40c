#!/bin/bash
# Experiment builds: scripts/build_variant.sh NAME "-DFLAG=.. ..."  ->  deep-image-prior_b200/libdip_NAME.so
# (select at run time with DIP_LIB=deep-image-prior_b200/libdip_NAME.so; *.so is git-ignored but travels with gpurun)
set -e
cd "$(dirname "$0")/../deep-image-prior_b200"
NAME=$1; EXTRA=$2
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $EXTRA"
mkdir -p build/$NAME
for f in deep downsample conv_simt engine; do   # deep.cu = conv_tc.cu + kernels_mem.cu in one translation unit
  $NVCC $FLAGS -c csrc/$f.cu -o build/$NAME/$f.o &
done
wait
$NVCC -shared -o libdip_$NAME.so build/$NAME/deep.o build/$NAME/downsample.o build/$NAME/conv_simt.o build/$NAME/engine.o -cudart static
echo "built libdip_$NAME.so"
