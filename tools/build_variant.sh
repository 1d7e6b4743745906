#!/bin/bash
# Dev-time A/B builds: tools/build_variant.sh NAME FILE.cu "-DMACRO=.. -DMACRO2=.." -> variants/NAME.so (the library with that one object rebuilt)
# Run a tool against it with RENDERTOY_B200_LIB=$PWD/variants/NAME.so.
set -e
NAME=$1; SRC=$2; DEFS=$3
mkdir -p variants /tmp/variants_build
OBJ=/tmp/variants_build/$NAME.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -ccbin /usr/bin/g++ -Xcompiler -fPIC,-O2,-fno-fast-math -Xptxas -v $DEFS -c rendertoy_b200/csrc/$SRC -o $OBJ 2>&1 | grep -E "spill|Used" | sort | uniq -c | sort -rn | head -${4:-3}
BASE=$(basename $SRC .cu)
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -ccbin /usr/bin/g++ -o variants/$NAME.so $OBJ $(ls rendertoy_b200/build/*.o | grep -v "/$BASE.o") -cudart static -ldl
echo "variants/$NAME.so"
