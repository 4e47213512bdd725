#!/bin/bash
# Build experimental variants of libl3d_b200.so that differ only in knn.cu compile-time knobs.
# usage: profiles/build_variants.sh "tag:-DFLAG=.. -DFLAG2=.." ...
set -e
cd "$(dirname "$0")/../learning3d_b200/csrc"
make -j8 >/dev/null
OTHERS=$(ls build/*.o | grep -v "build/knn.o" | grep -v "build/var_")
for spec in "$@"; do
  tag="${spec%%:*}"; flags="${spec#*:}"
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false \
     -Xcompiler -fPIC,-O2 -Xptxas -v $flags -c knn.cu -o build/var_knn_$tag.o 2> build/var_knn_$tag.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../profiles/variants/lib_$tag.so build/var_knn_$tag.o $OTHERS -lcudart
  echo "$tag: $(grep -A2 'knn_kernelILi0ELi1ELb1ELb1' build/var_knn_$tag.log | grep -E 'Used' | head -1)"
done
