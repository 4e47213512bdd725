#!/bin/bash
# A/B builds of knn_tpr.cu (compile-time knobs) linked against the current objects of the library:
#   ab/libl3d_tpr_<name>.so, selected at run time with L3D_B200_LIB (profiles/time_knn_paths.py).
set -e
cd "$(dirname "$0")/../learning3d_b200/csrc"
make -j8 > /dev/null
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -Xcompiler -fPIC,-O3,-fopenmp"
OBJS=$(ls build/*.o | grep -v knn_tpr.o)
build() {  # name, defines...
  name=$1; shift
  nvcc $FLAGS "$@" -c knn_tpr.cu -o ../../ab/tpr_$name.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../ab/libl3d_tpr_$name.so $OBJS ../../ab/tpr_$name.o -lcudart -lgomp
  rm -f ../../ab/tpr_$name.o
}
for v in "$@"; do
  case $v in
    dsetp)  build dsetp -DL3D_TPR_DSETP=1 & ;;
    nofuse) build nofuse -DL3D_TPR_FUSE_THR=0 & ;;
    g1)     build g1 -DL3D_TPR_UNROLL_G=1 & ;;
    g4)     build g4 -DL3D_TPR_UNROLL_G=4 & ;;
    loop)   build loop -DL3D_TPR_EXTRACT=0 & ;;
    dg2)    build dg2 -DL3D_DUO_UNROLL_G=2 & ;;
    dw2)    build dw2 -DL3D_DUO_UNROLL_W=2 & ;;
    dg2w2)  build dg2w2 -DL3D_DUO_UNROLL_G=2 -DL3D_DUO_UNROLL_W=2 & ;;
    r1)     build r1 -DL3D_TPR_R=1 & ;;
    r1dsetp) build r1dsetp -DL3D_TPR_R=1 -DL3D_TPR_DSETP=1 & ;;
    stop[1-4]) build $v -DL3D_TPR_STOP=${v#stop} & ;;
    *) echo "unknown variant $v"; exit 1 ;;
  esac
done
wait
ls -la ../../ab/*.so
