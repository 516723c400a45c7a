#!/bin/bash
# usage: profiles/tools/variant.sh NAME "-DX=1 ..."  — builds variants/liborbx_NAME.so (every TU recompiled under the extra defines).
# For A/B timing of kernel variants in ONE gpurun call:
#   for v in a b; do cp variants/liborbx_$v.so tc2li_slam_b200/lib/liborbx.so; profiles/tools/stages.sh; done
set -e
NAME=$1; DEFS=$2
cd "$(dirname "$0")/../../tc2li_slam_b200/csrc"
make -s -j8 VDIR=../../variants/$NAME OUT=../../variants/liborbx_$NAME.so NVFLAGS="-O3 -std=c++17 -lineinfo --fmad=false -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function,-ffp-contract=off $DEFS"
echo built variants/liborbx_$NAME.so
