#!/bin/bash
# build a tuning variant of the library: only diral_step_group.cu is recompiled with the given -D flags, the other
# objects come from the default build.  usage: scripts/build_group_variant.sh NAME -DDIRAL_VPD_HIST=1 ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
C=diral_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall -Xptxas -v \
  "$@" -c $C/diral_step_group.cu -o $C/_obj/group_$name.o > $C/_obj/group_$name.log 2>&1
objs=$(ls $C/_obj/diral_*.o | grep -v diral_step_group.o)
nvcc -shared -o diral_b200/libdiral_env_$name.so $objs $C/_obj/group_$name.o -gencode arch=compute_100a,code=sm_100a -lpthread
echo diral_b200/libdiral_env_$name.so
