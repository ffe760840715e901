#!/bin/bash
# A/B builds of one translation unit:  [TU=rnet] scripts/ab_build.sh NAME [-DFLAG ...]  ->  ab_build/libdmfg_NAME.so
# TU=api (default) rebuilds csrc/dmfg_api.cu (rollout / learner kernels), TU=rnet rebuilds csrc/dmfg_rnet_api.cu; the other
# object is reused from the main build.  Select the library at run time with DMFG_LIB_PATH; ab_build/NAME.ptxas holds
# the ptxas -v resource lines.
set -e
cd "$(dirname "$0")/.."
name=$1; shift
tu=${TU:-api}
if [ "$tu" = rnet ]; then src=dmfg_rnet_api; other=dmfg_api; else src=dmfg_api; other=dmfg_rnet_api; fi
mkdir -p ab_build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
     -c discrete_mean_field_game_b200/csrc/$src.cu -o ab_build/${src}_$name.o > ab_build/$name.ptxas 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ab_build/libdmfg_$name.so ab_build/${src}_$name.o \
     discrete_mean_field_game_b200/build/$other.o
echo ab_build/libdmfg_$name.so
