#!/bin/bash
# A/B builds of the rollout kernels: scripts/ab_build.sh NAME [-DFLAG ...]  ->  ab_build/libdmfg_NAME.so
# (select it at run time with DMFG_LIB_PATH; the reward-net translation unit is reused from the main build);
# ab_build/NAME.ptxas holds the ptxas -v resource lines
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p ab_build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
     -c discrete_mean_field_game_b200/csrc/dmfg_api.cu -o ab_build/dmfg_api_$name.o > ab_build/$name.ptxas 2>&1
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ab_build/libdmfg_$name.so ab_build/dmfg_api_$name.o \
     discrete_mean_field_game_b200/build/dmfg_rnet_api.o
echo ab_build/libdmfg_$name.so
