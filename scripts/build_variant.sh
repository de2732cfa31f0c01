#!/bin/bash
# build a tuning variant of the library next to the default one:
#   scripts/build_variant.sh oldgj -DNGT_OLD_GJ        -> kmc_rates_b200/csrc/libngt_b200_oldgj.so
# and run anything with it through NGT_B200_LIB=<that path>
name="$1"; shift
cd "$(dirname "$0")/../kmc_rates_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC "$@" -o libngt_b200_${name}.so engine.cu kernels.cu ordering.cu pathsample.cu symbolic.cpp && echo "$(pwd)/libngt_b200_${name}.so"
