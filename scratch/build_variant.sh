#!/bin/bash
# scratch/build_variant.sh <name> <nvcc -D flags...>: variant of libmopoe_b200.so with mopoe_daa.cu recompiled
set -e
name=$1; shift
cd /root/repo/2022_cambroise_interpret_multivae_b200/csrc
make -j4 >/dev/null
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c mopoe_daa.cu -o /tmp/daa_$name.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o /root/repo/scratch/variants/lib_$name.so mopoe_capi.o mopoe_model.o /tmp/daa_$name.o mopoe_rsa.o -lcudart
echo built scratch/variants/lib_$name.so
