#!/bin/bash
# build_attn_variant.sh <tag> <extra nvcc flags...>: libapex_b200_<tag>.so = the in-tree objects with attention.cu rebuilt with extra defines
set -e
cd "$(dirname "$0")/../apex-studio_b200/csrc"
tag=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c attention.cu -o /tmp/attention_$tag.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libapex_b200_$tag.so abi.o elementwise.o linear.o /tmp/attention_$tag.o conv.o vae_ops.o mmdit_ops.o wan_vae.o -lcudart
echo built libapex_b200_$tag.so
