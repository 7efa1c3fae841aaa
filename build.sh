#!/bin/bash
# Builds libbpgpu.so (sm_100a) in-tree.  __graft_entry__.build() calls this.
set -e
cd "$(dirname "$0")/mpc_bulletproof_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -diag-suppress 550 \
     -shared -Xcompiler -fPIC -I../../include -o ../libbpgpu.so bpgpu.cu host/protocol.cpp "$@"
