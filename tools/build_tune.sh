#!/bin/sh
# Experiment build of the kernel library (-DGNN_TUNE: plan parameters from GNN_TUNE_* environment variables,
# gather-roof microbenchmark).  Used by tools/tune_spmm.py and tools/gather_roof.py; never shipped.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/_build
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -Xlinker -Bsymbolic -Iinclude -DGNN_TUNE \
     -Ignn_b200/csrc -o tools/_build/libgnn_b200_tune.so gnn_b200/csrc/gnn_kernels.cu
