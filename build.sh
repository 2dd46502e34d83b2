#!/bin/bash
# Build libgwd_b200.so (sm_100a only).  Used by __graft_entry__.build().
set -e
cd "$(dirname "$0")/gw-depth_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
     -o ../libgwd_b200.so gwd_core.cu gwd_gemm.cu gwd_attn.cu gwd_attn_tc.cu gwd_attn_win.cu gwd_elem.cu gwd_select.cu gwd_stem.cu gwd_train.cu gwd_wgrad_tc.cu gwd_train_win.cu gwd_lsap.cu "$@"
