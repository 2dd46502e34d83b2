#!/bin/bash
# Build libgwd_b200.so (sm_100a only): one object per .cu, compiled in parallel (gw-depth_b200/csrc/Makefile).
# Used by __graft_entry__.build().
set -e
make -C "$(dirname "$0")/gw-depth_b200/csrc" -j"$(nproc)" "$@"
