#!/bin/sh
# Build the CPU-emulated copy of the SIMT kernels (TEST INFRASTRUCTURE ONLY, see cuda_emu.h).
set -e
cd "$(dirname "$0")/../.."
mkdir -p tests/emu/_build
SRC="so3k_api.cu so3k_graph.cu so3k_feat.cu so3k_node.cu so3k_agg.cu so3k_edge.cu so3k_nl.cu so3k_tc_test.cu so3k_edge_tc.cu so3k_train.cu"
FILES=""
for f in $SRC; do [ -f mlff_b200/csrc/$f ] && FILES="$FILES mlff_b200/csrc/$f"; done
g++ -std=c++20 -O2 -g -DSO3K_EMU -Itests/emu -Imlff_b200/csrc -shared -fPIC -pthread -x c++ $FILES \
    -o tests/emu/_build/libso3k_emu.so
