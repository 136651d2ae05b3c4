#!/bin/sh
# debug build of libso3k with the per-phase cycle counters of the tensor-core kernels (-DTC_TIMING=1) for profiles/tc_phase_timing.py
set -e
cd "$(dirname "$0")/.."
mkdir -p profiles/_scratch
cd mlff_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -DTC_TIMING=1 \
  so3k_api.cu so3k_graph.cu so3k_feat.cu so3k_node.cu so3k_agg.cu so3k_edge.cu so3k_nl.cu so3k_tc_test.cu so3k_edge_tc.cu so3k_train.cu \
  -o ../../profiles/_scratch/libso3k_timing.so
