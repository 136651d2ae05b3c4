#!/bin/sh
# The CPU-emulated kernels under AddressSanitizer (TEST INFRASTRUCTURE): builds tests/emu/_build/libso3k_emu_asan.so from the
# same kernel sources and runs the emulator-driven suites against it.  Every global-memory access of a kernel is then a
# checked access to a numpy / torch buffer, every shared-memory array an instrumented stack or static array: an index
# that runs past a workspace slice, a row or a tile is reported instead of silently reading a neighbour.
#   sh tests/emu/asan_check.sh            (about 12 minutes; last run: 69 passed, no reports)
set -e
cd "$(dirname "$0")/../.."
mkdir -p tests/emu/_build
SRC="so3k_api.cu so3k_graph.cu so3k_feat.cu so3k_node.cu so3k_agg.cu so3k_edge.cu so3k_nl.cu so3k_tc_test.cu so3k_edge_tc.cu so3k_train.cu"
FILES=""
for f in $SRC; do [ -f mlff_b200/csrc/$f ] && FILES="$FILES mlff_b200/csrc/$f"; done
g++ -std=c++20 -O1 -g -fsanitize=address -fno-omit-frame-pointer -DSO3K_EMU -Itests/emu -Imlff_b200/csrc -shared -fPIC -pthread \
    -x c++ $FILES -o tests/emu/_build/libso3k_emu_asan.so
cat > tests/emu/_build/asan_run.py <<'PY'
import os, sys
root = os.getcwd()
for p in ('tests', 'oracle', ''):
    sys.path.insert(0, os.path.join(root, p))
import emu_engine
so = os.path.join(root, 'tests', 'emu', '_build', 'libso3k_emu_asan.so')
emu_engine.EMU_SO = so
emu_engine.build_emu = lambda force=False: so
import pytest
sys.exit(pytest.main(['-x', '-q', '-p', 'no:cacheprovider'] + sys.argv[1:]))
PY
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
    python tests/emu/_build/asan_run.py tests/test_emu_parity.py tests/test_loss_kernels.py tests/test_training.py tests/test_dd.py \
    -k "not gloo"
