#!/bin/sh
# The CPU-emulated kernels under AddressSanitizer or ThreadSanitizer (TEST INFRASTRUCTURE): builds
# tests/emu/_build/libso3k_emu_<san>.so from the same kernel sources and runs the emulator-driven suites against it.
#   address: every global-memory access of a kernel is a checked access to a numpy / torch buffer, every shared-memory array
#            an instrumented stack or static array -- an index that runs past a workspace slice, a row or a tile is reported
#            instead of silently reading a neighbour;
#   thread:  the threads of a block are OS threads that only synchronise where the kernel says so (__syncthreads /
#            __syncwarp are std::barrier, a shuffle is a warp-wide rendezvous), so two threads of a block touching the same
#            shared or global word without one of those in between is reported as a data race.  (Reports whose two stacks
#            are torch's OpenMP workers -- libgomp is not instrumented -- are not about the kernels; the script counts the
#            reports that have an emulated-kernel frame on BOTH sides.)
#   undefined: misaligned vector loads / stores (the reinterpret_cast<float4*> accesses), signed overflow, bad shifts.
#   sh tests/emu/sanitizer_check.sh address     (about 12 minutes; last run on the final sources: 69 passed, no reports)
#   sh tests/emu/sanitizer_check.sh thread      (about 17 minutes; last run: 68 passed, no kernel-vs-kernel race)
#   sh tests/emu/sanitizer_check.sh undefined   (about 5 minutes; last run: 68 passed, no reports)
set -e
SAN=${1:-address}
cd "$(dirname "$0")/../.."
mkdir -p tests/emu/_build
SRC="so3k_api.cu so3k_graph.cu so3k_feat.cu so3k_node.cu so3k_agg.cu so3k_edge.cu so3k_nl.cu so3k_tc_test.cu so3k_edge_tc.cu so3k_train.cu"
FILES=""
for f in $SRC; do [ -f mlff_b200/csrc/$f ] && FILES="$FILES mlff_b200/csrc/$f"; done
SO=tests/emu/_build/libso3k_emu_$SAN.so
g++ -std=c++20 -O1 -g -fsanitize=$SAN -fno-sanitize=vptr -fno-omit-frame-pointer -DSO3K_EMU -Itests/emu -Imlff_b200/csrc -shared -fPIC -pthread \
    -x c++ $FILES -o $SO
cat > tests/emu/_build/san_run.py <<'PY'
import os, sys
root = os.getcwd()
for p in ('tests', 'oracle', ''):
    sys.path.insert(0, os.path.join(root, p))
import emu_engine
so = os.path.join(root, os.environ['SO3K_SAN_SO'])
emu_engine.EMU_SO = so
emu_engine.build_emu = lambda force=False: so
import pytest
sys.exit(pytest.main(['-x', '-q', '-p', 'no:cacheprovider'] + sys.argv[1:]))
PY
SUITES="tests/test_emu_parity.py tests/test_loss_kernels.py tests/test_training.py tests/test_dd.py"
if [ "$SAN" = undefined ]; then
  rm -f tests/emu/_build/ubsan_log*
  SO3K_SAN_SO=$SO LD_PRELOAD=$(gcc -print-file-name=libubsan.so) UBSAN_OPTIONS=print_stacktrace=1:log_path=tests/emu/_build/ubsan_log \
      python tests/emu/_build/san_run.py $SUITES -k "not gloo and not trainer_step"
  echo "undefined-behaviour reports: $(cat tests/emu/_build/ubsan_log* 2>/dev/null | grep -c 'runtime error')"
elif [ "$SAN" = address ]; then
  SO3K_SAN_SO=$SO LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 \
      python tests/emu/_build/san_run.py $SUITES -k "not gloo"
else
  rm -f tests/emu/_build/tsan_log*
  SO3K_SAN_SO=$SO LD_PRELOAD=$(gcc -print-file-name=libtsan.so) \
      TSAN_OPTIONS=halt_on_error=0:report_signal_unsafe=0:exitcode=0:log_path=tests/emu/_build/tsan_log \
      python tests/emu/_build/san_run.py $SUITES -k "not gloo and not trainer_step"
  python - <<'PY'
import glob
n = 0
for f in glob.glob('tests/emu/_build/tsan_log*'):
    for rep in open(f).read().split('==================\n'):
        if 'WARNING: ThreadSanitizer' not in rep:
            continue
        sides = rep.split('Previous ')
        if len(sides) == 2 and all('libso3k_emu_thread' in s and 'libtorch' not in s and 'libgomp' not in s for s in sides):
            n += 1
            print(rep[:3000])
print('data races with emulated-kernel frames on both sides:', n)
PY
fi
