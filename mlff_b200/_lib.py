"""Locate / build / load the CUDA shared library (the ONLY implementation of the hot path).

There is no CPU fallback: if ``mlff_b200/lib/libso3k.so`` is missing the import of the engine fails loudly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from typing import List, Optional

from .capi import So3kLib

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, 'csrc')
# SO3K_LIB_PATH: another build of the SAME library (profiling / A-B builds under profiles/_scratch); never a different backend
LIB_PATH = os.environ.get('SO3K_LIB_PATH') or os.path.join(PKG_DIR, 'lib', 'libso3k.so')
SOURCES = ['so3k_api.cu', 'so3k_graph.cu', 'so3k_feat.cu', 'so3k_node.cu', 'so3k_agg.cu', 'so3k_edge.cu',
           'so3k_nl.cu', 'so3k_tc_test.cu', 'so3k_edge_tc.cu', 'so3k_train.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']

_lib: Optional[So3kLib] = None


def sources() -> List[str]:
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(PKG_DIR, '..', 'include', 'so3k.h')]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into mlff_b200/lib/libso3k.so (in-tree, travels with gpurun)."""
    if not force and not needs_build():
        return LIB_PATH
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + sources() + ['-o', LIB_PATH + '.tmp']
    subprocess.check_call(cmd)
    os.replace(LIB_PATH + '.tmp', LIB_PATH)
    return LIB_PATH


# the files that define the kernels bench.py reports a roofline for (k_filter_tc, k_edge_bwd2_tc, k_alpha_aggregate_scan,
# k_qk_bwd_stream, k_alpha_bwd) and the device-code headers they include
PROFILED_SOURCES = ['so3k_edge_tc.cu', 'so3k_tc.cuh', 'so3k_agg.cu', 'so3k_math.h', 'so3k_portable.h']


def source_sha16() -> str:
    """sha256 (first 16 hex digits) over the sources of the profiled kernels: identifies the kernel code an ncu capture
    was taken on (the binary itself is not bit-reproducible across nvcc runs)."""
    import hashlib
    h = hashlib.sha256()
    for name in PROFILED_SOURCES:
        h.update(name.encode())
        h.update(open(os.path.join(CSRC, name), 'rb').read())
    return h.hexdigest()[:16]


def load() -> So3kLib:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f'{LIB_PATH} is missing: the CUDA extension has not been built (run `python -c "import '
                f'__graft_entry__ as g; g.build()"` or mlff_b200._lib.build()). mlff_b200 has no CPU fallback.')
        _lib = So3kLib(LIB_PATH)
    return _lib
