"""Where do the 64 rows of a cta_group::1 M=64 tcgen05.mma land in TMEM?  (groundwork for two 64-row half tiles per CTA)
   python profiles/tc_m64_layout_probe.py      (one B200; uses so3k_tc_selftest variants 3 / 4)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from mlff_b200 import _lib
lib = _lib.load()
dll = lib.dll
import ctypes as C
dll.so3k_tc_selftest.argtypes = [C.c_int32] * 4 + [C.c_void_p] * 4
dll.so3k_tc_selftest.restype = C.c_int
N, K = 32, 16
rng = np.random.default_rng(0)
A = np.round(rng.normal(size=(128, K)) * 4) / 4          # exactly representable in bf16
B = np.round(rng.normal(size=(N, K)) * 4) / 4
ref = A @ B.T
for variant in (3, 4):
    Ad, Bd = torch.tensor(A, dtype=torch.float32, device='cuda'), torch.tensor(B, dtype=torch.float32, device='cuda')
    D = torch.full((128, N), 777.0, dtype=torch.float32, device='cuda')
    rc = dll.so3k_tc_selftest(N, K, variant, 1, Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), None)
    torch.cuda.synchronize()
    Dh = D.cpu().numpy()
    lanes = {}
    for lane in range(128):
        if np.abs(Dh[lane]).max() == 0:
            continue
        hits = [m for m in range(64) if np.allclose(Dh[lane], ref[m])]
        lanes[lane] = hits[0] if hits else None
    print(f'variant {variant} (D lane offset {0 if variant == 3 else 16}): rc={rc}, non-zero lanes -> GEMM row:')
    print('  ', ', '.join(f'{l}:{m}' for l, m in sorted(lanes.items())))
