"""Per-phase cycle breakdown of the tensor-core edge kernel (debug build with -DTC_TIMING=1, see so3k_edge_tc.cu).
   nvcc <flags> -DTC_TIMING=1 <sources> -o profiles/_scratch/libso3k_timing.so ; python profiles/tc_phase_timing.py"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from mlff_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, 'profiles', '_scratch', 'libso3k_timing.so')
from mlff_b200.config import So3kratesConfig
from mlff_b200.calculator import mlffCalculator
from mlff_b200 import params as P
from common import jittered_lattice, water_like_numbers
cfg = So3kratesConfig(n_layer=3)
calc = mlffCalculator(cfg, P.init_params(cfg, 0), flags=3)
N, box = 100_000, 100.0
pos = jittered_lattice(N, box, 3); z = water_like_numbers(N, 3)
for _ in range(2):
    calc.calculate(pos, z, cell=np.eye(3) * box, pbc=(True, True, True))
torch.cuda.synchronize()
dll = calc.engine.lib.dll
dll.so3k_tc_phase_dump.restype = None
dll.so3k_tc_phase_dump()
sys.stdout.flush()
