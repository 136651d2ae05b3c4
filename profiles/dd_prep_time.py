import sys, os, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from mlff_b200.config import So3kratesConfig
from mlff_b200.calculator import mlffCalculator
from mlff_b200 import params as P, dist as dd
from mlff_b200.engine import NL_ASE
from common import jittered_lattice, water_like_numbers
dev='cuda:0'
cfg = So3kratesConfig(n_layer=3)
calc = mlffCalculator(cfg, P.init_params(cfg, 0), device=dev, flags=3)
eng = calc.engine
N, box = 100_000, 100.0
pos = jittered_lattice(N, box, 3); z = water_like_numbers(N, 3)
cell = np.eye(3) * box; pbc=(True,True,True)
pos_d = torch.as_tensor(pos, device=dev); z_d = torch.as_tensor(z, device=dev, dtype=torch.int32)
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
rank = 0
cap = int(5.3e6 * 1.6 / world) + 4096
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for rep in range(3):
    t0 = T()
    keep = dd.brick_halo_mask(pos_d, cell, world, rank, 5.0); t1 = T()
    sub = torch.nonzero(keep).reshape(-1); pos_s, z_s = pos_d[sub], z_d[sub]; t2 = T()
    ii, jj, S, count = eng.neighbor_list(pos_s, 5.0, cap, cell=cell, pbc=pbc, mode=NL_ASE); nv = int(count.item()); t3 = T()
    owner = dd.assign_owners_bricks(pos_s, cell, world); t4 = T()
    plan = dd.build_plan(ii[:nv].long(), jj[:nv].long(), S[:nv], owner, rank, world, check=False); t5 = T()
    print(f'world {world}: mask {1e3*(t1-t0):.3f} subset {1e3*(t2-t1):.3f} NL {1e3*(t3-t2):.3f} owners {1e3*(t4-t3):.3f} plan {1e3*(t5-t4):.3f} ms | n_sub {len(sub)} edges {nv} own {plan.n_own} ghost {plan.n_ghost}')
# inside build_plan
import torch
iiL, jjL = ii[:nv].long(), jj[:nv].long()
def timeit(label, fn):
    t0 = T(); r = fn(); t1 = T(); print(f'   {label}: {1e3*(t1-t0):.3f} ms'); return r
timeit('long conversions', lambda: (ii[:nv].long(), jj[:nv].long()))
oi = timeit('owner[ii]', lambda: owner[iiL]); oj = timeit('owner[jj]', lambda: owner[jjL])
own_recv = oi == rank; mirror = (oj == rank) & (oi != rank)
gh = timeit('unique ghosts', lambda: torch.unique(jjL[own_recv & (oj != rank)]))
timeit('argsort ghosts', lambda: gh[torch.argsort(owner[gh] * len(owner) + gh)])
timeit('nonzero owned', lambda: torch.nonzero(owner == rank).reshape(-1))
sel = own_recv | mirror
timeit('select edges', lambda: (iiL[sel], jjL[sel], S[:nv][sel]))
timeit('unique need', lambda: torch.unique(oi[mirror] * len(owner) + jjL[mirror]))
