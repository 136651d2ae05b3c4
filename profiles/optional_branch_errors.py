"""Measured E / F errors of the optional-branch models (LayerNorm, ResidualMLP) against the float64 oracle, per
arithmetic mode.  Run on the GPU box:  python profiles/optional_branch_errors.py > gpurun_out/optional_branch_errors.txt
(test infrastructure: imports the oracle; not part of the product)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle'), os.path.join(ROOT, 'tests')]
import so3k_oracle as o  # noqa: E402
from test_zz_optional_branches_gpu import OPTS, make, dev  # noqa: E402

d = np.load(os.path.join(ROOT, 'tests', 'golden', 'ethanol_32.npz'))
B = 2
R = d['R'][:B]
z = np.tile(d['z'][None], (B, 1))
nl = o.get_indices(R, z, 5.0)
print('model            mode  max|dE|   |E|      sum|e_atom|  max|dF|   max|F|')
for name in ['default'] + list(OPTS):
    cfg = o.OracleConfig(F=132, n_layers=3, **OPTS.get(name, {}))
    p = o.init_params(cfg, 0, embed_scale=4.0, scale_shift=True)
    Eo, Fo = o.energy_forces_batch(cfg, p, R, z, nl['idx_i'], nl['idx_j'])
    Eo, Fo = Eo.numpy().reshape(-1), Fo.numpy()
    for flags in (0, 1, 3):
        eng = make(cfg, p, flags)
        E, F, ea = eng.energy_forces(dev(R, torch.float32), dev(z, torch.int32), dev(nl['idx_i'], torch.int32),
                                     dev(nl['idx_j'], torch.int32), want_e_atom=True)
        E, F, ea = E.cpu().numpy().reshape(-1), F.cpu().numpy(), ea.cpu().numpy()
        print(f'{name:16s} {flags:4d}  {np.abs(E - Eo).max():.2e}  {np.abs(Eo).min():.4f}  {np.abs(ea).sum(-1).min():.4f}'
              f'      {np.abs(F - Fo).max():.2e}  {np.abs(Fo).max():.3f}')
