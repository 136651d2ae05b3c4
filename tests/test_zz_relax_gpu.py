"""Structure relaxation (mlff_b200/md.py: LBFGS, GradientDescent -- mlff/mdx/optimizer.py) driven by the CUDA engine.
The optimiser logic itself is tested on the CPU in tests/test_md.py; this checks it end to end on the device and
verifies the relaxed structure with the float64 oracle."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(__file__))
import so3k_oracle as o
from mlff_b200 import md
from test_md import ethanol_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('flags', [0, 3], ids=['fp32', 'tensor+keep'])
def test_lbfgs_relaxes_ethanol_on_the_device(flags):
    from mlff_b200.config import So3kratesConfig
    from mlff_b200.engine import So3kEngine
    cfg = o.OracleConfig(F=36, n_layers=2)
    params = o.init_params(cfg, 0, embed_scale=4.0)
    eng = So3kEngine(So3kratesConfig(F=36, n_layer=2), 'cuda:0', flags)
    eng.load_params({k: np.asarray(v, np.float32) for k, v in params.items()})
    _, z, _, _, atoms = ethanol_state('cuda:0')
    atoms = atoms.init_spatial_partitioning(eng, cutoff=5.0, skin=0.5)
    calc = md.CalculatorX(eng)
    f0 = float(torch.linalg.norm(calc(atoms)['forces'], dim=-1).max())
    opt = md.LBFGS.create(atoms, calc)
    energies = [float(calc(atoms)['energy'])]
    f_tol = 0.05                       # eV/A; f0 = 1.39.  The float64 oracle-driven run gets there in 30 iterations
    for _ in range(80):                # (0.178 after 25; the device run measured 0.177 fp32 / 0.173 tensor at that point)
        atoms, grads = opt.step(atoms)
        energies.append(float(calc(atoms)['energy']))
        if float(torch.linalg.norm(grads, dim=-1).max()) < f_tol:
            break
    assert eng.check_status() == 0
    tol = 1e-5 * max(abs(e) for e in energies) + 1e-6                    # fp32 energies
    assert all(b <= a + tol for a, b in zip(energies, energies[1:]))
    assert f0 > 1.0 and float(torch.linalg.norm(grads, dim=-1).max()) < f_tol
    assert energies[-1] < energies[0] - 1.0
    # the oracle agrees that the relaxed structure is (nearly) stationary
    R = atoms.get_positions().cpu().numpy()
    nl = o.get_indices(R[None], z[None], 5.0)
    _, F, _ = o.energy_forces(cfg, params, R, z, nl['idx_i'][0], nl['idx_j'][0])
    assert np.abs(F.numpy() + grads.cpu().numpy()).max() < 1e-4
    assert float(np.linalg.norm(F.numpy(), axis=-1).max()) < f_tol + 1e-4
