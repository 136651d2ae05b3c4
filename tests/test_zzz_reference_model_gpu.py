"""The CUDA path DIRECTLY against energies and finite-difference forces produced by the reference's own model code
(tests/golden/reference_model.npz: init_so3krates -> StackNet.__call__ executed from the reference's source files in
float64 on numpy / apply-only flax stand-ins, see tests/golden/make_reference_model_golden.py) -- no oracle arithmetic
in between (the oracle module only supplies the seeded parameter generator the fixture was made with).
The same comparison runs on the CPU-emulated kernels: tests/test_emu_parity.py::test_kernels_match_reference_model_code."""
import numpy as np
import pytest
import torch

import so3k_oracle as o
from conftest import golden
from test_zz_optional_branches_gpu import E_RTOL, F_ATOL, dev, tolerances

pytestmark = pytest.mark.gpu

CASES = {
    'ethanol_default': (dict(), dict(scale_shift=True)),
    'ethanol_1223': (dict(F=32, n_layers=2, degrees=(1, 2, 2, 3), n_heads=2), {}),
    'ethanol_all_branches': (dict(F=36, n_layers=2, layer_normalization=True, final_layer=True, residual_mlp_1=True,
                                  residual_mlp_2=True), {}),
    'ethanol_zbl': (dict(F=36, n_layers=2, zbl_repulsion=True, zbl_repulsion_shift=0.37), dict(scale_shift=True)),
    'solid_periodic': (dict(F=36, n_layers=2), {})}


def engine(ocfg, p, flags):
    from mlff_b200.config import So3kratesConfig
    from mlff_b200.engine import So3kEngine
    cfg = So3kratesConfig(F=ocfg.F, n_layer=ocfg.n_layers, degrees=tuple(ocfg.degrees), n_heads=ocfg.n_heads,
                          n_rbf=ocfg.n_rbf, r_cut=ocfg.r_cut, zbl_repulsion=ocfg.zbl_repulsion,
                          zbl_repulsion_shift=ocfg.zbl_repulsion_shift, layer_normalization=ocfg.layer_normalization,
                          final_layer=ocfg.final_layer, residual_mlp_1=ocfg.residual_mlp_1,
                          residual_mlp_2=ocfg.residual_mlp_2)
    eng = So3kEngine(cfg, 'cuda:0', flags)
    eng.load_params({k: np.asarray(v, np.float32) for k, v in p.items()})
    return eng


@pytest.mark.parametrize('tag,flags', [(t, 0) for t in CASES] + [('ethanol_default', 3), ('solid_periodic', 3)])
def test_cuda_path_matches_reference_model_code(tag, flags):
    g = golden('reference_model.npz')
    ckw, pkw = CASES[tag]
    cfg = o.OracleConfig(**ckw)
    p = o.init_params(cfg, int(g[f'{tag}/seed']), embed_scale=4.0, **pkw)      # the generator's parameters
    ins = {k.split('/in/')[1]: g[k] for k in g.files if k.startswith(tag + '/in/')}
    eng = engine(cfg, p, flags)
    E, F, ea = eng.energy_forces(dev(ins['R'][None], torch.float32), dev(ins['z'][None], torch.int32),
                                 dev(ins['idx_i'][None], torch.int32), dev(ins['idx_j'][None], torch.int32),
                                 dev(ins['unit_cell'][None], torch.float32) if 'unit_cell' in ins else None,
                                 dev(ins['cell_offset'][None], torch.int32) if 'cell_offset' in ins else None,
                                 want_e_atom=True)
    assert eng.check_status() == 0
    E_ref, F_fd, atoms = float(g[f'{tag}/E']), g[f'{tag}/F_fd'], g[f'{tag}/fd_atoms']
    e_rtol, f_atol = tolerances(flags, float(np.abs(F_fd).max()))
    assert abs(float(E.cpu()) - E_ref) <= e_rtol * abs(E_ref)
    assert np.abs(F.cpu().numpy()[0][atoms] - F_fd).max() < f_atol
