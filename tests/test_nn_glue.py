"""Host logic of mlff_b200.nn's observable functions on the CPU: key names, single / batched inputs, numpy in -> numpy out,
and get_energy_force_stress_fn (mlff/nn/stacknet/observable_function.py:152-186) -- with the float64 oracle standing in
for the CUDA engine (the engine itself is exercised by tests/test_gpu_parity.py)."""
import numpy as np
import pytest
import torch

import so3k_oracle as o
from conftest import golden
from mlff_b200 import nn
from mlff_b200.config import So3kratesConfig


class OracleBackedEngine:
    device = torch.device('cpu')

    def __init__(self, cfg, p):
        self.cfg, self.p = cfg, p

    def energy_forces(self, R, z, ii, jj, cell=None, off=None, want_e_atom=False, want_stress=False):
        E, F, S = [], [], []
        for b in range(R.shape[0]):
            kw = dict(cell=None if cell is None else cell[b].double().numpy(),
                      cell_offset=None if off is None else off[b].numpy())
            e, f, s = o.energy_forces_stress(self.cfg, self.p, R[b].double().numpy(), z[b].numpy(), ii[b].numpy(),
                                             jj[b].numpy(), **kw)
            E.append(e), F.append(f), S.append(s)
        out = (torch.stack(E).float()[:, None], torch.stack(F).float(), None)
        return out + (torch.stack(S).float(),) if want_stress else out


def _net(cfg, p):
    pc = So3kratesConfig(F=cfg.F, n_layer=cfg.n_layers, degrees=tuple(cfg.degrees), n_heads=cfg.n_heads)
    net = nn.StackNet(pc, device='cpu')
    net.engine = lambda params: OracleBackedEngine(cfg, p)
    return net


def test_energy_force_stress_fn_keys_shapes_and_values():
    sol = golden('solid_0.npz')
    cfg = o.OracleConfig(F=16, n_layers=1, degrees=(1, 2), n_heads=2)
    p = o.init_params(cfg, 2, embed_scale=4.0)
    R, z, cell = sol['R'].astype(np.float64), sol['z'].astype(np.int32), sol['unit_cell'].astype(np.float64)
    ii, jj, S = o.primitive_neighbor_list(R, cell, (True, True, True), cfg.r_cut)
    Eo, Fo, So = o.energy_forces_stress(cfg, p, R, z, ii, jj, cell=cell, cell_offset=S)
    net = _net(cfg, p)
    fn = nn.get_energy_force_stress_fn(net)
    x = {'R': R, 'z': z, 'idx_i': ii, 'idx_j': jj, 'unit_cell': cell, 'cell_offset': S}
    out = fn(None, x)                                              # one structure, numpy in -> numpy out
    assert set(out) == {'E', 'F', 'stress'} and all(isinstance(v, np.ndarray) for v in out.values())
    assert out['E'].shape == (1,) and out['F'].shape == (len(z), 3) and out['stress'].shape == (3, 3)
    assert abs(out['E'][0] - float(Eo)) <= 1e-5 * abs(float(Eo))
    np.testing.assert_allclose(out['F'], Fo.numpy(), atol=1e-4)
    np.testing.assert_allclose(out['stress'], So.numpy(), rtol=1e-4, atol=1e-4 * float(So.abs().max()))
    xb = {k: torch.as_tensor(np.stack([v, v])) for k, v in x.items()}      # batched torch tensors in -> tensors out
    outb = fn(None, xb)
    assert outb['stress'].shape == (2, 3, 3) and outb['E'].shape == (2, 1) and isinstance(outb['F'], torch.Tensor)
    assert torch.equal(outb['stress'][0], outb['stress'][1])
    # the force function of the same net has no stress entry; without a cell the stress function refuses
    assert set(nn.get_obs_and_force_fn(net)(None, x)) == {'E', 'F'}
    with pytest.raises(ValueError):
        fn(None, {k: v for k, v in x.items() if k not in ('unit_cell', 'cell_offset')})
    # renamed property keys travel through (reset_prop_keys of the reference's StackNet)
    net.reset_prop_keys({'stress': 'virial_like', 'energy': 'U'})
    assert set(nn.get_energy_force_stress_fn(net)(None, x)) == {'U', 'F', 'virial_like'}


def test_obs_and_grad_obs_fn_with_the_reference_derivative_spec():
    """get_obs_and_grad_obs_fn(net, derivatives) as the reference's training CLI calls it: the force as the negative
    gradient of the energy with respect to the positions, `lambda y: -y.squeeze(-3)` on the (1, n, 3) gradient."""
    g = golden('ethanol_32.npz')
    cfg = o.OracleConfig(F=16, n_layers=1, degrees=(1, 2), n_heads=2)
    p = o.init_params(cfg, 2, embed_scale=4.0)
    R, z = g['R'][:2].astype(np.float64), np.tile(g['z'][None], (2, 1)).astype(np.int32)
    nl = o.get_indices(R, z, cfg.r_cut)
    net = _net(cfg, p)
    D_force = ('F', ('E', 'R', lambda y: -y.squeeze(-3)))
    fn = nn.get_obs_and_grad_obs_fn(net, derivatives=(D_force,))
    x1 = {'R': R[0], 'z': z[0], 'idx_i': nl['idx_i'][0], 'idx_j': nl['idx_j'][0]}
    out = fn(None, x1)
    Eo, Fo, _ = o.energy_forces(cfg, p, R[0], z[0], nl['idx_i'][0], nl['idx_j'][0])
    assert set(out) == {'E', 'F'} and out['F'].shape == (9, 3)
    np.testing.assert_allclose(out['F'], Fo.numpy(), atol=1e-4)
    assert abs(out['E'][0] - float(Eo)) <= 1e-5 * abs(float(Eo))
    raw = nn.get_grad_observable_fn(net, (('dE_dR', ('E', 'R', lambda y: y)),))(None, x1)['dE_dR']
    assert raw.shape == (1, 9, 3) and np.allclose(raw[0], -out['F'])
    xb = {'R': R, 'z': z, 'idx_i': nl['idx_i'], 'idx_j': nl['idx_j']}
    outb = fn(None, xb)                                              # batched: (B, 1, n, 3) -> squeeze(-3) -> (B, n, 3)
    assert outb['F'].shape == (2, 9, 3) and np.allclose(outb['F'][0], out['F'], atol=1e-6)
    assert nn.get_obs_and_grad_obs_fn(net)(None, x1).keys() == {'E'}
    with pytest.raises(NotImplementedError):
        nn.get_grad_observable_fn(net, (('q', ('E', 'z', lambda y: y)),))
