"""Golden loss values and parameter-gradient projections from the REFERENCE'S OWN TRAINING CODE.

The reference's loss (mlff/training/loss.py: scaled_safe_masked_mse_loss :13-29, get_loss_fn :42-73) is executed
unmodified from /root/reference on the numpy stand-ins of ref_shim.py, on top of an obs_fn that evaluates the reference's
own model code (init_so3krates + StackNet.__call__, as in make_reference_model_golden.py) with forces by float64 central
differences (the reference uses jax.jacrev, observable_function.py:127-149; jax is absent).  jax.value_and_grad over
3e4 parameters cannot be replaced by finite differences entry by entry, so the gradient is recorded as PROJECTIONS:
for a few seeded random parameter directions v the central difference (loss(theta + h v) - loss(theta - h v)) / 2h of the
reference loss.  tests/test_training.py compares <oracle grad, v> (torch double backward) and, through the oracle, the
CUDA kernels' gradient blob with these numbers.

Batch: two ethanol frames + one padding atom each, padded pair slots, one NaN energy target and NaN force targets on
one atom (both must be masked out of numerator and denominator).

    python tests/golden/make_reference_training_golden.py        (build container only; needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [HERE, ROOT, os.path.join(ROOT, 'oracle')]
import ref_shim as rs  # noqa: E402
from make_reference_model_golden import load_reference, PROP_KEYS, _f64  # noqa: E402

W_ENERGY, W_FORCE = 0.01, 0.99
N_DIRECTIONS, H_THETA, H_R = 4, 1e-3, 1e-4


def main():
    so3krates = load_reference()
    import types
    sys.modules['mlff.training'] = types.ModuleType('mlff.training')      # declared without running its __init__
    sys.modules['mlff.training'].__path__ = []
    loss_mod = rs.load('mlff.training.loss')
    import so3k_oracle as o
    from mlff_b200 import params as P
    from mlff_b200.config import So3kratesConfig

    eth = np.load(os.path.join(rs.REF, 'tests/test_data/ethanol.npz'))
    ocfg = o.OracleConfig(F=32, n_rbf=8, degrees=(1, 2), n_heads=4, n_layers=2)
    pcfg = So3kratesConfig(F=32, n_rbf=8, n_layer=2, degrees=(1, 2), n_heads=4)
    seed = 11
    p = o.init_params(ocfg, seed, embed_scale=4.0, scale_shift=True)
    B = 2
    R = np.concatenate([eth['R'][[0, 7]].astype(np.float64), np.zeros((B, 1, 3))], 1)
    z = np.concatenate([np.tile(eth['z'][None].astype(np.int64), (B, 1)), np.zeros((B, 1), np.int64)], 1)
    nl = o.get_indices(R, z, 5.0)
    ii = np.concatenate([nl['idx_i'], -np.ones((B, 3), np.int64)], 1)
    jj = np.concatenate([nl['idx_j'], -np.ones((B, 3), np.int64)], 1)
    rng = np.random.default_rng(17)
    E_true = rng.normal(size=(B, 1)) - 25.0
    F_true = rng.normal(size=(B, R.shape[1], 3))
    F_true[0, 4] = np.nan                         # a real atom without a force label
    node_mask = z != 0

    Energy = sys.modules['mlff.nn.observable'].Energy

    def make_net():       # a fresh module tree per evaluation (the apply-only stand-in binds parameters when first applied)
        obs = [Energy(prop_keys=PROP_KEYS, per_atom_scale=list(p['energy/per_atom_scale']),
                      per_atom_shift=list(p['energy/per_atom_shift']))]
        return so3krates.init_so3krates(PROP_KEYS, F=ocfg.F, n_layer=ocfg.n_layers, obs=obs,
                                        so3krates_layer_kwargs={'degrees': list(ocfg.degrees), 'n_heads': ocfg.n_heads},
                                        geometry_embed_kwargs={'degrees': list(ocfg.degrees), 'n_rbf': ocfg.n_rbf,
                                                               'r_cut': ocfg.r_cut, 'mic': False})

    def tree_of(flat):
        t = P.to_flax_tree(pcfg, {k: np.asarray(v, np.float64) for k, v in flat.items()})
        return {'params': _f64(t['params'], flat, pcfg)}

    def obs_fn(flat, inputs):
        """vmap(get_obs_and_force_fn(net)) (observable_function.py:127-149): E (B,1), F (B,n,3) = -dE/dR."""
        tree = tree_of(flat)
        Es, Fs = [], []
        for b in range(B):
            def energy(Rb):
                q = {'R': Rb, 'z': inputs['z'][b], 'idx_i': inputs['idx_i'][b], 'idx_j': inputs['idx_j'][b]}
                return float(np.asarray(make_net().apply(tree, q)['E']).reshape(()))
            R0 = np.asarray(inputs['R'][b], np.float64)
            Es.append([energy(R0)])
            F = np.zeros_like(R0)
            for a in range(R0.shape[0]):
                if inputs['z'][b][a] == 0:
                    continue                      # a padding atom has no edges: its force is exactly zero
                for c in range(3):
                    Rp, Rm = R0.copy(), R0.copy()
                    Rp[a, c] += H_R
                    Rm[a, c] -= H_R
                    F[a, c] = -(energy(Rp) - energy(Rm)) / (2 * H_R)
            Fs.append(F)
        return {'E': np.array(Es), 'F': np.array(Fs)}

    loss_fn = loss_mod.get_loss_fn(obs_fn=obs_fn, weights={'energy': W_ENERGY, 'force': W_FORCE}, prop_keys=PROP_KEYS)
    inputs = {'R': R, 'z': z, 'idx_i': ii, 'idx_j': jj, 'node_mask': node_mask}
    out = {'seed': np.array(seed), 'w_energy': np.array(W_ENERGY), 'w_force': np.array(W_FORCE), 'h_theta': np.array(H_THETA)}
    for k, v in inputs.items():
        out['in/' + k] = v

    names = [k for k in p if not k.startswith('energy/per_atom_')]
    for tag, Et in (('full', E_true), ('nan_energy', np.where(np.arange(B)[:, None] == 1, np.nan, E_true))):
        targets = {'E': Et, 'F': F_true}
        loss, metrics = loss_fn(p, (inputs, targets))
        out[f'{tag}/E_true'], out[f'{tag}/F_true'] = Et, F_true
        out[f'{tag}/loss'] = np.array(float(loss))
        out[f'{tag}/mse_E'], out[f'{tag}/mse_F'] = np.array(float(metrics['E'])), np.array(float(metrics['F']))
        print(tag, 'loss', float(loss), 'E', float(metrics['E']), 'F', float(metrics['F']))
        if tag != 'full':
            continue
        # projections of the parameter gradient on seeded random directions (unit norm over all trainable leaves)
        proj = []
        for k in range(N_DIRECTIONS):
            drng = np.random.default_rng(100 + k)
            v = {n: drng.normal(size=np.shape(p[n])) for n in names}
            nrm = np.sqrt(sum(float((a ** 2).sum()) for a in v.values()))
            vals = []
            for sgn in (1.0, -1.0):
                q = dict(p)
                for n in names:
                    q[n] = np.asarray(p[n], np.float64) + sgn * H_THETA * v[n] / nrm
                vals.append(float(loss_fn(q, (inputs, targets))[0]))
            proj.append((vals[0] - vals[1]) / (2 * H_THETA))
            print('direction', k, 'dloss/dtheta . v =', proj[-1])
        out['full/grad_projections'] = np.array(proj)
    np.savez_compressed(os.path.join(HERE, 'reference_training.npz'), **out)
    print('wrote reference_training.npz')


if __name__ == '__main__':
    main()
