"""Golden energies and forces from the REFERENCE'S OWN MODEL CODE.

jax / flax are absent from this image, so the reference cannot run as shipped.  This script executes the reference's
source files unmodified (loaded by path from /root/reference) on top of the stand-ins of tests/golden/ref_shim.py --
numpy for jax.numpy, and an apply-only re-implementation of the handful of flax.linen pieces the SO3krates path uses
(Module naming, Dense, Embed, LayerNorm, vmap over stacked parameters).  The network is built by the reference's own
`init_so3krates` (mlff/nn/representation/so3krates.py:13-49) and evaluated through `StackNet.__call__`
(mlff/nn/stacknet/stacknet.py:40-87), i.e. every arithmetic line of the model is the reference's, in float64.

Parameters: oracle.init_params(seed) laid out as the reference's flax tree by mlff_b200.params.to_flax_tree -- every
leaf is looked up by the path flax would use, so this also verifies that name map.
Forces: central finite differences of the reference energy in float64 (step 1e-4 A, error ~1e-8).

    python tests/golden/make_reference_model_golden.py           (build container only; needs /root/reference)

Writes tests/golden/reference_model.npz; tests/test_oracle.py::test_oracle_matches_reference_model_code compares.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [HERE, ROOT, os.path.join(ROOT, 'oracle')]
import ref_shim as rs  # noqa: E402

PROP_KEYS = {'energy': 'E', 'force': 'F', 'atomic_type': 'z', 'atomic_position': 'R', 'unit_cell': 'unit_cell',
             'cell_offset': 'cell_offset', 'displacement_vector': 'displacements', 'atomic_energy': 'atomic_energy',
             'idx_i': 'idx_i', 'idx_j': 'idx_j', 'node_mask': 'node_mask'}


def load_reference():
    rs.install()
    rs.load('mlff.masking.mask')
    rs.load('mlff.properties.property_names')
    sys.modules['mlff.properties'].property_names = sys.modules['mlff.properties.property_names']
    rs.load('mlff.basis_function.spherical')
    rs.load('mlff.basis_function.radial')
    pbc = rs.load('mlff.cutoff_function.pbc')
    rad = rs.load('mlff.cutoff_function.radial')
    rs.export('mlff.cutoff_function', pbc, 'add_cell_offsets')
    rs.export('mlff.cutoff_function', rad, 'get_cutoff_fn')
    rs.export('mlff.sph_ops', rs.load('mlff.sph_ops.contract'), 'make_l0_contraction_fn')
    rs.load('mlff.nn.base.sub_module')
    rs.load('mlff.nn.activation_function.activation_function')
    rs.export('mlff.nn.mlp', rs.load('mlff.nn.mlp.mlp'), 'MLP', 'ResidualMLP')
    rs.export('mlff.nn.embed', rs.load('mlff.nn.embed.embed'), 'AtomTypeEmbed', 'GeometryEmbed')
    rs.export('mlff.nn.layer', rs.load('mlff.nn.layer.so3krates_layer'), 'So3kratesLayer')
    rs.export('mlff.nn.observable', rs.load('mlff.nn.observable.observable'), 'Energy')
    # stacknet.py imports three factory look-ups and read_json that only its JSON constructor uses
    for pkg, name in (('mlff.nn.layer', 'get_layer'), ('mlff.nn.embed', 'get_embedding_module'),
                      ('mlff.nn.observable', 'get_observable_module'), ('mlff.io', 'read_json')):
        setattr(sys.modules[pkg], name, None)
    rs.export('mlff.nn.stacknet', rs.load('mlff.nn.stacknet.stacknet'), 'StackNet')
    return rs.load('mlff.nn.representation.so3krates')


def main():
    so3krates = load_reference()
    import so3k_oracle as o
    from mlff_b200 import params as P
    from mlff_b200.config import So3kratesConfig

    eth = np.load(os.path.join(rs.REF, 'tests/test_data/ethanol.npz'))
    sol = np.load(os.path.join(rs.REF, 'tests/test_data/test_solid.npz'))
    out = {}

    def run(tag, ocfg, seed, inputs, mic, fd_atoms, layer_kwargs=None, scale_shift=False, energy_kwargs=None):
        p = o.init_params(ocfg, seed, embed_scale=4.0, scale_shift=scale_shift)
        pcfg = So3kratesConfig(F=ocfg.F, n_layer=ocfg.n_layers, degrees=tuple(ocfg.degrees), n_heads=ocfg.n_heads,
                               n_rbf=ocfg.n_rbf, r_cut=ocfg.r_cut, **(layer_kwargs or {}), **(energy_kwargs or {}))
        tree = P.to_flax_tree(pcfg, {k: np.asarray(v, np.float64) for k, v in p.items()})
        tree = {'params': _f64(tree['params'], p, pcfg)}
        lk = {'degrees': list(ocfg.degrees), 'n_heads': ocfg.n_heads}
        lk.update(layer_kwargs or {})
        Energy = sys.modules['mlff.nn.observable'].Energy
        obs = [Energy(prop_keys=PROP_KEYS, per_atom_scale=list(p['energy/per_atom_scale']),
                      per_atom_shift=list(p['energy/per_atom_shift']), **(energy_kwargs or {}))]

        def energy(R):
            net = so3krates.init_so3krates(PROP_KEYS, F=ocfg.F, n_layer=ocfg.n_layers, obs=obs,
                                           so3krates_layer_kwargs=lk,
                                           geometry_embed_kwargs={'degrees': list(ocfg.degrees), 'n_rbf': ocfg.n_rbf,
                                                                  'r_cut': ocfg.r_cut, 'mic': mic})
            q = dict(inputs)
            q['R'] = R
            return float(np.asarray(net.apply(tree, q)['E']).reshape(()))

        R0 = np.asarray(inputs['R'], np.float64)
        out[f'{tag}/E'] = np.array(energy(R0))
        h = 1e-4
        F = np.zeros((len(fd_atoms), 3))
        for a, atom in enumerate(fd_atoms):
            for c in range(3):
                Rp, Rm = R0.copy(), R0.copy()
                Rp[atom, c] += h
                Rm[atom, c] -= h
                F[a, c] = -(energy(Rp) - energy(Rm)) / (2 * h)
        out[f'{tag}/F_fd'], out[f'{tag}/fd_atoms'] = F, np.array(fd_atoms)
        out[f'{tag}/seed'] = np.array(seed)
        for k, v in inputs.items():
            out[f'{tag}/in/{k}'] = np.asarray(v)
        print(tag, 'E =', out[f'{tag}/E'])

    # ethanol, default model (F = 132, 3 layers, degrees 1-3, 4 heads) with padding atoms / pairs
    R, z = eth['R'][0].astype(np.float64), eth['z'].astype(np.int64)
    nl = o.get_indices(R[None], z[None], 5.0)
    ii = np.concatenate([nl['idx_i'][0], [-1, -1]])
    jj = np.concatenate([nl['idx_j'][0], [-1, -1]])
    mol = {'R': np.concatenate([R, np.zeros((1, 3))]), 'z': np.concatenate([z, [0]]), 'idx_i': ii, 'idx_j': jj}
    run('ethanol_default', o.OracleConfig(), 0, mol, False, list(range(9)), scale_shift=True)
    # other degree list / heads
    run('ethanol_1223', o.OracleConfig(F=32, n_layers=2, degrees=(1, 2, 2, 3), n_heads=2), 1, mol, False, [0, 4, 8])
    # every optional layer branch
    allk = dict(layer_normalization=True, final_layer=True, residual_mlp_1=True, residual_mlp_2=True)
    run('ethanol_all_branches', o.OracleConfig(F=36, n_layers=2, **allk), 2, mol, False, [0, 3, 7], layer_kwargs=allk)
    # ZBL repulsion (Energy(zbl_repulsion=True), observable.py:81-84, 110-183), per_structure shift
    zk = dict(zbl_repulsion=True, zbl_repulsion_shift=0.37)
    run('ethanol_zbl', o.OracleConfig(F=36, n_layers=2, **zk), 4, mol, False, [0, 2, 5], scale_shift=True, energy_kwargs=zk)
    # periodic solid with cell offsets (all images within 5 A of the 9 A cell)
    pos, cell, zs = sol['R'][0].astype(np.float64), sol['unit_cell'][0].astype(np.float64), sol['z'][0].astype(np.int64)
    si, sj, S = o.primitive_neighbor_list(pos, cell, [True, True, True], 5.0)
    per = {'R': pos, 'z': zs, 'idx_i': si, 'idx_j': sj, 'unit_cell': cell, 'cell_offset': S}
    run('solid_periodic', o.OracleConfig(F=36, n_layers=2), 3, per, True, [0, 57])

    # ---- MD seam: what MLFFPotential.potential_fn computes (mdx/potential/mlff_potential.py:58-109): displacements in,
    # per-atom energies out, masked edges -> idx -1, energy scale applied outside the net
    def run_md_seam(tag, ocfg, seed, energy_kwargs, scale):
        p = o.init_params(ocfg, seed, embed_scale=4.0, scale_shift=True)
        pcfg = So3kratesConfig(F=ocfg.F, n_layer=ocfg.n_layers, degrees=tuple(ocfg.degrees), n_heads=ocfg.n_heads,
                               **energy_kwargs)
        tree = P.to_flax_tree(pcfg, {k: np.asarray(v, np.float64) for k, v in p.items()})
        tree = {'params': _f64(tree['params'], p, pcfg)}
        Energy = sys.modules['mlff.nn.observable'].Energy
        i0, j0 = nl['idx_i'][0], nl['idx_j'][0]
        edges0 = R[j0] - R[i0]
        mask = np.ones(len(i0), bool)
        mask[[3, 17]] = False                                  # two edges switched off by the graph mask (not symmetric
        mask[[k for k in range(len(i0)) if (i0[k], j0[k]) in ((j0[3], i0[3]), (j0[17], i0[17]))]] = False  # ... made so)

        def atomic_energies(edges):
            obs = [Energy(prop_keys=PROP_KEYS, per_atom_scale=list(p['energy/per_atom_scale']),
                          per_atom_shift=list(p['energy/per_atom_shift']), **energy_kwargs)]
            net = so3krates.init_so3krates(PROP_KEYS, F=ocfg.F, n_layer=ocfg.n_layers, obs=obs,
                                           so3krates_layer_kwargs={'degrees': list(ocfg.degrees), 'n_heads': ocfg.n_heads},
                                           geometry_embed_kwargs={'degrees': list(ocfg.degrees)})
            net.reset_input_convention('displacements')        # mlff_potential.py:65-66
            net.reset_output_convention('per_atom')
            x = {'displacements': edges, 'z': z, 'idx_i': np.where(mask, i0, -1), 'idx_j': np.where(mask, j0, -1)}
            return scale * np.asarray(net.apply(tree, x)['atomic_energy']).reshape(-1)     # :86-87, :104-107

        out[f'{tag}/e_atom'] = atomic_energies(edges0)
        fd_edges = [0, 3, 10, 40, 71]
        h = 1e-4
        g = np.zeros((len(fd_edges), 3))
        for a, e in enumerate(fd_edges):
            for c in range(3):
                ep, em = edges0.copy(), edges0.copy()
                ep[e, c] += h
                em[e, c] -= h
                g[a, c] = (atomic_energies(ep).sum() - atomic_energies(em).sum()) / (2 * h)
        out[f'{tag}/dE_dedges_fd'], out[f'{tag}/fd_edges'] = g, np.array(fd_edges)
        out[f'{tag}/seed'], out[f'{tag}/scale'] = np.array(seed), np.array(scale)
        for k, v in (('edges', edges0), ('z', z), ('centers', i0), ('others', j0), ('mask', mask)):
            out[f'{tag}/in/{k}'] = v
        print(tag, 'sum e_atom =', out[f'{tag}/e_atom'].sum())

    # ---- stress: get_energy_force_stress_fn (mlff/nn/stacknet/observable_function.py:152-186) run as written, with
    # jax.value_and_grad replaced by central differences; a sparse periodic system (30 atoms of the solid in its cell)
    obs_fn_mod = rs.load('mlff.nn.stacknet.observable_function')
    scfg = o.OracleConfig(F=32, n_layers=2, degrees=(1, 2), n_heads=2)
    sp = o.init_params(scfg, 7, embed_scale=4.0)
    spc = So3kratesConfig(F=32, n_layer=2, degrees=(1, 2), n_heads=2)
    stree = {'params': _f64(P.to_flax_tree(spc, {k: np.asarray(v, np.float64) for k, v in sp.items()})['params'], sp, spc)}
    Rp, zp = pos[:30], zs[:30]
    qi, qj, qS = o.primitive_neighbor_list(Rp, cell, [True, True, True], 5.0)
    keys = dict(PROP_KEYS, stress='stress')
    snet = so3krates.init_so3krates(keys, F=32, n_layer=2,
                                    so3krates_layer_kwargs={'degrees': [1, 2], 'n_heads': 2},
                                    geometry_embed_kwargs={'degrees': [1, 2], 'mic': True})
    res = obs_fn_mod.get_energy_force_stress_fn(snet)(stree, {'R': Rp, 'z': zp, 'idx_i': qi, 'idx_j': qj,
                                                              'unit_cell': cell, 'cell_offset': qS})
    out['stress/E'], out['stress/F'], out['stress/stress'] = (np.asarray(res[k], np.float64) for k in ('E', 'F', 'stress'))
    for k, v in (('R', Rp), ('z', zp), ('idx_i', qi), ('idx_j', qj), ('unit_cell', cell), ('cell_offset', qS)):
        out[f'stress/in/{k}'] = v
    print('stress: E =', out['stress/E'], ' |stress|max =', np.abs(out['stress/stress']).max())

    # ---- hyperparameters.json as the reference writes it: StackNet.__dict_repr__ (stacknet.py:89-111) of nets built by
    # init_so3krates -- the schema So3kratesConfig.from_hyperparameters / to_hyperparameters must read and write
    import json
    md17 = {'energy': 'E', 'force': 'F', 'atomic_type': 'z', 'atomic_position': 'R', 'hirshfeld_volume': 'V_eff',
            'total_charge': 'Q', 'total_spin': 'S', 'partial_charge': 'q', 'idx_i': 'idx_i', 'idx_j': 'idx_j',
            'unit_cell': 'unit_cell', 'cell_offset': 'cell_offset', 'node_mask': 'node_mask', 'stress': 'stress',
            'pbc': 'pbc', 'displacement_vector': 'displacements', 'atomic_energy': 'atomic_energy'}
    Energy = sys.modules['mlff.nn.observable'].Energy
    hp = {}
    for tag, F_, L_, lk, ek in (('default', 132, 3, {}, {}),
                                ('optional', 64, 2, dict(allk, n_heads=2, degrees=[1, 2, 3, 4]),
                                 dict(zbl_repulsion=True, zbl_repulsion_shift=0.37))):
        hnet = so3krates.init_so3krates(dict(md17), F=F_, n_layer=L_, so3krates_layer_kwargs=lk or None,
                                        geometry_embed_kwargs={'degrees': lk['degrees'], 'mic': True} if lk else None,
                                        obs=[Energy(prop_keys=dict(md17), **ek)])
        hp[tag] = hnet.__dict_repr__()
    with open(os.path.join(HERE, 'reference_hyperparameters.json'), 'w', encoding='utf-8') as f:
        json.dump(hp, f, ensure_ascii=False, indent=1)
    print('wrote reference_hyperparameters.json')

    run_md_seam('md_seam_default', o.OracleConfig(F=36, n_layers=2), 5, {}, 1.3)
    zk = dict(zbl_repulsion=True, zbl_repulsion_shift=0.37)
    run_md_seam('md_seam_zbl', o.OracleConfig(F=36, n_layers=2, **zk), 6, zk, 0.7)
    np.savez_compressed(os.path.join(HERE, 'reference_model.npz'), **out)
    print('wrote reference_model.npz')


def _f64(tree, flat, cfg):
    """to_flax_tree stores float32; put the float64 values back leaf by leaf (same paths)."""
    from mlff_b200 import params as P
    for name, shape in P.param_shapes(cfg):
        path, vm = P._flax_path(name, cfg.residual_mlp_1)
        if not path:
            continue
        node = tree
        for k in path[:-1]:
            node = node[k]
        a = np.asarray(flat[name], np.float64).reshape(shape)
        node[path[-1]] = a[None] if vm else a
    return tree


if __name__ == '__main__':
    main()
