"""Generate the committed golden fixtures.  Run ONCE in the build container (needs /root/reference):

    python tests/golden/make_golden.py

Writes (all small):
  l0_cg_diag.npy      diag of CG(l,l->0) for l = 0..4 read from the reference's mlff/sph_ops/cgmatrix.npz
                      (what make_l0_contraction_fn uses, contract.py:162-177)
  ethanol_32.npz      R (32,9,3), z (9,)  = frames 0..31 of the reference fixture tests/test_data/ethanol.npz
  solid_0.npz         R (110,3), z (110,), unit_cell (3,3) = frame 0 of tests/test_data/test_solid.npz
  oracle_*.npz        float64 oracle outputs (E, F, ...) for fixed-seed weights on those inputs, so that the oracle
                      itself is regression-pinned and GPU tests can compare without re-deriving anything.
These are ORACLE outputs (regression pins).  Outputs of the reference's own code are produced by the two sibling
scripts make_reference_function_golden.py / make_reference_model_golden.py (reference source executed on numpy /
apply-only flax stand-ins, since jax / flax / ase are not installed here).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import so3k_oracle as o  # noqa: E402

REF = '/root/reference'


def main():
    cg = np.load(os.path.join(REF, 'mlff/sph_ops/cgmatrix.npz'))['cg']
    diag = np.diagonal(cg[0:1, :25, :25], axis1=1, axis2=2)[0]
    np.save(os.path.join(HERE, 'l0_cg_diag.npy'), diag)

    eth = np.load(os.path.join(REF, 'tests/test_data/ethanol.npz'))
    np.savez_compressed(os.path.join(HERE, 'ethanol_32.npz'), R=eth['R'][:32], z=eth['z'].astype(np.int64))
    sol = np.load(os.path.join(REF, 'tests/test_data/test_solid.npz'))
    np.savez_compressed(os.path.join(HERE, 'solid_0.npz'), R=sol['R'][0], z=sol['z'][0], unit_cell=sol['unit_cell'][0])

    # ---- oracle outputs ------------------------------------------------------------------------
    R = eth['R'][:4]
    z = np.tile(eth['z'][None].astype(np.int64), (4, 1))
    nl = o.get_indices(R, z, 5.0)
    cfg = o.OracleConfig()
    for tag, es in (('default', 1.0), ('amplified', 4.0)):
        p = o.init_params(cfg, seed=0, embed_scale=es)
        E, F = o.energy_forces_batch(cfg, p, R, z, nl['idx_i'], nl['idx_j'])
        np.savez_compressed(os.path.join(HERE, f'oracle_ethanol_{tag}.npz'), E=E.numpy(), F=F.numpy(),
                            idx_i=nl['idx_i'], idx_j=nl['idx_j'], seed=0, embed_scale=es)
    # periodic solid, all-image neighbour list (cell 9.05 A < 2 r_cut -> several images per pair)
    pos, cell, zz = sol['R'][0], sol['unit_cell'][0], sol['z'][0]
    ii, jj, S = o.primitive_neighbor_list(pos, cell, [True, True, True], 5.0)
    p = o.init_params(cfg, seed=1, embed_scale=4.0)
    E, F, ea = o.energy_forces(cfg, p, pos, zz, ii, jj, cell=cell, cell_offset=S)
    np.savez_compressed(os.path.join(HERE, 'oracle_solid_pbc.npz'), E=E.numpy(), F=F.numpy(), e_atom=ea.numpy(),
                        idx_i=ii, idx_j=jj, shifts=S, seed=1, embed_scale=4.0)
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
