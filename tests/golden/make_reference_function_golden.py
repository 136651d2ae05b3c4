"""Golden vectors from the REFERENCE'S OWN SOURCE for the stateless pieces of the path.

jax / flax are not installed in this image, so the reference cannot be run as a whole.  Its geometry embedding,
basis functions, cutoff, masks and l=0 contraction, however, are plain array code: this script executes those
reference files UNMODIFIED, where they lie under /root/reference, with thin stand-ins for the three imports they need --
`jax.numpy` (numpy, float64), `jax.jit / jax.vmap / jax.ops.segment_sum` (Python loops) and `flax.linen.Module`
(tests/golden/ref_shim.py; no parameters are involved in these modules).  Nothing is copied: the files are loaded by path.

    python tests/golden/make_reference_function_golden.py        (build container only; needs /root/reference)

Writes tests/golden/reference_functions.npz:
  GeometryEmbed.__call__   mlff/nn/embed/embed.py:66-169        positions (+ mic cell offsets) and displacements
                           conventions, padded pairs: r_ij, d_ij, rbf_ij, phi_r_cut, unit_r_ij, sph_ij (l = 0..4)
     -> PhysNetBasis       mlff/basis_function/radial.py:140-166
     -> cosine_cutoff_fn   mlff/cutoff_function/radial.py:39-51
     -> init_sph_fn        mlff/basis_function/spherical.py:6-152
     -> add_cell_offsets   mlff/cutoff_function/pbc.py:8-18
     -> safe_mask / safe_scale   mlff/masking/mask.py:5-54
  make_l0_contraction_fn   mlff/sph_ops/contract.py:162-190     degrees [1,2,3], [1,2,2,3], [0,1,2,3,4]
  get_indices              mlff/indexing/indices.py:15-84       padded batch of ethanol frames, two cutoffs, a pair at
                           exactly r_cut, an isolated atom, padding atoms
tests/test_oracle.py::test_oracle_matches_reference_source_functions compares the oracle with these arrays.
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim as rs  # noqa: E402

REF = rs.REF
load = rs.load


def install_stand_ins():
    rs.install()
    for name in ('mlff.geometric', 'mlff.indexing', 'mlff.padding', 'mlff.utils', 'ase', 'ase.neighborlist'):
        m = types.ModuleType(name)
        m.__path__ = []
        sys.modules[name] = m
    sys.modules['ase.neighborlist'].primitive_neighbor_list = None      # ASE is absent; only get_pbc_neighbors needs it
    sys.modules['mlff.utils'].PrimitiveNeighbors = None


def main():
    install_stand_ins()
    load('mlff.masking.mask')
    load('mlff.properties.property_names')
    load('mlff.basis_function.spherical')
    load('mlff.basis_function.radial')
    pbc = load('mlff.cutoff_function.pbc')
    rad = load('mlff.cutoff_function.radial')
    cf = sys.modules['mlff.cutoff_function']             # what mlff/cutoff_function/__init__.py re-exports
    cf.add_cell_offsets, cf.get_cutoff_fn, cf.pbc_diff = pbc.add_cell_offsets, rad.get_cutoff_fn, pbc.pbc_diff
    load('mlff.nn.base.sub_module')
    embed = load('mlff.nn.embed.embed')
    contract = load('mlff.sph_ops.contract')

    out = {}
    rng = np.random.default_rng(0)
    degrees = [0, 1, 2, 3, 4]
    prop_keys = {'atomic_position': 'R', 'atomic_type': 'z', 'unit_cell': 'unit_cell', 'cell_offset': 'cell_offset',
                 'displacement_vector': 'displacements'}

    def geometry(tag, convention, mic, inputs):
        ge = embed.GeometryEmbed(prop_keys=prop_keys, degrees=degrees, radial_basis_function='phys', n_rbf=32,
                                 radial_cutoff_fn='cosine_cutoff_fn', r_cut=5.0, sphc=True, mic=mic,
                                 input_convention=convention)
        g = ge.apply({'params': {}}, inputs)
        for k in ('r_ij', 'd_ij', 'rbf_ij', 'phi_r_cut', 'unit_r_ij', 'sph_ij', 'chi'):
            out[f'{tag}/{k}'] = np.asarray(g[k], np.float64)
        for k, v in inputs.items():
            out[f'{tag}/in/{k}'] = np.asarray(v)

    # molecule: ethanol frame 0, all pairs within 5 A, three padding pairs, one pair at distance 0 < d and one beyond r_cut
    eth = np.load(os.path.join(REF, 'tests/test_data/ethanol.npz'))
    R = eth['R'][0].astype(np.float64)
    z = eth['z'].astype(np.int64)
    n = len(z)
    ii, jj = np.nonzero(~np.eye(n, dtype=bool))
    R_far = np.concatenate([R, R[:1] + np.array([[6.5, 0.0, 0.0]])])           # an atom outside the cutoff of atom 0
    ii = np.concatenate([ii, [0, n, -1, -1, -1]])
    jj = np.concatenate([jj, [n, 0, -1, -1, -1]])
    pm = (ii != -1).astype(np.float64)
    zz = np.concatenate([z, [1]])
    geometry('molecule', 'positions', False,
             {'R': R_far, 'z': zz, 'idx_i': ii, 'idx_j': jj, 'pair_mask': pm, 'point_mask': (zz != 0).astype(np.float64)})

    # periodic: frame 0 of the reference's solid fixture, random pairs with cell offsets in {-1,0,1}^3
    sol = np.load(os.path.join(REF, 'tests/test_data/test_solid.npz'))
    Rs, cell, zs = sol['R'][0].astype(np.float64), sol['unit_cell'][0].astype(np.float64), sol['z'][0].astype(np.int64)
    P = 400
    pi, pj = rng.integers(0, len(zs), P), rng.integers(0, len(zs), P)
    S = rng.integers(-1, 2, (P, 3))
    keep = (pi != pj) | (np.abs(S).sum(-1) > 0)
    pi, pj, S = pi[keep], pj[keep], S[keep]
    pi = np.concatenate([pi, [-1, -1]])
    pj = np.concatenate([pj, [-1, -1]])
    S = np.concatenate([S, np.zeros((2, 3), S.dtype)])
    geometry('periodic', 'positions', True,
             {'R': Rs, 'z': zs, 'idx_i': pi, 'idx_j': pj, 'pair_mask': (pi != -1).astype(np.float64),
              'point_mask': (zs != 0).astype(np.float64), 'unit_cell': cell, 'cell_offset': S})

    # MD seam: displacement vectors in, including a zero vector (unit vector placeholder) and masked slots
    disp = rng.normal(0.0, 2.0, (64, 3))
    disp[5] = 0.0
    di, dj = rng.integers(0, 9, 64), rng.integers(0, 9, 64)
    dm = np.ones(64)
    dm[60:] = 0.0
    di[60:] = -1
    dj[60:] = -1
    geometry('displacements', 'displacements', False,
             {'displacements': disp, 'z': z, 'idx_i': di, 'idx_j': dj, 'pair_mask': dm,
              'point_mask': (z != 0).astype(np.float64)})

    # l = 0 contraction
    for tag, degs in (('123', [1, 2, 3]), ('1223', [1, 2, 2, 3]), ('01234', [0, 1, 2, 3, 4])):
        fn = contract.make_l0_contraction_fn(degs, dtype=np.float64)
        chi = rng.normal(size=(7, sum(2 * l + 1 for l in degs)))
        out[f'l0/{tag}/chi'] = chi
        out[f'l0/{tag}/degrees'] = np.array(degs)
        out[f'l0/{tag}/out'] = np.asarray(fn(chi), np.float64)

    # training-set neighbour lists: get_indices (mlff/indexing/indices.py:15-84) on padded ethanol frames
    load('mlff.geometric.metric')
    indices = load('mlff.indexing.indices')
    Rb = np.concatenate([eth['R'][:6].astype(np.float64), np.zeros((6, 2, 3))], axis=1)       # two padding atoms (z = 0)
    zb = np.tile(np.concatenate([z, [0, 0]])[None], (6, 1))
    Rb[3, 1] = Rb[3, 0] + np.array([4.0, 3.0, 0.0])                           # a pair at exactly r_cut (d <= r_cut counts)
    Rb[4, 2] = Rb[4, 0] + 30.0                                                 # an atom with no neighbours at all
    for rc in (5.0, 2.0):
        nlr = indices.get_indices(Rb, zb, rc)
        out[f'get_indices/{rc}/idx_i'], out[f'get_indices/{rc}/idx_j'] = nlr['idx_i'], nlr['idx_j']
    out['get_indices/R'], out['get_indices/z'] = Rb, zb

    # batch padding helpers (mlff/padding/padding.py:47-108, mlff/indexing/indices.py:87-108)
    padding = load('mlff.padding.padding')
    zz_ = np.arange(1, 7).reshape(2, 3)
    RR_ = rng.normal(size=(2, 3, 3))
    ij_ = rng.integers(0, 3, (2, 4))
    out['pad/in/z'], out['pad/in/R'], out['pad/in/idx'] = zz_, RR_, ij_
    out['pad/atomic_types'] = padding.pad_atomic_types(zz_, 5)
    out['pad/coordinates'] = padding.pad_coordinates(RR_, 5)
    out['pad/forces'] = padding.pad_forces(RR_, 4)
    out['pad/per_atom_quantity'] = padding.pad_per_atom_quantity(RR_[..., :1], 6)
    out['pad/indices_i'], out['pad/indices_j'] = padding.pad_indices(ij_, ij_[:, ::-1], 7)
    out['pad/index_list'] = indices.pad_index_list(ij_[0], 6)
    out['pad/shift'] = indices.pad_shift(ij_[:, :3], 5)

    np.savez_compressed(os.path.join(HERE, 'reference_functions.npz'), **out)
    print('wrote', os.path.join(HERE, 'reference_functions.npz'), len(out), 'arrays')


if __name__ == '__main__':
    main()
