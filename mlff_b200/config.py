"""Hyper-parameters of the SO3krates model, mirroring the reference's constructor and JSON schema.

  So3krates(...)                       mlff/nn/representation/so3krates.py:13-67
  hyperparameters.json -> stack_net    mlff/nn/stacknet/stacknet.py:89-106, 135-145
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

md17_property_keys = {'energy': 'E', 'force': 'F', 'atomic_type': 'z', 'atomic_position': 'R', 'hirshfeld_volume': 'V_eff',
                      'total_charge': 'Q', 'total_spin': 'S', 'partial_charge': 'q', 'idx_i': 'idx_i', 'idx_j': 'idx_j',
                      'unit_cell': 'unit_cell', 'cell_offset': 'cell_offset', 'node_mask': 'node_mask',
                      'stress': 'stress', 'pbc': 'pbc', 'displacement_vector': 'displacements',
                      'atomic_energy': 'atomic_energy'}


@dataclass
class So3kratesConfig:
    F: int = 132
    n_layer: int = 6                      # python-API default (so3krates.py:16); the CLI default is 3
    degrees: Sequence[int] = (1, 2, 3)
    n_heads: int = 4
    n_rbf: int = 32
    r_cut: float = 5.0
    num_embeddings: int = 100
    mic: bool = False
    zbl_repulsion: bool = False           # Energy(zbl_repulsion=...), mlff/nn/observable/observable.py:39 (CLI --zbl_repulsion)
    zbl_repulsion_shift: float = 0.0      # observable.py:40
    # optional So3kratesLayer branches (so3krates_layer.py:30-42), set for every layer through
    # So3krates(so3krates_layer_kwargs=...) (so3krates.py:23-25, 38)
    layer_normalization: bool = False     # pre-LayerNorm before the blocks and before the interaction block
    final_layer: bool = False             # with layer_normalization: post-LayerNorm on the layer output
    residual_mlp_1: bool = False          # ResidualMLP after the first skip connection
    residual_mlp_2: bool = False          # ResidualMLP after the second skip connection
    # Energy(per_atom_scale, per_atom_shift): constants indexed by atomic number, baked into the observable by
    # init_stack_net from hyperparameters.json (observable.py:36-37, 49-58); None = 1 / 0
    per_atom_scale: Optional[Sequence[float]] = None
    per_atom_shift: Optional[Sequence[float]] = None
    prop_keys: Dict[str, str] = field(default_factory=lambda: dict(md17_property_keys))

    def validate(self):
        unsupported = []
        if self.F % self.n_heads or self.F % len(self.degrees):
            raise ValueError('F must be divisible by n_heads and by len(degrees) (equal_head_split, '
                             'so3krates_layer.py:611-615)')
        if any(l < 0 or l > 4 for l in self.degrees):
            raise NotImplementedError('Spherical harmonics are only defined up to order l = 4.')
        for k in ('per_atom_scale', 'per_atom_shift'):
            v = getattr(self, k)
            if v is not None and len(v) > self.num_embeddings:
                raise ValueError(f'{k} has {len(v)} entries for num_embeddings = {self.num_embeddings}')
        return unsupported

    @property
    def M(self) -> int:
        return sum(2 * int(l) + 1 for l in self.degrees)

    # ---- hyperparameters.json -----------------------------------------------------------------
    @classmethod
    def from_hyperparameters(cls, h: Dict) -> 'So3kratesConfig':
        """Parse the dict the reference writes with StackNet.to_json (stacknet.py:89-111)."""
        sn = h['stack_net'] if 'stack_net' in h else h
        geo = next(iter(sn['geometry_embeddings'][0].values()))
        feat = next(iter(sn['feature_embeddings'][0].values()))
        layers = [next(iter(x.values())) for x in sn['layers']]
        l0 = layers[0]
        checks = {
            'radial_basis_function': (geo.get('radial_basis_function', 'phys'), ('phys',)),
            'radial_cutoff_fn': (geo.get('radial_cutoff_fn', 'cosine_cutoff_fn'), ('cosine_cutoff_fn',)),
            'sphc_normalization': (geo.get('sphc_normalization', None), (None,)),
            'solid_harmonic': (geo.get('solid_harmonic', False), (False,)),
            'fb_filter': (l0.get('fb_filter', 'radial_spherical'), ('radial_spherical',)),
            'gb_filter': (l0.get('gb_filter', 'radial_spherical'), ('radial_spherical',)),
            'non_local_feature': (l0.get('non_local_feature', False), (False,)),
            'non_local_sphc': (l0.get('non_local_sphc', False), (False,)),
        }
        opt_keys = ('layer_normalization', 'final_layer', 'residual_mlp_1', 'residual_mlp_2')
        for k in opt_keys:      # So3krates() gives every layer the same keyword arguments; mixed stacks are not implemented
            if any(bool(l.get(k, False)) != bool(l0.get(k, False)) for l in layers):
                raise NotImplementedError(f'hyper-parameter {k} differs between layers: outside the hot path implemented '
                                          'by mlff_b200')
        for k, (v, ok) in checks.items():
            if v not in ok:
                raise NotImplementedError(f'hyper-parameter {k}={v!r} is outside the hot path implemented by mlff_b200')
        F = int(feat['features'])
        for l in layers:
            for blk in ('fb', 'gb'):
                if (list(l.get(f'{blk}_rad_filter_features', [F, F])) != [F, F]
                        or list(l.get(f'{blk}_sph_filter_features', [int(F / 4), F])) != [int(F / 4), F]):
                    raise NotImplementedError(f'{blk} filter widths: only the defaults [F,F] / [F/4,F] are implemented')
            for k, ok in (('fb_attention', 'conv_att'), ('gb_attention', 'conv_att'), ('chi_cut', None),
                          ('chi_cut_dynamic', False), ('sphc_normalization', False),
                          ('neighborhood_normalization', False), ('fast_attention_kwargs', None)):
                if l.get(k, ok) not in (ok, None if ok is False else ok):
                    raise NotImplementedError(f'layer hyper-parameter {k}={l.get(k)!r} is outside the hot path '
                                              'implemented by mlff_b200')
            if [int(x) for x in l.get('degrees', geo['degrees'])] != [int(x) for x in geo['degrees']]:
                raise NotImplementedError('layer degrees differ from the geometry embedding degrees')
            if int(l.get('n_heads', 4)) != int(l0.get('n_heads', 4)):
                raise NotImplementedError('n_heads differs between layers')
        if len(sn['geometry_embeddings']) != 1 or len(sn['feature_embeddings']) != 1:
            raise NotImplementedError('exactly one geometry embedding and one feature embedding are implemented')
        if geo.get('input_convention', 'positions') not in ('positions', 'displacements'):
            raise NotImplementedError(f"input_convention={geo.get('input_convention')!r}")
        if not geo.get('sphc', True):
            raise NotImplementedError('sphc=False is outside the hot path implemented by mlff_b200')
        if len(sn.get('observables', [])) > 1 or (sn.get('observables') and 'energy' not in sn['observables'][0]):
            raise NotImplementedError('only the Energy observable is implemented')
        obs = next(iter(sn['observables'][0].values())) if sn.get('observables') else {}
        if obs.get('output_convention', 'per_structure') not in ('per_structure', 'per_atom'):
            raise NotImplementedError(f"output_convention={obs.get('output_convention')!r}")
        scale_shift = {}
        for k in ('per_atom_scale', 'per_atom_shift'):
            v = obs.get(k, None)
            scale_shift[k] = None if v is None else tuple(float(x) for x in v)
        return cls(F=F, n_layer=len(layers), degrees=tuple(int(x) for x in geo['degrees']),
                   n_heads=int(l0.get('n_heads', 4)), n_rbf=int(geo['n_rbf']), r_cut=float(geo['r_cut']),
                   num_embeddings=int(feat.get('num_embeddings', 100)), mic=bool(geo.get('mic', False)),
                   zbl_repulsion=bool(obs.get('zbl_repulsion', False)),
                   zbl_repulsion_shift=float(obs.get('zbl_repulsion_shift', 0.0) or 0.0),
                   **{k: bool(l0.get(k, False)) for k in opt_keys}, **scale_shift,
                   prop_keys=dict(sn.get('prop_keys', md17_property_keys)))

    def to_hyperparameters(self) -> Dict:
        """The same nested dict StackNet.__dict_repr__ produces (stacknet.py:89-106)."""
        F = self.F
        layer = {'so3krates_layer': {
            'fb_filter': 'radial_spherical', 'fb_rad_filter_features': [F, F], 'fb_sph_filter_features': [int(F / 4), F],
            'fb_attention': 'conv_att', 'gb_filter': 'radial_spherical', 'gb_rad_filter_features': [F, F],
            'gb_sph_filter_features': [int(F / 4), F], 'gb_attention': 'conv_att', 'n_heads': self.n_heads,
            'residual_mlp_1': self.residual_mlp_1, 'residual_mlp_2': self.residual_mlp_2, 'non_local_sphc': False,
            'non_local_feature': False,
            'fast_attention_kwargs': None, 'chi_cut': None, 'chi_cut_dynamic': False, 'degrees': list(self.degrees),
            'parity': True, 'layer_normalization': self.layer_normalization, 'sphc_normalization': False,
            'neighborhood_normalization': False, 'final_layer': self.final_layer}}
        geo = {'geometry_embed': {'degrees': list(self.degrees), 'radial_basis_function': 'phys', 'n_rbf': self.n_rbf,
                                  'radial_cutoff_fn': 'cosine_cutoff_fn', 'r_cut': self.r_cut, 'sphc': True,
                                  'sphc_normalization': None, 'solid_harmonic': False, 'mic': self.mic,
                                  'input_convention': 'positions', 'prop_keys': self.prop_keys}}
        feat = {'atom_type_embed': {'num_embeddings': self.num_embeddings, 'features': F, 'prop_keys': self.prop_keys}}
        obs = {'energy': {'per_atom_scale': None if self.per_atom_scale is None else [float(x) for x in self.per_atom_scale],
                          'per_atom_shift': None if self.per_atom_shift is None else [float(x) for x in self.per_atom_shift],
                          'num_embeddings': self.num_embeddings,
                          'output_convention': 'per_structure', 'zbl_repulsion': self.zbl_repulsion,
                          'zbl_repulsion_shift': self.zbl_repulsion_shift,
                          'prop_keys': self.prop_keys}}
        return {'stack_net': {'geometry_embeddings': [geo], 'feature_embeddings': [feat],
                              'layers': [copy.deepcopy(layer) for _ in range(self.n_layer)], 'observables': [obs],
                              'prop_keys': self.prop_keys, 'n_layers': self.n_layer}}
