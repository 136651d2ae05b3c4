"""ctypes declarations for the C ABI in ``include/so3k.h``.

``So3kLib`` only knows about raw addresses (Python ints): device pointers of torch CUDA tensors in
the product, host pointers of numpy arrays when the test-suite drives the CPU-emulated build of the
same kernels.  Nothing here imports torch.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

SO3K_MAX_DEGREES = 8

SO3K_OK = 0
ERR_NAMES = {0: 'SO3K_OK', 1: 'SO3K_ERR_CONFIG', 2: 'SO3K_ERR_ARG', 3: 'SO3K_ERR_WORKSPACE',
             4: 'SO3K_ERR_CUDA', 5: 'SO3K_ERR_NOPARAMS'}
STATUS_OVERFLOW = 1
STATUS_ASYMMETRIC = 2
STATUS_BAD_Z = 4
STATUS_TC_TIMEOUT = 8
FLAG_TENSOR = 1
FLAG_KEEP_FILTERS = 2    # with FLAG_TENSOR: keep the forward's filters for the backward (memory for time)
FLAG_ZBL = 4             # Energy(zbl_repulsion=True): ZBL pair repulsion, ten extra scalars at the end of the blob
FLAG_LAYER_NORM = 8      # So3kratesLayer(layer_normalization=True)
FLAG_RESIDUAL_MLP_1 = 16 # So3kratesLayer(residual_mlp_1=True)
FLAG_RESIDUAL_MLP_2 = 32 # So3kratesLayer(residual_mlp_2=True)
FLAG_FINAL_LAYER = 64    # So3kratesLayer(final_layer=True): post LayerNorm (with FLAG_LAYER_NORM)
NL_ASE = 0
NL_DENSE = 1
NL_MIC = 2

# every symbol include/so3k.h declares (checked by tests/test_capi.py)
EXPORTS = ['so3k_create', 'so3k_destroy', 'so3k_version', 'so3k_param_count', 'so3k_param_lookup',
           'so3k_load_params', 'so3k_workspace_bytes', 'so3k_energy_forces', 'so3k_potential_displacements',
           'so3k_featurise', 'so3k_neighbor_list_workspace_bytes', 'so3k_neighbor_list', 'so3k_launch_count',
           'so3k_profile_categories', 'so3k_profile_enable', 'so3k_profile_collect', 'so3k_tc_selftest',
           'so3k_dd_begin', 'so3k_dd_stage', 'so3k_dd_stage_owned', 'so3k_dd_buffer', 'so3k_dd_pack', 'so3k_dd_unpack', 'so3k_stress', 'so3k_train_workspace_bytes',
           'so3k_energy_forces_vjp', 'so3k_loss_sums', 'so3k_loss_cotangents', 'so3k_adam_step']


class So3kConfigC(C.Structure):
    _fields_ = [('F', C.c_int32), ('n_rbf', C.c_int32), ('n_layers', C.c_int32), ('n_heads', C.c_int32),
                ('n_degrees', C.c_int32), ('degrees', C.c_int32 * SO3K_MAX_DEGREES), ('r_cut', C.c_float),
                ('num_embeddings', C.c_int32), ('flags', C.c_int32), ('zbl_shift', C.c_float)]


class So3kError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc != SO3K_OK:
        raise So3kError(f'{what} failed: {ERR_NAMES.get(rc, rc)}')


class So3kLib:
    """Typed view of a loaded libso3k shared object."""

    def __init__(self, path: str):
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        vp, i32, i64, f32, sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
        d.so3k_create.argtypes = [C.POINTER(So3kConfigC), C.POINTER(vp)]
        d.so3k_create.restype = C.c_int
        d.so3k_destroy.argtypes = [vp]
        d.so3k_destroy.restype = C.c_int
        d.so3k_version.argtypes = []
        d.so3k_version.restype = C.c_char_p
        d.so3k_param_count.argtypes = [vp]
        d.so3k_param_count.restype = sz
        d.so3k_param_lookup.argtypes = [vp, C.c_char_p, C.POINTER(sz), C.POINTER(sz)]
        d.so3k_param_lookup.restype = C.c_int
        d.so3k_load_params.argtypes = [vp, vp, sz, vp]
        d.so3k_load_params.restype = C.c_int
        d.so3k_workspace_bytes.argtypes = [vp, i64, i64]
        d.so3k_workspace_bytes.restype = sz
        d.so3k_energy_forces.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp]
        d.so3k_energy_forces.restype = C.c_int
        d.so3k_stress.argtypes = [vp, i32, i32, i32, vp, sz, vp, vp]
        d.so3k_stress.restype = C.c_int
        d.so3k_potential_displacements.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, f32, vp, vp, vp, vp, sz, vp, vp]
        d.so3k_potential_displacements.restype = C.c_int
        d.so3k_featurise.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        d.so3k_featurise.restype = C.c_int
        d.so3k_neighbor_list_workspace_bytes.argtypes = [i64, i64]
        d.so3k_neighbor_list_workspace_bytes.restype = sz
        d.so3k_neighbor_list.argtypes = [i32, i64, vp, vp, vp, vp, C.c_double, i64, vp, vp, vp, vp, vp, sz, vp, vp]
        d.so3k_neighbor_list.restype = C.c_int
        d.so3k_launch_count.argtypes = []
        d.so3k_launch_count.restype = C.c_uint64
        d.so3k_dd_begin.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp]
        d.so3k_dd_begin.restype = C.c_int
        d.so3k_dd_stage.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp, vp]
        d.so3k_dd_stage.restype = C.c_int
        d.so3k_dd_stage_owned.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp]
        d.so3k_dd_stage_owned.restype = C.c_int
        d.so3k_dd_pack.argtypes = [vp, i32, i32, i32, i32, vp, i32, vp, vp, vp]
        d.so3k_dd_pack.restype = C.c_int
        d.so3k_dd_unpack.argtypes = [vp, i32, i32, i32, i32, i32, vp, vp, vp]
        d.so3k_dd_unpack.restype = C.c_int
        d.so3k_dd_buffer.argtypes = [vp, i32, i32, i32, i32, C.POINTER(sz), C.POINTER(sz)]
        d.so3k_dd_buffer.restype = C.c_int
        d.so3k_tc_selftest.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp]
        d.so3k_tc_selftest.restype = C.c_int
        d.so3k_profile_categories.argtypes = []
        d.so3k_profile_categories.restype = C.c_char_p
        d.so3k_profile_enable.argtypes = [vp, i32]
        d.so3k_profile_enable.restype = C.c_int
        d.so3k_profile_collect.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_int64), i32, i32]
        d.so3k_profile_collect.restype = C.c_int
        d.so3k_train_workspace_bytes.argtypes = [vp, i64, i64]
        d.so3k_train_workspace_bytes.restype = sz
        d.so3k_energy_forces_vjp.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, sz, vp, vp]
        d.so3k_energy_forces_vjp.restype = C.c_int
        d.so3k_loss_sums.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp]
        d.so3k_loss_sums.restype = C.c_int
        d.so3k_loss_cotangents.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, f32, f32, vp, vp, vp, vp]
        d.so3k_loss_cotangents.restype = C.c_int
        d.so3k_adam_step.argtypes = [sz, vp, vp, vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, f32, vp, vp, vp]
        d.so3k_adam_step.restype = C.c_int

    # ---- handle -----------------------------------------------------------------------------
    def create(self, F: int, n_rbf: int, n_layers: int, n_heads: int, degrees: Sequence[int], r_cut: float,
               num_embeddings: int = 100, flags: int = 0, zbl_shift: float = 0.0) -> int:
        cfg = So3kConfigC()
        cfg.F, cfg.n_rbf, cfg.n_layers, cfg.n_heads = F, n_rbf, n_layers, n_heads
        cfg.n_degrees = len(degrees)
        if len(degrees) > SO3K_MAX_DEGREES:
            raise So3kError('too many degrees')
        for k, l in enumerate(degrees):
            cfg.degrees[k] = int(l)
        cfg.r_cut = float(r_cut)
        cfg.num_embeddings = num_embeddings
        cfg.flags = flags
        cfg.zbl_shift = float(zbl_shift)
        h = C.c_void_p()
        _check(self.dll.so3k_create(C.byref(cfg), C.byref(h)), 'so3k_create')
        return h.value

    def destroy(self, h: int):
        self.dll.so3k_destroy(h)

    def version(self) -> str:
        return self.dll.so3k_version().decode()

    def param_count(self, h: int) -> int:
        return int(self.dll.so3k_param_count(h))

    def param_lookup(self, h: int, name: str):
        off, cnt = C.c_size_t(), C.c_size_t()
        rc = self.dll.so3k_param_lookup(h, name.encode(), C.byref(off), C.byref(cnt))
        if rc != SO3K_OK:
            raise KeyError(name)
        return int(off.value), int(cnt.value)

    def load_params(self, h: int, blob_ptr: int, n: int, stream: int = 0):
        _check(self.dll.so3k_load_params(h, blob_ptr, n, stream), 'so3k_load_params')

    def workspace_bytes(self, h: int, n_atoms: int, n_edge_slots: int) -> int:
        return int(self.dll.so3k_workspace_bytes(h, n_atoms, n_edge_slots))

    # ---- compute ----------------------------------------------------------------------------
    def energy_forces(self, h, B, n, P, R, z, idx_i, idx_j, cell, cell_offset, E, F, e_atom, ws, ws_bytes,
                      status, stream=0):
        _check(self.dll.so3k_energy_forces(h, B, n, P, R, z, idx_i, idx_j, cell, cell_offset, E, F, e_atom, ws,
                                           ws_bytes, status, stream), 'so3k_energy_forces')

    def stress(self, h, B, n, P, ws, ws_bytes, stress, stream=0):
        _check(self.dll.so3k_stress(h, B, n, P, ws, ws_bytes, stress, stream), 'so3k_stress')

    def potential_displacements(self, h, N, P, edges, z, centers, others, mask, energy_scale, e_atom, dE_dedges,
                                forces, ws, ws_bytes, status, stream=0):
        _check(self.dll.so3k_potential_displacements(h, N, P, edges, z, centers, others, mask, energy_scale, e_atom,
                                                     dE_dedges, forces, ws, ws_bytes, status, stream),
               'so3k_potential_displacements')

    def featurise(self, h, n, P, R, idx_i, idx_j, cell, cell_offset, disp, rbf, phi, d, unit, sph, stream=0):
        _check(self.dll.so3k_featurise(h, n, P, R, idx_i, idx_j, cell, cell_offset, disp, rbf, phi, d, unit, sph,
                                       stream), 'so3k_featurise')

    def neighbor_list_workspace_bytes(self, N: int, capacity: int) -> int:
        return int(self.dll.so3k_neighbor_list_workspace_bytes(N, capacity))

    def neighbor_list(self, mode, N, positions, z, cell, pbc, cutoff, capacity, idx_i, idx_j, shifts, count, ws,
                      ws_bytes, status, stream=0):
        _check(self.dll.so3k_neighbor_list(mode, N, positions, z, cell, pbc, float(cutoff), capacity, idx_i, idx_j,
                                           shifts, count, ws, ws_bytes, status, stream), 'so3k_neighbor_list')

    # ---- training seam ------------------------------------------------------------------------
    def train_workspace_bytes(self, h: int, n_atoms: int, n_edge_slots: int) -> int:
        return int(self.dll.so3k_train_workspace_bytes(h, n_atoms, n_edge_slots))

    def energy_forces_vjp(self, h, B, n, P, R, z, idx_i, idx_j, cell, cell_offset, E_bar, F_bar, grad, E_out, ws,
                          ws_bytes, status, stream=0, S_bar=None):
        _check(self.dll.so3k_energy_forces_vjp(h, B, n, P, R, z, idx_i, idx_j, cell, cell_offset, E_bar, F_bar, S_bar,
                                               grad, E_out, ws, ws_bytes, status, stream), 'so3k_energy_forces_vjp')

    def loss_sums(self, B, n, E, F, E_true, F_true, z, sums, stream=0):
        _check(self.dll.so3k_loss_sums(B, n, E, F, E_true, F_true, z, sums, stream), 'so3k_loss_sums')

    def loss_cotangents(self, B, n, E, F, E_true, F_true, z, sums, w_energy, w_force, E_bar, F_bar, loss, stream=0):
        _check(self.dll.so3k_loss_cotangents(B, n, E, F, E_true, F_true, z, sums, float(w_energy), float(w_force),
                                             E_bar, F_bar, loss, stream), 'so3k_loss_cotangents')

    def adam_step(self, n, params, grad, m, v, frozen, decay_mask, step, lr, b1, b2, eps, eps_root, weight_decay,
                  clip_norm, grad_sumsq, stream=0, step_dev=None):
        _check(self.dll.so3k_adam_step(n, params, grad, m, v, frozen, decay_mask, int(step), float(lr), float(b1),
                                       float(b2), float(eps), float(eps_root), float(weight_decay), float(clip_norm),
                                       grad_sumsq, step_dev, stream), 'so3k_adam_step')

    # ---- stepwise / domain decomposition --------------------------------------------------------
    def dd_begin(self, h, N, P, R, z, idx_i, idx_j, cell, cell_offset, ws, ws_bytes, status, stream=0):
        _check(self.dll.so3k_dd_begin(h, N, P, R, z, idx_i, idx_j, cell, cell_offset, ws, ws_bytes, status, stream),
               'so3k_dd_begin')

    def dd_stage(self, h, stage, t, N, P, z, out, ws, status, stream=0):
        _check(self.dll.so3k_dd_stage(h, stage, t, N, P, z, out, ws, status, stream), 'so3k_dd_stage')

    def dd_stage_owned(self, h, stage, t, N, n_owned, P, z, out, ws, status, stream=0):
        _check(self.dll.so3k_dd_stage_owned(h, stage, t, N, n_owned, P, z, out, ws, status, stream), 'so3k_dd_stage_owned')

    def dd_pack(self, h, pair, t, N, P, send_idx, n_send, out, ws, stream=0):
        _check(self.dll.so3k_dd_pack(h, pair, t, N, P, send_idx, n_send, out, ws, stream), 'so3k_dd_pack')

    def dd_unpack(self, h, pair, t, N, n_owned, P, inp, ws, stream=0):
        _check(self.dll.so3k_dd_unpack(h, pair, t, N, n_owned, P, inp, ws, stream), 'so3k_dd_unpack')

    def dd_buffer(self, h, which, t, N, P):
        off, rf = C.c_size_t(), C.c_size_t()
        _check(self.dll.so3k_dd_buffer(h, which, t, N, P, C.byref(off), C.byref(rf)), 'so3k_dd_buffer')
        return int(off.value), int(rf.value)

    def tc_selftest(self, N, K, variant, passes, A, B, D, stream=0):
        _check(self.dll.so3k_tc_selftest(N, K, variant, passes, A, B, D, stream), 'so3k_tc_selftest')

    def launch_count(self) -> int:
        return int(self.dll.so3k_launch_count())

    def profile_enable(self, h: int, on: bool):
        _check(self.dll.so3k_profile_enable(h, 1 if on else 0), 'so3k_profile_enable')

    def profile_collect(self, h: int, reset: bool = True):
        names = self.dll.so3k_profile_categories().decode().split(',')
        ms = (C.c_double * len(names))()
        cnt = (C.c_int64 * len(names))()
        _check(self.dll.so3k_profile_collect(h, ms, cnt, len(names), 1 if reset else 0), 'so3k_profile_collect')
        return {n: (float(ms[k]), int(cnt[k])) for k, n in enumerate(names)}
