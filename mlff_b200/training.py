"""Training step on the device: the host-side mirror of the reference's training seam.

Reference interfaces mirrored here (paths relative to the reference checkout):
  * ``Optimizer`` / ``optimizer``            mlff/training/optimizer.py:12-97  (clip -> adam -> decayed weights -> -lr)
  * ``get_loss_fn``                          mlff/training/loss.py:42-73       (weighted, masked, NaN-safe MSE of E and F)
  * ``train_step_fn`` / ``TrainState``       mlff/training/run.py:34-50        (value_and_grad + apply_gradients,
                                                                               train_metrics incl. 'gradients_norm')
  * data parallelism: the reference vmaps the loss over the batch (cAPI/mlff_train_so3krates.py:381) and has no
    multi-device path; here each rank owns B / world structures and the ranks all-reduce the four loss sums (so that
    every rank divides by the GLOBAL denominators of scaled_safe_masked_mse_loss) and the parameter-gradient blob.

Everything numeric happens in libso3k (so3k_energy_forces, so3k_loss_*, so3k_energy_forces_vjp, so3k_adam_step);
PyTorch provides device memory, the stream and ``torch.distributed`` (NCCL on GPUs, gloo in the CPU test-suite, where
the same code drives the CPU-emulated build of the kernels through an engine stand-in).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, Optional, Tuple

import numpy as np
import torch


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


@dataclass
class Optimizer:
    """mlff/training/optimizer.py:12-62 (same field names and defaults).  ``get(learning_rate)`` returns the settings
    the device update needs; the scheduled decay (`transition_steps`, `decay_rate`) raises like the reference does."""
    b1: float = 0.9
    b2: float = 0.999
    eps: float = 1e-8
    eps_root: float = 0.0
    transition_steps: Optional[int] = None
    decay_rate: Optional[float] = None
    weight_decay: Optional[float] = None
    clip_by_global_norm: Optional[float] = None

    def get(self, learning_rate: float) -> 'Optimizer':
        if self.transition_steps is not None and self.decay_rate is not None:
            # optimizer.py:79-81
            raise DeprecationWarning('Scheduled LR decay has been moved to the train_state class. No LR decay is used.')
        self.learning_rate = float(learning_rate)
        return self

    def __dict_repr__(self):
        return {'optimizer': {'learning_rate': getattr(self, 'learning_rate', None),
                              'transition_steps': self.transition_steps, 'decay_rate': self.decay_rate,
                              'weight_decay': self.weight_decay, 'clip_by_global_norm': self.clip_by_global_norm}}


def lr_decay_exponential(step: int, transition_steps: int, decay_factor: float) -> float:
    """mlff/training/lr_decay/lr_decay.py:1-2."""
    return decay_factor ** (step / transition_steps)


def lr_decay_on_plateau(plateau_count: int, plateau_length: int, patience: int, decay_factor: float) -> Tuple[float, int, int]:
    """mlff/training/lr_decay/lr_decay.py:5-22: (factor of the n-th plateau, plateau length carried over, plateaus completed
    by this length)."""
    return decay_factor ** plateau_count, plateau_length % patience, plateau_length // patience


class TrainState:
    """step counter + device parameter blob + Adam moments (CustomTrainState of mlff/training/train_state.py:17-127).  The
    blob IS the engine's parameter blob: apply_gradients updates it in place and re-loads it into the handle.
    As in the reference the update of a step is multiplied by lr_decay_fn(step) and by the plateau factor of
    reduce_lr_on_plateau_fn(plateau_count, plateau_length) (train_state.py:75-83; defaults: no decay), `improved(x)` feeds
    the plateau counter (:104-110), and with polyak_step_size `valid_params` = step_size * new + (1 - step_size) * previous
    parameters (:85-90, optax.incremental_update) -- otherwise `valid_params` is `params`."""

    def __init__(self, engine, tx: Optimizer, lr_decay_fn: Optional[Callable] = None,
                 reduce_lr_on_plateau_fn: Optional[Callable] = None, polyak_step_size: Optional[float] = None):
        if engine._blob is None:
            raise RuntimeError('load parameters into the engine before creating a TrainState')
        self.engine, self.tx, self.step = engine, tx, 0
        self.lr_decay_fn, self.reduce_lr_on_plateau_fn = lr_decay_fn, reduce_lr_on_plateau_fn
        self.polyak_step_size = polyak_step_size
        self.plateau_length, self.plateau_count = 0, 0
        self.params = engine._blob
        self.valid_params = self.params if polyak_step_size is None else self.params.clone()
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        lib, h = engine.lib, engine.h
        n = self.params.numel()
        decay = np.ones(n, np.uint8)
        frozen = np.zeros(n, np.uint8)
        for name in engine_param_names(engine):
            off, cnt = lib.param_lookup(h, name)
            leaf = name.rsplit('/', 1)[-1]
            if leaf == 'bias' or name.startswith('energy/per_atom_'):
                decay[off:off + cnt] = 0          # flattened_traversal(path[-1] != 'bias'), optimizer.py:40
            if name.startswith('energy/per_atom_'):
                frozen[off:off + cnt] = 1         # constants of the Energy module, not flax parameters
        self.decay_mask = torch.from_numpy(decay).to(engine.device)
        self.frozen = torch.from_numpy(frozen).to(engine.device)
        self.grad_sumsq = torch.zeros(1, dtype=torch.float64, device=engine.device)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=engine.device)      # the update count, device resident

    @property
    def scheduled(self) -> bool:
        """True when a step depends on host-side schedule state (such a step cannot be replayed from a captured graph)."""
        return self.lr_decay_fn is not None or self.reduce_lr_on_plateau_fn is not None or self.polyak_step_size is not None

    def improved(self, x: bool) -> 'TrainState':
        """train_state.py:104-110: an epoch without improvement lengthens the current plateau."""
        self.plateau_length = 0 if x else self.plateau_length + 1
        return self

    def apply_gradients(self, grads: torch.Tensor) -> 'TrainState':
        """One optimiser update in place + re-load into the handle.  The update count lives on the device (incremented by
        the call), so the same enqueued work can be replayed from a CUDA graph; `self.step` mirrors it on the host."""
        e, tx = self.engine, self.tx
        factor = 1.0
        if self.reduce_lr_on_plateau_fn is not None:
            cd, self.plateau_length, inc = self.reduce_lr_on_plateau_fn(plateau_count=self.plateau_count,
                                                                        plateau_length=self.plateau_length)
            self.plateau_count += int(inc)
            factor *= float(cd)
        if self.lr_decay_fn is not None:
            factor *= float(self.lr_decay_fn(self.step))         # the counter BEFORE this update, as in the reference
        previous = self.params.clone() if self.polyak_step_size is not None else None
        self.step += 1
        wd = 0.0 if tx.weight_decay is None else tx.weight_decay
        clip = 0.0 if tx.clip_by_global_norm is None else tx.clip_by_global_norm
        e.lib.adam_step(self.params.numel(), _p(self.params), _p(grads), _p(self.m), _p(self.v), _p(self.frozen),
                        _p(self.decay_mask) if tx.weight_decay is not None else None, self.step, tx.learning_rate * factor,
                        tx.b1, tx.b2, tx.eps, tx.eps_root, wd, clip, _p(self.grad_sumsq), e._stream(),
                        step_dev=_p(self.step_dev))
        e.lib.load_params(e.h, _p(self.params), self.params.numel(), e._stream())
        if previous is not None:
            torch.lerp(previous, self.params, float(self.polyak_step_size), out=self.valid_params)
        return self


def engine_param_names(engine):
    from .params import param_shapes
    return [name for name, _ in param_shapes(engine.cfg)]


class Trainer:
    """value_and_grad of the reference's loss + one optimiser update, for the rank's shard of the batch.

    ``weights`` = {'energy': w_E, 'force': w_F} (the CLI's effective_loss_weights); ``scales`` = the optional per-target
    scalars of get_loss_fn (loss.py:42-48; the CLI's variance scaling 1 / nanvar(target), mlff_train_so3krates.py:330-342):
    the loss term of a target is weight * scale * MSE while the reported metric stays the plain MSE (`_l / scale`).
    ``group``: a torch.distributed process group (or True for the default group) for data-parallel training; None =
    single process.  ``lr_decay_fn`` / ``reduce_lr_on_plateau_fn`` / ``polyak_step_size``: the schedule hooks of the
    reference's train state (see TrainState)."""

    def __init__(self, engine, weights: Dict[str, float], tx: Optional[Optimizer] = None, group=None,
                 scales: Optional[Dict[str, float]] = None, lr_decay_fn: Optional[Callable] = None,
                 reduce_lr_on_plateau_fn: Optional[Callable] = None, polyak_step_size: Optional[float] = None):
        self.engine = engine
        scales = scales or {}
        unknown = (set(weights) | set(scales)) - {'energy', 'force', 'stress'}
        if unknown:
            raise NotImplementedError(f'loss targets {sorted(unknown)}: energy, force and stress targets are built')
        self.w_E = float(weights.get('energy', 0.0)) * float(scales.get('energy', 1.0))
        self.w_F = float(weights.get('force', 0.0)) * float(scales.get('force', 1.0))
        self.w_S = float(weights.get('stress', 0.0)) * float(scales.get('stress', 1.0)) if 'stress' in weights else None
        self.state = TrainState(engine, tx, lr_decay_fn, reduce_lr_on_plateau_fn, polyak_step_size) if tx is not None else None
        self.group = group
        dev = engine.device
        self._sums = torch.zeros(6, dtype=torch.float64, device=dev)       # E: sum sq, n ; F: sum sq, n ; stress: sum sq, n
        self._loss = torch.zeros(3, dtype=torch.float32, device=dev)
        self._grad = torch.zeros(engine.lib.param_count(engine.h), dtype=torch.float32, device=dev)
        self._ws: Optional[torch.Tensor] = None
        self._graph = None                  # captured training step (train_step(..., cuda_graph=True)) for one batch shape
        self._graph_key = None

    def _dist(self):
        if self.group is None:
            return None
        import torch.distributed as dist
        return dist

    def _workspace(self, n_atoms: int, n_slots: int) -> torch.Tensor:
        nb = self.engine.lib.train_workspace_bytes(self.engine.h, n_atoms, n_slots)
        if self._ws is None or self._ws.numel() < nb:
            self._ws = torch.empty(int(nb) + 1024, dtype=torch.uint8, device=self.engine.device)
        return self._ws

    def loss_and_grad(self, R, z, idx_i, idx_j, E_true, F_true, cell=None, cell_offset=None, S_true=None):
        """-> (loss[3] device tensor = (loss, mse_E, mse_F), grad blob, E (B,1), F (B,n,3)); jax.value_and_grad(loss_fn)
        of run.py:46 for this rank's structures, all-reduced over the group.  With a 'stress' weight and S_true (B,3,3) the
        loss gains w_S mse(stress) (get_energy_force_stress_fn, observable_function.py:152-186; masks[pn.stress] all true,
        NaN labels skipped) and loss becomes (loss, mse_E, mse_F, mse_S)."""
        e = self.engine
        dev = e.device
        R = R.to(dev, torch.float32).contiguous()
        z = z.to(dev, torch.int32).contiguous()
        idx_i = idx_i.to(dev, torch.int32).contiguous()
        idx_j = idx_j.to(dev, torch.int32).contiguous()
        E_true = E_true.to(dev, torch.float32).reshape(-1).contiguous()
        F_true = F_true.to(dev, torch.float32).contiguous()
        if cell is not None:
            cell = cell.to(dev, torch.float32).contiguous()
            cell_offset = cell_offset.to(dev, torch.int32).contiguous()
        B, n = z.shape
        P = idx_i.shape[1]
        lib, h, st = e.lib, e.h, e._stream()
        with_stress = self.w_S is not None and S_true is not None
        if with_stress:
            E, F, _, S = e.energy_forces(R, z, idx_i, idx_j, cell, cell_offset, want_stress=True)
            S_true = S_true.to(dev, torch.float32).contiguous()
        else:
            E, F, _ = e.energy_forces(R, z, idx_i, idx_j, cell, cell_offset)
        E1 = E.reshape(-1)
        self._sums.zero_()
        lib.loss_sums(B, n, _p(E1), _p(F), _p(E_true), _p(F_true), _p(z), _p(self._sums), st)
        if with_stress:
            ok = ~torch.isnan(S_true)
            resid = torch.where(ok, S - torch.where(ok, S_true, torch.zeros_like(S_true)), torch.zeros_like(S))
            self._sums[4] = (resid.double() ** 2).sum()
            self._sums[5] = ok.sum().double()
        dist = self._dist()
        grp = None if self.group is True else self.group
        if dist is not None:
            dist.all_reduce(self._sums, group=grp)
        E_bar = torch.empty(B, dtype=torch.float32, device=dev)
        F_bar = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
        lib.loss_cotangents(B, n, _p(E1), _p(F), _p(E_true), _p(F_true), _p(z), _p(self._sums), self.w_E, self.w_F,
                            _p(E_bar), _p(F_bar), _p(self._loss), st)
        S_bar = None
        loss = self._loss
        if with_stress:
            nS = self._sums[5].clamp_min(1.0)
            S_bar = (2.0 * self.w_S * resid.double() / nS).float().contiguous()          # 0 where the label is NaN
            mse_S = torch.where(self._sums[5] > 0, self._sums[4] / nS, torch.zeros_like(nS)).float()
            loss = torch.cat([(self._loss[0] + self.w_S * mse_S).reshape(1), self._loss[1:3], mse_S.reshape(1)])
        ws = self._workspace(B * n, B * P)
        lib.energy_forces_vjp(h, B, n, P, _p(R), _p(z), _p(idx_i), _p(idx_j), _p(cell), _p(cell_offset), _p(E_bar),
                              _p(F_bar), _p(self._grad), None, _p(ws), ws.numel(), _p(e._status), st, S_bar=_p(S_bar))
        if dist is not None:
            dist.all_reduce(self._grad, group=grp)
        return loss, self._grad, E, F

    def eval_step(self, R, z, idx_i, idx_j, E_true, F_true, cell=None, cell_offset=None):
        """valid_step_fn (run.py:88-104): the loss metrics of a batch without gradients -- one E+F pass and the loss sums
        (all-reduced over the group); returns a dict of device scalars."""
        e = self.engine
        dev = e.device
        R = R.to(dev, torch.float32).contiguous()
        z = z.to(dev, torch.int32).contiguous()
        idx_i = idx_i.to(dev, torch.int32).contiguous()
        idx_j = idx_j.to(dev, torch.int32).contiguous()
        E_true = E_true.to(dev, torch.float32).reshape(-1).contiguous()
        F_true = F_true.to(dev, torch.float32).contiguous()
        if cell is not None:
            cell = cell.to(dev, torch.float32).contiguous()
            cell_offset = cell_offset.to(dev, torch.int32).contiguous()
        B, n = z.shape
        E, F, _ = e.energy_forces(R, z, idx_i, idx_j, cell, cell_offset)
        E1 = E.reshape(-1)
        self._sums.zero_()
        e.lib.loss_sums(B, n, _p(E1), _p(F), _p(E_true), _p(F_true), _p(z), _p(self._sums), e._stream())
        dist = self._dist()
        if dist is not None:
            dist.all_reduce(self._sums, group=None if self.group is True else self.group)
        s = self._sums.clone()                                   # (the stress target is evaluated by loss_and_grad only)
        mse_E = torch.where(s[1] > 0, s[0] / s[1].clamp_min(1.0), torch.zeros_like(s[0]))
        mse_F = torch.where(s[3] > 0, s[2] / s[3].clamp_min(1.0), torch.zeros_like(s[2]))
        return {'loss': (self.w_E * mse_E + self.w_F * mse_F).float(), 'energy': mse_E.float(), 'force': mse_F.float()}

    def train_step(self, R, z, idx_i, idx_j, E_true, F_true, cell=None, cell_offset=None, cuda_graph: bool = False,
                   S_true=None):
        """train_step_fn (run.py:34-50): returns a dict of DEVICE scalars (loss, per-target MSE, gradients_norm); nothing
        is read back to the host here.  cuda_graph=True: the whole step for this batch shape -- E+F pass, loss, dual-number
        sweep, the two all-reduces, Adam, parameter re-load: ~500 short launches -- is captured once and replayed (a
        small-molecule batch is launch-bound, above all when it is split over several ranks)."""
        if cuda_graph:
            if S_true is not None:
                raise NotImplementedError('cuda_graph=True with a stress target')
            if self.state.scheduled:
                raise NotImplementedError('cuda_graph=True with a learning-rate schedule or Polyak averaging: the step '
                                          'size is a host value of the captured launch')
            return self._train_step_graphed(R, z, idx_i, idx_j, E_true, F_true, cell, cell_offset)
        loss, grad, _, _ = self.loss_and_grad(R, z, idx_i, idx_j, E_true, F_true, cell, cell_offset, S_true)
        self.state.apply_gradients(grad)
        loss = loss.clone()              # the metrics outlive the next step's buffers
        out = {'loss': loss[0], 'energy': loss[1], 'force': loss[2], 'gradients_norm': self.state.grad_sumsq.sqrt()[0]}
        if loss.numel() > 3:
            out['stress'] = loss[3]
        return out

    def _train_step_graphed(self, R, z, idx_i, idx_j, E_true, F_true, cell, cell_offset):
        dev = self.engine.device
        dts = (torch.float32, torch.int32, torch.int32, torch.int32, torch.float32, torch.float32, torch.float32, torch.int32)
        args = [R, z, idx_i, idx_j, E_true, F_true, cell, cell_offset]
        key = tuple(None if a is None else tuple(a.shape) for a in args)
        if self._graph_key != key:
            self.close()
            self._static = [None if a is None else a.to(dev, dt).contiguous().clone() for a, dt in zip(args, dts)]

            def step():
                loss, grad, _, _ = self.loss_and_grad(*self._static)
                self.state.apply_gradients(grad)
                return torch.cat([loss, self.state.grad_sumsq.sqrt().to(torch.float32)])
            # one eager step on a side stream (workspaces, NCCL warm-up) -- undone afterwards so that the captured sequence
            # starts from the caller's state
            st = self.state
            saved = (st.params.clone(), st.m.clone(), st.v.clone(), st.step, st.step_dev.clone())
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            st.params.copy_(saved[0]); st.m.copy_(saved[1]); st.v.copy_(saved[2]); st.step_dev.copy_(saved[4])
            st.step = saved[3]
            g = torch.cuda.CUDAGraph()
            n0 = self.engine.lib.launch_count()
            with torch.cuda.graph(g):
                self._metrics = step()
            self.captured_launches = self.engine.lib.launch_count() - n0      # libso3k kernels one replay runs
            st.params.copy_(saved[0]); st.m.copy_(saved[1]); st.v.copy_(saved[2]); st.step_dev.copy_(saved[4])
            st.step = saved[3]                       # capturing enqueues nothing, but the host counter moved
            e = self.engine
            e.lib.load_params(e.h, _p(st.params), st.params.numel(), e._stream())
            self._graph, self._graph_key = g, key
        else:
            for s_, a in zip(self._static, args):
                if s_ is not None:
                    s_.copy_(a, non_blocking=True)
        self._graph.replay()
        self.state.step += 1
        m = self._metrics.clone()
        return {'loss': m[0], 'energy': m[1], 'force': m[2], 'gradients_norm': m[3]}

    def close(self):
        """Release a captured step (it holds NCCL kernels: the process group must outlive it)."""
        if self._graph is not None:
            torch.cuda.synchronize(self.engine.device)
            self._graph = None
            self._graph_key = None
