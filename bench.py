#!/usr/bin/env python
"""bench.py -- SO3krates energy+force throughput (atom-steps/s) on B200, BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c4|c3|c5|c1|c2] [--impl ours|reference]

Workloads (synthetic seeded positions, random-init weights):
  c4  100 000-atom periodic box (100 A, water-like composition), default model F=132 L=3 r_cut=5   [default]
  c3  3 000-atom periodic water box (31.072 A)
  c5  1 000 atoms, F=256, degrees [1,2,3,4]
  c1  32 x 9-atom ethanol-sized molecules (latency-bound; fits in L2)
  c2  TRAINING step (BASELINE.json configs[1]): 128 x 21-atom aspirin-sized molecules, loss = 0.01 mse(E) + 0.99 mse(F),
      value_and_grad (second order: the loss contains F = -dE/dR) + one Adam update; N > 1: the batch is split over the
      ranks (128 / N structures each), the four loss sums and the 1.3 MB gradient blob are all-reduced over NCCL
A step = one pass of the hot path: device neighbour list (cell list) -> CSR -> L layers -> read-out -> analytic
forces.  `value` times steps with the positions already resident in HBM (CUDA events, max over ranks);
`e2e` times the public calculator call with HOST numpy buffers (pinned H2D of positions, D2H of energy+forces).
N > 1, periodic workloads: ONE system is decomposed into bricks (mlff_b200/dist.py: owned + ghost atoms, mirror edges, a
halo exchange of feature rows per layer and direction over NCCL) -- strong scaling, labelled so at every N including
N = 1; `--parallel replicas` (and the molecular workload c1) runs an independent copy per rank instead (weak scaling).
`--impl reference`: the reference is pure Python on JAX (not installable here), so the arm times the float32
oracle restatement (torch-CPU, all host threads) on a bounded sub-box of the same workload.
The filter GEMMs run on tcgen05 (split-bf16); shapes the tensor-core kernels do not cover make the run FAIL unless
`--fp32-simt` asks for the fp32 CUDA-core kernels explicitly -- there is no silent change of code path.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

from common import jittered_lattice, water_like_numbers  # noqa: E402

WORKLOADS = {
    'c4': dict(n_atoms=100_000, box=100.0, F=132, degrees=(1, 2, 3), L=3, seed=3, desc='100k-atom periodic box, E+F'),
    'c3': dict(n_atoms=3_000, box=31.072, F=132, degrees=(1, 2, 3), L=3, seed=2, desc='3k-atom periodic water box, E+F'),
    'c5': dict(n_atoms=1_000, box=21.544, F=256, degrees=(1, 2, 3, 4), L=3, seed=4, desc='1k atoms, F=256, L_max=4, E+F'),
    'c1': dict(n_atoms=288, box=None, F=132, degrees=(1, 2, 3), L=3, seed=0, desc='32 x 9-atom molecules, E+F'),
    'c2': dict(n_atoms=128 * 21, box=None, F=132, degrees=(1, 2, 3), L=3, seed=1, train=True, batch=128, n_per=21,
               desc='training step: 128 x 21-atom aspirin-sized molecules, value_and_grad of 0.01 mse(E) + 0.99 mse(F) + Adam'),
}


def build_training_batch(w):
    """aspirin composition (C9 H8 O4), one rejection-sampled template in a 7 A ball (min distance 0.9 A)
    + N(0, 0.05 A) jitter per structure, dense neighbour list at r_cut = 5 A padded to the batch maximum with -1,
    targets E ~ N(0,1), F ~ N(0,1).  Seeded numpy only."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import so3k_oracle as o
    rng = np.random.default_rng(w['seed'])
    B, n = w['batch'], w['n_per']
    tmpl = []
    while len(tmpl) < n:
        x = rng.uniform(-3.5, 3.5, 3)
        if np.linalg.norm(x) <= 3.5 and all(np.linalg.norm(x - y) >= 0.9 for y in tmpl):
            tmpl.append(x)
    R = np.asarray(tmpl)[None] + rng.normal(0, 0.05, (B, n, 3))
    z = np.tile(np.array([6] * 9 + [8] * 4 + [1] * 8, np.int64)[None], (B, 1))
    nl = o.get_indices(R, z, 5.0)
    E_true = rng.normal(size=(B, 1))
    F_true = rng.normal(size=(B, n, 3))
    return R, z, nl['idx_i'], nl['idx_j'], E_true, F_true


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


def build_inputs(w, rank=0):
    if w['box'] is None:
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'ethanol_32.npz'))
        pos = (g['R'] + np.arange(32)[:, None, None] * 50.0).reshape(-1, 3)     # 32 molecules 50 A apart
        z = np.tile(g['z'], 32)
        return pos.astype(np.float64), z.astype(np.int64), None, (False, False, False)
    pos = jittered_lattice(w['n_atoms'], w['box'], seed=w['seed'] + 1000 * rank)
    if w['n_atoms'] == 3000:
        z = np.where(np.arange(w['n_atoms']) % 3 == 0, 8, 1).astype(np.int64)
    else:
        z = water_like_numbers(w['n_atoms'], w['seed'])
    order = os.environ.get('SO3K_BENCH_ATOM_ORDER', 'lattice')      # experiment knob: sensitivity to the caller's atom order
    if order != 'lattice':
        if order == 'random':
            perm = np.random.default_rng(5).permutation(len(pos))
        else:                                                      # 'morton': Z-order of 4 A cells
            c = np.floor(pos / 4.0).astype(np.int64) & 1023
            key = np.zeros(len(pos), np.int64)
            for b in range(10):
                for ax in range(3):
                    key |= ((c[:, ax] >> b) & 1) << (3 * b + ax)
            perm = np.argsort(key, kind='stable')
        pos, z = pos[perm], z[perm]
    return pos, z, np.eye(3) * w['box'], (True, True, True)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.idx), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def wait_ready(self, timeout: float = 3.0):
        """Block until the first sample has been written (NVML is initialised by then) or the timeout expires."""
        t0 = time.perf_counter()
        while self.proc is not None and time.perf_counter() - t0 < timeout:
            try:
                if os.path.getsize(self.f.name) > 0:
                    return
            except OSError:
                return
            time.sleep(0.02)

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        rows = [r.strip().split(',') for r in open(self.f.name).read().strip().splitlines() if r.strip()]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
                for k, nm in enumerate(names):
                    if 'Active' in r[5 + k] and 'Not' not in r[5 + k]:
                        reasons.add(nm)
            except Exception:
                pass
        if sm:
            hi = sorted(sm)[len(sm) // 2:]        # under-load half
            out = {'sm_mhz': float(np.median(hi)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                   'samples': len(sm)}
        try:
            os.unlink(self.f.name)
        except Exception:
            pass
        return out


# ------------------------------------------------------------------------------------------------------
def oracle_sample(w, n_sample, threads):
    """float32 oracle (torch-CPU) on a bounded sub-box of the workload: returns a callable doing one E+F."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import so3k_oracle as o
    torch.set_num_threads(threads)
    cfg = o.OracleConfig(F=w['F'], degrees=w['degrees'], n_layers=w['L'])
    p = o.init_params(cfg, 0)
    if w['box'] is None:
        pos, z, cell, pbc = build_inputs(w)
        n_sample = len(pos)
        ii, jj, S = o.primitive_neighbor_list(pos, np.eye(3), (False,) * 3, cfg.r_cut)
        cell = None
    else:
        box = w['box'] * (n_sample / w['n_atoms']) ** (1.0 / 3.0)        # same density
        pos = jittered_lattice(n_sample, box, seed=w['seed'])
        z = water_like_numbers(n_sample, w['seed'])
        cell = np.eye(3) * box
        ii, jj, S = o.periodic_neighbor_list_kdtree(pos, box, cfg.r_cut)

    def step():
        E, F, _ = o.energy_forces(cfg, p, pos, z, ii, jj, dtype=torch.float32, cell=cell, cell_offset=S if cell is not None else None)
        return float(E)
    return step, n_sample, len(ii)


def run_reference(args, w):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_sample = min(w['n_atoms'], 10_000)          # the O(N) model scales linearly; 10k atoms is what the CPU finishes in seconds
    step, n_sample, n_edges = oracle_sample(w, n_sample, threads)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = n_sample / dt
    line = {'impl': 'reference', 'metric': 'SO3krates energy+force atom-steps/s', 'value': val, 'unit': 'atom-steps/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
            'higher_is_better': True, 'scaling': 'strong' if (w['box'] is not None and args.parallel != 'replicas') else 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': args.workload, 'description': w['desc'], 'n_atoms': int(n_sample), 'n_edges': int(n_edges),
                       'n_atoms_full_workload': int(w['n_atoms']), 'F': w['F'], 'n_layers': w['L'],
                       'degrees': list(w['degrees']), 'r_cut': 5.0,
                       'parallelism': 'host threads (torch-CPU), rank 0 only'},
            'cpu_baseline': {'value': val, 'unit': 'atom-steps/s', 'cores': threads, 'kind': 'port',
                             'sample': f'{n_sample}-atom sub-box at the workload density ({n_edges} edges), float32 oracle '
                                       f'restatement on torch-CPU (the JAX reference is not installable here)'},
            'e2e': {'value': val, 'unit': 'atom-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
def oracle_training_sample(w, n_structs, threads):
    """float32 oracle (torch-CPU, double backward) of the training step's value_and_grad on the first structures."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import so3k_oracle as o
    torch.set_num_threads(threads)
    cfg = o.OracleConfig(F=w['F'], degrees=w['degrees'], n_layers=w['L'])
    p = o.init_params(cfg, 0)
    R, z, ii, jj, Et, Ft = (a[:n_structs] for a in build_training_batch(w))

    def step():
        return o.loss_and_param_grads(cfg, p, R, z, ii, jj, Et, Ft, 0.01, 0.99, dtype=torch.float32)[0]
    return step, n_structs * w['n_per'], int((ii >= 0).sum())


def run_reference_training(args, w):
    if int(os.environ.get('RANK', '0')) != 0:
        return
    threads = os.cpu_count() or 1
    step, n_atoms, n_edges = oracle_training_sample(w, 16, threads)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    val = n_atoms / dt
    line = {'impl': 'reference', 'metric': 'SO3krates training-step atom-steps/s', 'value': val, 'unit': 'atom-steps/s',
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt * 1e3,
            'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'c2', 'description': w['desc'], 'n_atoms': int(n_atoms), 'n_edges': int(n_edges),
                       'n_atoms_full_workload': int(w['n_atoms']), 'F': w['F'], 'n_layers': w['L'],
                       'degrees': list(w['degrees']), 'r_cut': 5.0, 'parallelism': 'host threads (torch-CPU), rank 0 only'},
            'cpu_baseline': {'value': val, 'unit': 'atom-steps/s', 'cores': threads, 'kind': 'port',
                             'sample': f'16 of the 128 structures ({n_atoms} atoms, {n_edges} edges): value_and_grad of the loss '
                                       f'by torch double backward of the float32 oracle restatement (no optimiser update; '
                                       f'the JAX reference is not installable here)'},
            'e2e': {'value': val, 'unit': 'atom-steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def run_training(args, w):
    """c2: one training step = E+F pass (tensor mode) + loss sums / cotangents + dual-number sweep (parameter gradient)
    + [all-reduce] + Adam + parameter re-load.  `value` with the batch resident in HBM; `e2e` with HOST batch arrays
    (pinned H2D of R, z, idx_i, idx_j, E_true, F_true every step) and a D2H read of the loss metrics."""
    import torch
    import torch.distributed as dist
    from mlff_b200.config import So3kratesConfig
    from mlff_b200.engine import So3kEngine
    from mlff_b200 import params as P
    from mlff_b200.training import Trainer, Optimizer

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))
    cfg = So3kratesConfig(F=w['F'], n_layer=w['L'], degrees=w['degrees'])
    prm = P.init_params(cfg, 0)
    rng = np.random.default_rng(7)
    for k in prm:
        if k.endswith('/bias'):
            prm[k] = rng.normal(0, 0.01, prm[k].shape).astype(np.float32)
    flags = 0 if args.fp32_simt else 3
    eng = So3kEngine(cfg, dev, flags).load_params(prm)
    tr = Trainer(eng, {'energy': 0.01, 'force': 0.99}, Optimizer().get(learning_rate=1e-4), group=True if world > 1 else None)
    R, z, ii, jj, Et, Ft = build_training_batch(w)
    B, n = z.shape
    assert B % world == 0, 'the batch must divide over the ranks'
    sl = slice(rank * (B // world), (rank + 1) * (B // world))
    host = [np.ascontiguousarray(R[sl], np.float32), np.ascontiguousarray(z[sl], np.int32),
            np.ascontiguousarray(ii[sl], np.int32), np.ascontiguousarray(jj[sl], np.int32),
            np.ascontiguousarray(Et[sl], np.float32), np.ascontiguousarray(Ft[sl], np.float32)]
    pinned = [torch.from_numpy(a).pin_memory() for a in host]
    devb = [t.to(dev) for t in pinned]

    def step_resident():
        return tr.train_step(*devb, cuda_graph=args.train_graph)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    for _ in range(args.warmup):
        m = step_resident()
    barrier()
    l0 = eng.lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        m = step_resident()
    e1.record()
    barrier()
    launches = eng.lib.launch_count() - l0
    if args.train_graph:                  # replays do not pass through the host-side launch counter
        launches = tr.captured_launches * args.steps
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    N_all = B * n
    value = N_all / (ms_step * 1e-3)
    # phases of one step on this rank (CUDA events on the launching stream, outside the timed region)
    phases = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = [ev() for _ in range(4)]
    reps = 5
    acc = [0.0, 0.0, 0.0]
    for _ in range(reps):
        marks[0].record()
        Eo, Fo, _ = eng.energy_forces(devb[0], devb[1], devb[2], devb[3])
        marks[1].record()
        loss, grad, _, _ = tr.loss_and_grad(*devb)
        marks[2].record()
        tr.state.apply_gradients(grad)
        marks[3].record()
        torch.cuda.synchronize()
        acc[0] += marks[0].elapsed_time(marks[1])
        acc[1] += marks[1].elapsed_time(marks[2])
        acc[2] += marks[2].elapsed_time(marks[3])
    phases = {'energy_forces_ms': acc[0] / reps, 'loss_and_grad_ms (E+F pass + loss + dual sweep + all-reduce)': acc[1] / reps,
              'adam_and_reload_ms': acc[2] / reps}
    # ---- e2e: host batch in, metrics out ---------------------------------------------------------------------
    def e2e_call():
        for a, pt in zip(host, pinned):
            pt.copy_(torch.from_numpy(a))
        batch = [pt.to(dev, non_blocking=True) for pt in pinned]
        met = tr.train_step(*batch, cuda_graph=args.train_graph)
        return {k: float(v) for k, v in met.items()}          # D2H of the loss metrics
    for _ in range(2):
        res = e2e_call()
    barrier()
    k_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        res = e2e_call()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / k_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = N_all / float(t_e2e.item())
    h2d = sum(a.nbytes for a in host)
    if rank == 0:
        pk, pk_kind = peaks()
        Fd, K, nl_ = w['F'], 32, len(w['degrees'])
        Fs_ = Fd // 4
        n_edges = int((ii >= 0).sum())
        slots = (B // world) * ii.shape[1]
        # dense work of the dual sweep on this rank: both planes, per filter block and layer -- forward (K F + F F + Fs F),
        # input gradients (F F + Fs F), weight gradients (F F + K F + Fs F); x 2 FLOP, x 2 planes, x 2 blocks, x L layers
        flops = 2.0 * 2 * slots * 2 * w['L'] * ((K * Fd + Fd * Fd + Fs_ * Fd) + (Fd * Fd + Fs_ * Fd) + (Fd * Fd + K * Fd + Fs_ * Fd))
        ms_sweep = max(phases['loss_and_grad_ms (E+F pass + loss + dual sweep + all-reduce)'] - phases['energy_forces_ms'], 1e-6)
        tf = flops / (ms_sweep * 1e-3) / 1e12
        peak_tf = pk.get('bf16_tflops_sustained', pk.get('bf16_tflops'))
        roof = {'kernel': ('k_tc_gemm (tcgen05 split-bf16: forward + input-gradient layers) / k_tr_gemm_tn (fp32 FMA: weight gradients)'
                           if flags else 'k_tr_gemm / k_tr_gemm_tn (fp32 FMA)') + ' of the dual-number sweep, so3k_energy_forces_vjp',
                'bound': 'tensor',
                'achieved': tf, 'peak': peak_tf, 'unit': 'TFLOP/s', 'frac': tf / peak_tf, 'traffic': None,
                'peak_kind': f'{pk_kind} sustained bf16 (kernels timed inside a long step)',
                'ms_per_step': ms_sweep, 'algorithmic_flops_per_step': flops,
                'note': 'algorithmic FLOPs of all dense layers of the sweep (both planes; forward, input gradients, weight '
                        'gradients) over the time of the WHOLE sweep (GEMMs + elementwise + gather kernels), against the bf16 tensor '
                        'peak: a lower bound of the GEMM kernels\' own rate'}
        threads = os.cpu_count() or 1
        cpu = None
        if not args.no_cpu_baseline:
            step, n_s, n_e = oracle_training_sample(w, 8, threads)
            step()
            reps_c, t0 = 0, time.perf_counter()
            while reps_c < 2 or (time.perf_counter() - t0 < 10.0 and reps_c < 50):
                step()
                reps_c += 1
            dt = (time.perf_counter() - t0) / reps_c
            cpu = {'value': n_s / dt, 'unit': 'atom-steps/s', 'cores': threads, 'kind': 'port',
                   'sample': f'8 of the 128 structures ({n_s} atoms, {n_e} edges), {reps_c} value_and_grad evaluations (torch '
                             f'double backward of the float32 oracle restatement; JAX reference not installable)'}
        line = {'metric': 'SO3krates training-step atom-steps/s', 'value': value, 'unit': 'atom-steps/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
                'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': 'c2', 'description': w['desc'], 'n_atoms': N_all, 'n_edges': n_edges,
                           'global_batch': B, 'F': w['F'], 'n_layers': w['L'], 'degrees': list(w['degrees']), 'r_cut': 5.0,
                           'loss_weights': {'energy': 0.01, 'force': 0.99}, 'optimizer': 'adam (optax chain of the reference CLI)',
                           'parallelism': 'single' if world == 1 else
                                          f'data parallel x{world}: {B // world} structures per rank, all-reduce of 4 loss sums + '
                                          f'{tr._grad.numel() * 4} B gradient blob (NCCL)',
                           'first_pass': 'tcgen05 split-operand E+F' if flags else 'fp32 CUDA-core E+F',
                           'step_replayed_as_cuda_graph': bool(args.train_graph),
                           'l2': 'working set of a rank fits in L2 at N > 1; 0.8 GB of kept activations at N = 1'},
                'e2e': {'value': e2e_val, 'unit': 'atom-steps/s', 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': 16},
                'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof, 'cpu_baseline': cpu,
                'phases_ms': phases, 'loss_check': res}
        print(json.dumps(line), flush=True)
    tr.close()                            # a captured step holds NCCL kernels: release it before the process group goes
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
def run_ours(args, w):
    import torch
    import torch.distributed as dist
    from mlff_b200.config import So3kratesConfig
    from mlff_b200.calculator import mlffCalculator
    from mlff_b200 import params as P
    from mlff_b200.capi import NL_ASE

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = f'cuda:{local}'
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(dev))

    cfg = So3kratesConfig(F=w['F'], n_layer=w['L'], degrees=w['degrees'])
    prm = P.init_params(cfg, 0)
    rng = np.random.default_rng(7)
    for k in prm:                                  # exercise the bias paths too
        if k.endswith('/bias'):
            prm[k] = rng.normal(0, 0.01, prm[k].shape).astype(np.float32)
    flags = 0 if args.fp32_simt else (1 if args.no_keep_filters else 3)
    # no fallback: a shape outside the tensor-core kernels' limits raises So3kError(SO3K_ERR_CONFIG) here; the fp32
    # CUDA-core kernels are a different code path and must be asked for with --fp32-simt
    calc = mlffCalculator(cfg, prm, device=dev, flags=flags)
    args.tensor = bool(flags)
    eng = calc.engine
    pos, z, cell, pbc = build_inputs(w, rank)
    N = len(pos)
    periodic = cell is not None

    # ---- warm-up through the public call (sizes the neighbour-list capacity) ----------------------------
    res = calc.calculate(pos, z, cell=cell, pbc=pbc)
    n_edges = calc.n_edges
    cap = calc.capacity
    pos_d = torch.as_tensor(pos, device=dev)
    z_d = torch.as_tensor(z, device=dev, dtype=torch.int32)
    cell32 = None if not periodic else torch.as_tensor(cell, dtype=torch.float32, device=dev)[None]

    def step_resident():
        ii, jj, S, count = eng.neighbor_list(pos_d, calc.cutoff, cap, cell=cell, pbc=pbc, mode=NL_ASE)
        R32 = pos_d.to(torch.float32)
        return eng.energy_forces(R32[None], z_d[None], ii[None], jj[None], cell32, S[None] if periodic else None)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    dd_workload = periodic and args.parallel in ('auto', 'dd')      # one system split over the ranks (strong scaling)
    use_dd = world > 1 and dd_workload
    if use_dd:
        # ---- spatial domain decomposition: every rank holds the (replicated) positions, builds the full neighbour
        # list, keeps the rows of its slab + mirror edges, and exchanges ghost rows once per layer and direction
        from mlff_b200 import dist as dd
        pos, z, cell, pbc = build_inputs(w, 0)              # the same system on every rank
        pos_d = torch.as_tensor(pos, device=dev)
        z_d = torch.as_tensor(z, device=dev, dtype=torch.int32)
        N = len(pos)
        grid = dd.brick_grid(cell, world)
        e_tot = torch.zeros(1, dtype=torch.float64, device=dev)
        # the plan (brick subset, neighbour list with r_cut + skin, owned / ghost order, mirror edges, send lists) is
        # kept while no atom has moved more than skin / 2 -- checked on the device every step; --dd-skin 0 rebuilds it
        # every step
        sess = dd.DDSession(eng, cfg.n_layer, cell, world, rank, calc.cutoff, skin=args.dd_skin, cuda_graph=args.dd_graph,
                            capacity=int(cap * 1.6 / world) + 4096)
        dd_phase = {}

        def step_resident():
            e, f_own, ids = sess.energy_forces(pos_d, z_d, timing=args.dd_timing)
            if args.dd_timing:
                dd_phase.update(sess.phase_times())
            step_resident.p_local = sess.dd.P
            step_resident.n_local = sess.dd.N
            step_resident.ids = ids
            e_tot[0] = e
            dist.all_reduce(e_tot)
            return e_tot, f_own, None

    # the clock / throttle sampler (nvidia-smi polling) is started BEFORE the warm-up so that its start-up (NVML
    # initialisation touches every GPU of the box) does not land in the first timed steps; it keeps sampling through the
    # timed region, which is where its samples are needed
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    # ---- timed region: K steps, CUDA events, per-kernel-category events inside libso3k --------------------
    # per-kernel CUDA-event scopes (roofline numbers): inside the timed region on one GPU; on N > 1 ranks the ~150 event
    # records per step are a visible share of a 10 ms step, so they run in a separate pass after the timed region
    prof_in_loop = world == 1
    eng.lib.profile_collect(eng.h, reset=True)
    eng.lib.profile_enable(eng.h, prof_in_loop)
    nl_ev = []
    l0 = eng.lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    launches = eng.lib.launch_count() - l0
    if use_dd and args.dd_graph and sess._graph is not None:      # replays do not pass through the host-side launch counter
        launches = sess.captured_launches * args.steps
    if use_dd and args.dd_timing and rank == 0:
        print('dd step phases of the last step (ms): ' + ', '.join(f'{k} {v:.3f}' for k, v in dd_phase.items())
              + f' | plan builds so far: {sess.n_builds}', file=sys.stderr)
    ms_total = e0.elapsed_time(e1)
    if not prof_in_loop:
        eng.lib.profile_enable(eng.h, True)
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(args.steps):
            step_resident()
        pe1.record()
        barrier()
        ms_prof_total = pe0.elapsed_time(pe1)
    prof = eng.lib.profile_collect(eng.h, reset=True)
    eng.lib.profile_enable(eng.h, False)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = (N if use_dd else world * N) / (ms_step * 1e-3)

    # ---- e2e: host buffers through the public calculator call ---------------------------------------------
    if use_dd:
        # e2e: host positions in (pinned H2D of the replicated positions on every rank), owned forces + energy out
        pin = torch.empty(pos.shape, dtype=torch.float64).pin_memory()

        f_full = torch.zeros(N, 3, dtype=torch.float32, device=dev)

        def e2e_call():
            # host positions in on every rank; the WHOLE force array is assembled on rank 0 (owned rows scattered into a
            # full-size array, summed onto rank 0 over NVLink) and read back to its host together with the energy
            pin.copy_(torch.from_numpy(pos))
            pos_d.copy_(pin, non_blocking=True)
            e, f, _ = step_resident()
            f_full.zero_()
            f_full[step_resident.ids] = f
            dist.reduce(f_full, dst=0)
            if rank == 0:
                return {'energy': float(e.item()), 'forces': f_full.cpu().numpy()}
            return {'energy': float(e.item()), 'forces': None}
    else:
        def e2e_call():
            return calc.calculate(pos, z, cell=cell, pbc=pbc)
    for _ in range(2):
        e2e_call()
    barrier()
    k_e2e = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(k_e2e):
        res = e2e_call()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / k_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = (N if use_dd else world * N) / float(t_e2e.item())
    h2d = N * 3 * 8 + (0 if use_dd else N * 4)
    d2h = N * 3 * 4 + 8 if use_dd else (N * 3 + 3) * 8      # dd: rank 0 reads the assembled float32 forces + the energy

    # standalone featurisation kernel (rbf, phi, d, unit vector, Y_lm written out; not on the fused path, where these are
    # recomputed on chip): achieved bytes against the HBM peak, north_star's "featurisation kernel GB/s"
    feat_info = None
    if world == 1 and rank == 0:
        ii_f, jj_f, S_f, cnt_f = eng.neighbor_list(pos_d, calc.cutoff, cap, cell=cell, pbc=pbc, mode=NL_ASE)
        nv_f = int(cnt_f.item())
        R32_f = pos_d.to(torch.float32)
        fe0, fe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fargs = dict(cell=cell32[0] if periodic else None, cell_offset=S_f[:nv_f] if periodic else None)
        for _ in range(3):                                   # warm-up: the output buffers come from the caching allocator
            fout = eng.featurise(ii_f[:nv_f], jj_f[:nv_f], N, R=R32_f, **fargs)
        fe0.record()
        for _ in range(3):
            fout = eng.featurise(ii_f[:nv_f], jj_f[:nv_f], N, R=R32_f, **fargs)
        fe1.record()
        torch.cuda.synchronize()
        ms_f = fe0.elapsed_time(fe1) / 3
        M_f = sum(2 * l + 1 for l in w['degrees'])
        bytes_f = nv_f * (4 * (32 + 1 + 1 + 3 + M_f) + 8 + (12 if periodic else 0)) + 12 * N
        feat_info = {'kernel': 'k_featurise', 'edges': nv_f, 'ms': ms_f, 'algorithmic_bytes': bytes_f,
                     'GB/s': bytes_f / (ms_f * 1e-3) / 1e9}
        del fout, ii_f, jj_f, S_f

    md_info = None
    if args.md_steps > 0 and use_dd:
        # the same MD loop with the force call decomposed over the ranks (md.CalculatorDD: one all-reduce of forces + energy)
        from mlff_b200 import md
        masses = torch.as_tensor(np.where(z == 8, 15.999, np.where(z == 6, 12.011, 1.008)), device=dev)
        md_sess = dd.DDSession(eng, cfg.n_layer, cell, world, rank, calc.cutoff, skin=args.md_skin, cuda_graph=args.dd_graph,
                               capacity=int(cap * 1.6 / world) + 4096)
        atoms = md.AtomsX(positions=pos_d.clone(), momenta=md.maxwell_boltzmann_momenta(masses, 300.0, seed=1), masses=masses,
                          numbers=z_d, cell=cell, pbc=pbc).init_domain_decomposition(md_sess)
        integ = md.VelocityVerletX.create(0.5 * md.FS, md.CalculatorDD(md_sess))
        atoms, _ = md.simulate(integ, atoms, 3)
        barrier()
        b0, t0 = md_sess.n_builds, time.perf_counter()
        atoms, st_md = md.simulate(integ, atoms, args.md_steps)
        barrier()
        dt_md = (time.perf_counter() - t0) / args.md_steps
        etot = st_md['kinetic_energy'] + st_md['potential_energy']
        md_info = {'steps': args.md_steps, 'ms_per_step': dt_md * 1e3, 'atom_steps_per_s': N / dt_md, 'timestep_fs': 0.5,
                   'skin_A': args.md_skin, 'ranks': world, 'plan_rebuilds': md_sess.n_builds - b0,
                   'total_energy_drift_over_kinetic': float(np.abs(etot - etot[0]).max() / st_md['kinetic_energy'].mean())}
        md_sess.close()
    if args.md_steps > 0 and world == 1:
        # MD inner loop on the device: velocity Verlet, skin list re-used until an atom moved skin / 2
        from mlff_b200 import md
        masses = torch.as_tensor(np.where(z == 8, 15.999, np.where(z == 6, 12.011, 1.008)), device=dev)
        atoms = md.AtomsX(positions=pos_d.clone(), momenta=md.maxwell_boltzmann_momenta(masses, 300.0, seed=1), masses=masses,
                          numbers=z_d, cell=cell, pbc=pbc if periodic else (False, False, False)
                          ).init_spatial_partitioning(eng, calc.cutoff, args.md_skin)
        integ = md.VelocityVerletX.create(0.5 * md.FS, md.CalculatorX(eng, cuda_graph=args.md_graph))
        atoms, _ = md.simulate(integ, atoms, 3)
        torch.cuda.synchronize()
        b0, t0 = atoms.neighbors.n_builds, time.perf_counter()
        atoms, st_md = md.simulate(integ, atoms, args.md_steps)
        torch.cuda.synchronize()
        dt_md = (time.perf_counter() - t0) / args.md_steps
        etot = st_md['kinetic_energy'] + st_md['potential_energy']
        md_info = {'steps': args.md_steps, 'ms_per_step': dt_md * 1e3, 'atom_steps_per_s': N / dt_md, 'timestep_fs': 0.5,
                   'skin_A': args.md_skin, 'cuda_graph': bool(args.md_graph), 'edges_with_skin': int(atoms.neighbors.count),
                   'list_rebuilds': atoms.neighbors.n_builds - b0,
                   'total_energy_drift_over_kinetic': float(np.abs(etot - etot[0]).max() / st_md['kinetic_energy'].mean())}

    if rank == 0:
        pk, pk_kind = peaks()
        nl_ = len(w['degrees'])
        Fd, K = w['F'], 32
        p_rank = getattr(step_resident, 'p_local', n_edges)        # edges this rank's kernels processed (owned + mirror)
        Fs_ = Fd // 4
        M_ = sum(2 * l + 1 for l in w['degrees'])
        Ast_ = (4 + nl_ + 3) // 4 * 4
        n_rank = getattr(step_resident, 'n_local', N)
        flop_unit = 2.0 * (K * Fd + Fd * Fd + nl_ * Fs_ + Fs_ * Fd)          # one filter block, one edge
        peak_tf = pk.get('bf16_tflops_sustained', pk.get('bf16_tflops'))
        peak_bw = pk['hbm_gbs']

        ms_share_ref = ms_total if prof_in_loop else ms_prof_total        # the pass the scopes were recorded in

        def tensor_roof(kernel, cat, flops_scope, note):
            ms, n = prof[cat]
            ms_scope = ms / max(n, 1)
            tf = flops_scope / (ms_scope * 1e-3) / 1e12 if ms_scope > 0 else 0.0
            return {'kernel': kernel, 'bound': 'tensor', 'achieved': tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                    'frac': tf / peak_tf, 'traffic': None,
                    'peak_kind': f'{pk_kind} sustained bf16 (kernel timed inside a long step)',
                    'ms_per_layer_scope': ms_scope, 'algorithmic_flops_per_scope': flops_scope, 'note': note,
                    'share_of_step': ms / (ms_share_ref if ms_share_ref > 0 else 1.0)}

        def hbm_roof(kernel, cat, bytes_scope, note):
            ms, n = prof[cat]
            ms_scope = ms / max(n, 1)
            gbs = bytes_scope / (ms_scope * 1e-3) / 1e9 if ms_scope > 0 else 0.0
            return {'kernel': kernel, 'bound': 'hbm', 'achieved': gbs, 'peak': peak_bw, 'unit': 'GB/s',
                    'frac': gbs / peak_bw, 'traffic': None, 'peak_kind': f'{pk_kind} HBM copy bandwidth',
                    'ms_per_layer_scope': ms_scope, 'algorithmic_bytes_per_scope': bytes_scope, 'note': note,
                    'share_of_step': ms / (ms_share_ref if ms_share_ref > 0 else 1.0)}

        if args.tensor:
            # tensor mode works on undirected pairs (the filter is symmetric): units = p_rank / 2 pairs, 2 blocks per layer
            n_pairs = p_rank / 2.0
            split = 'algorithmic fp32-equivalent FLOPs; the split-bf16 scheme issues 3 MMAs per product on padded shapes'
            roofs = [
                tensor_roof('k_filter_tc (per layer: 2 blocks)', 'edge_fwd', 2.0 * n_pairs * flop_unit, split),
                tensor_roof('k_edge_bwd2_tc (per layer: 2 blocks)', 'edge_bwd',
                            2.0 * n_pairs * 2.0 * (2 * K * Fd + Fd * (Fd + Fs_) + nl_ * Fs_),
                            split + '; two radial-basis GEMMs + the W2^T GEMM + the l0 transposes per pair'),
                hbm_roof('k_alpha_aggregate', 'aggregate',
                         n_pairs * 8 * Fd + n_rank * (5 * 4 * Fd + 8 * Fd + 8 * M_) + p_rank * (4 * Ast_ + 8 + 16),
                         'compulsory bytes: both filters once per pair, 5 projection rows + x/chi in and out per atom, '
                         'alpha out + (col, pid, geom) per edge; the sender-row gathers are served by L2'),
                hbm_roof('k_qk_bwd_stream (per layer: 2 blocks)', 'qk_stream',
                         n_pairs * 16 * Fd + n_rank * 32 * Fd + p_rank * (8 * Ast_ + 4 + 12 + 16),
                         'compulsory bytes: filters in + filter cotangents out once per pair, q/k rows in + their '
                         'gradients out per atom, alphabar + alpha in, ebar out, (col, rev, pid, geom) per edge'),
                hbm_roof('k_alpha_bwd', 'alpha_bwd',
                         n_rank * (4 * Fd * 3 + 8 * M_) + p_rank * (4 * 4 + 4 * Ast_ + 8 + 16),
                         'compulsory bytes: v and xbar1 rows in, v gradient out per atom; alpha (reverse edge) in, '
                         'alphabar out, (col, rev, geom) per edge'),
            ]
        else:
            roofs = [
                tensor_roof('k_edge_bwd', 'edge_bwd', 4.0 * p_rank * flop_unit,
                            'fp32 FMA on the CUDA cores, reported against the bf16 tensor peak; backward ~ 2 x forward'),
                tensor_roof('k_edge_fwd', 'edge_fwd', 2.0 * p_rank * flop_unit,
                            'fp32 FMA on the CUDA cores, reported against the bf16 tensor peak'),
                hbm_roof('k_aggregate', 'aggregate',
                         (4 * 8 + 4 + 16) * p_rank + (8 * Fd + 8 * M_ + 4) * n_rank + 4 * Fd * n_rank,
                         'compulsory bytes (DESIGN.md)'),
            ]
        # `traffic` (ncu dram__bytes_read + dram__bytes_write per scope) cannot be measured inside this run (a number taken
        # under a profiler is not a bench value): it is filled from profiles/ncu_traffic.json, written by
        # profiles/summarize_traffic.py from `ncu --set full` captures of THIS workload, and only when that file names the
        # kernel sources it was captured on (sha256 over the source files of the profiled kernels, mlff_b200._lib.PROFILED_SOURCES); otherwise it stays null.
        tpath = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
        if os.path.exists(tpath) and world == 1:
            from mlff_b200 import _lib as _l
            tj = json.load(open(tpath))
            if tj.get('source_sha16') == _l.source_sha16() and tj.get('workload') == args.workload:
                for r_ in roofs:
                    for k_, v_ in tj.get('bytes_per_scope', {}).items():
                        if r_['kernel'].startswith(k_ + ' ') or r_['kernel'] == k_:
                            r_['traffic'] = v_
        roofs.sort(key=lambda r: -r['share_of_step'])
        roof, agg = roofs[0], roofs[1:]          # `roofline` = the kernel with the largest share of the step
        # CPU baseline: bounded sample, all host threads
        threads = os.cpu_count() or 1
        cpu = None
        if not args.no_cpu_baseline:
            n_s = min(N, 10_000)
            step, n_s, n_e = oracle_sample(w, n_s, threads)
            step()
            reps, t0 = 0, time.perf_counter()
            while reps < 2 or (time.perf_counter() - t0 < 10.0 and reps < 50):
                step()
                reps += 1
            dt = (time.perf_counter() - t0) / reps
            cpu = {'value': n_s / dt, 'unit': 'atom-steps/s', 'cores': threads, 'kind': 'port',
                   'sample': f'{n_s}-atom sub-box at the workload density ({n_e} edges), {reps} E+F evaluations of the '
                             f'float32 oracle restatement on torch-CPU (JAX reference not installable)'}
        line = {'metric': 'SO3krates energy+force atom-steps/s', 'value': value, 'unit': 'atom-steps/s', 'n_gpus': world,
                'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step, 'higher_is_better': True,
                'scaling': 'strong' if dd_workload else 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': args.workload, 'description': w['desc'], 'n_atoms': N, 'n_edges': int(n_edges),
                           'F': w['F'], 'n_layers': w['L'], 'degrees': list(w['degrees']), 'r_cut': 5.0,
                           'parallelism': 'single' if world == 1 else (
                               f'domain decomposition: {grid[0]}x{grid[1]}x{grid[2]} bricks, mirror edges, per-layer halo exchange (NCCL all_to_all), '
                               f'plan re-used within a {args.dd_skin} A skin ({sess.n_builds} builds)' + (', sweep replayed as one CUDA graph' if args.dd_graph else '')
                               if use_dd else f'replicas x{world} (no collective)'),
                           'filter_gemm': 'tcgen05 split-operand' if args.tensor else 'fp32 CUDA-core',
                           'filters': 'kept for the backward' if (args.tensor and not args.no_keep_filters) else 'recomputed',
                           'l2': 'working set (edge + node arrays) larger than the 126 MB L2' if N >= 50_000 else
                                 'working set fits in L2 (latency/L2-bound configuration)'},
                'e2e': {'value': e2e_val, 'unit': 'atom-steps/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
                'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roof, 'roofline_others': agg,
                'cpu_baseline': cpu,
                'kernel_ms_per_step': {k: round(v[0] / args.steps, 4) for k, v in prof.items()},
                'featurise': (dict(feat_info, peak=pk['hbm_gbs'], frac=feat_info['GB/s'] / pk['hbm_gbs']) if feat_info else None),
                'md': md_info, 'energy_check': float(res['energy']),
                'rank0_local': {'atoms': int(getattr(step_resident, 'n_local', N)), 'edges': int(p_rank)}}
        print(json.dumps(line), flush=True)
    if world > 1:
        if use_dd:
            sess.close()                  # a captured sweep holds NCCL kernels: release it before the process group goes
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c4', choices=sorted(WORKLOADS))
    ap.add_argument('--md-graph', action='store_true', help='replay the force evaluation of the MD loop as one CUDA graph')
    ap.add_argument('--md-skin', type=float, default=0.5, help='skin (Angstrom) of the MD neighbour list')
    ap.add_argument('--md-steps', type=int, default=0, help='additionally run this many velocity-Verlet MD steps with a '
                    '0.5 A skin list (mlff_b200.md, one GPU) and report them under "md"')
    ap.add_argument('--dd-skin', type=float, default=0.2, help='domain decomposition: skin (Angstrom) of the per-rank plan; the plan '
                    'is re-used while no atom has moved more than skin / 2 (0: rebuild every step)')
    ap.add_argument('--dd-eager', action='store_true', help='domain decomposition: launch the sweep kernel by kernel instead of replaying '
                    'the sweep of a kept plan as one CUDA graph (kernels + halo pack / NCCL exchange / unpack; the default)')
    ap.add_argument('--train-graph', action='store_true', help='c2: capture the training step once and replay it as one CUDA graph')
    ap.add_argument('--dd-timing', action='store_true', help='print the per-phase device time of one domain-decomposed step '
                    '(rank 0, stderr)')
    ap.add_argument('--no-keep-filters', action='store_true', help='recompute the filters in the backward instead of keeping '
                    'them in HBM (2F floats per edge and layer)')
    ap.add_argument('--fp32-simt', action='store_true', help='run the filter GEMMs as fp32 FMA on the CUDA cores instead of '
                    'tcgen05 split-bf16 (the default where the shapes are supported)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--parallel', default='auto', choices=['auto', 'dd', 'replicas'],
                    help='N > 1: dd = spatial domain decomposition of the ONE system with per-layer halo exchange over '
                         'NCCL (strong scaling; default for the periodic workloads), replicas = one independent copy per rank')
    args = ap.parse_args()
    args.dd_graph = not args.dd_eager and args.dd_skin > 0 and not args.dd_timing
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    w = WORKLOADS[args.workload]
    if w.get('train'):
        (run_reference_training if args.impl == 'reference' else run_training)(args, w)
    elif args.impl == 'reference':
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == '__main__':
    main()
