#!/usr/bin/env python
"""Benchmark of the PV-RAFT hot path (BASELINE.json metric: RAFT iters/sec at N=8192, iters=32;
corr-kernel HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one full `RSF.forward(p, num_iters=32)` (encoders + correlation build + 32 RAFT
iterations) on a batch of synthetic N=8192 cloud pairs with seeded random-init weights.
value = sample-iterations/s = global_batch * iters / T_forward (CUDA events, max over ranks).
Rank 0 prints ONE JSON line.  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
import types

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS, TRUNC_K, ITERS, LEVELS, BASE_SCALE = 8192, 512, 32, 3, 0.25
BATCH_PER_GPU = 8        # 8 x 32 MiB of (corr, index) state = 268 MB > the 126 MB L2: the lookup streams from HBM


def alg_bytes_lookup(n, k, levels=LEVELS):
    """ALGORITHMIC bytes of one sample-iteration of the lookup kernel:
    K*(4 B corr + 4 B index) + 12 B coords in, levels*27*4 B voxel means + 32*16 B kNN vectors out."""
    return n * (k * 8 + 12 + levels * 27 * 4 + 32 * 16)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        try:
            return json.load(open(path)), 'measured'
        except Exception:   # noqa: BLE001
            pass
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:   # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:   # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, flag in zip(names, f[3:7]):
                if flag.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': mx, 'samples': len(sm),
                'reasons': sorted(reasons)}


def make_args():
    return types.SimpleNamespace(corr_levels=LEVELS, base_scales=BASE_SCALE, truncate_k=TRUNC_K)


def synthetic_clouds(b, n, seed):
    g = torch.Generator().manual_seed(seed)
    pc1 = 10.0 * torch.rand(b, n, 3, generator=g)
    pc2 = pc1 + 0.1 * torch.randn(b, n, 3, generator=g)
    return pc1, pc2


# --------------------------------------------------------------------------------------------------
# CPU side: the reference's own formulation (oracle port, torch CPU ops) on the host cores
# --------------------------------------------------------------------------------------------------
_CPU_STATE = {}


def _cpu_weights():
    if 'W' not in _CPU_STATE:
        from pvraft_b200 import RSF
        torch.manual_seed(0)
        _CPU_STATE['W'] = {k: v.detach().clone() for k, v in RSF(make_args()).state_dict().items()}
    return _CPU_STATE['W']


def cpu_pick_threads():
    """torch CPU ops stop scaling (and then collapse) long before 100+ threads on these op sizes: time a
    small forward at a few thread counts and keep the fastest (reported as `cores`)."""
    if 'threads' in _CPU_STATE:
        return _CPU_STATE['threads']
    from oracle import pvraft_oracle as O
    W = _cpu_weights()
    pc1, pc2 = synthetic_clouds(1, 2048, 7)
    ncpu = os.cpu_count() or 1
    best, best_t = 1, float('inf')
    for th in sorted({t for t in (4, 8, 16, 32, 64, ncpu) if t <= ncpu}):
        torch.set_num_threads(th)
        with torch.no_grad():
            t0 = time.perf_counter()
            O.rsf_forward(W, pc1, pc2, 1, LEVELS, BASE_SCALE, TRUNC_K)
            t = time.perf_counter() - t0
        if t < best_t:
            best, best_t = th, t
    _CPU_STATE['threads'] = best
    return best


def cpu_sample(threads, loop_iters=4, n=N_POINTS):
    """One bounded CPU sample of the bench workload at B=1: everything before the loop (encoders, graphs,
    correlation build) + `loop_iters` RAFT iterations, timed separately.  Returns (t_prepare, t_per_iteration)."""
    from oracle import pvraft_oracle as O
    W = _cpu_weights()
    torch.set_num_threads(threads)
    pc1, pc2 = synthetic_clouds(1, n, 1234)
    with torch.no_grad():
        t0 = time.perf_counter()
        li = O.prepare(W, pc1, pc2, TRUNC_K)
        t1 = time.perf_counter()
        O.raft_loop(W, li, pc1, loop_iters, LEVELS, BASE_SCALE)
        t2 = time.perf_counter()
    return t1 - t0, (t2 - t1) / loop_iters


def cpu_value(t_prep, t_iter):
    """sample-iterations/s of a full forward (ITERS iterations) from the two measured parts."""
    return ITERS / (t_prep + ITERS * t_iter)


def run_reference(a):
    """`--impl reference`: the reference's CPU formulation (oracle port; the reference itself is pure PyTorch
    and is not present on the GPU box) timed on the host cores, same metric / unit / config.  A step is the
    bounded sample of cpu_sample(): the full pre-loop work + 4 of the 32 iterations at B=1, scaled to 32."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    threads = cpu_pick_threads()
    warm = min(a.warmup, 1)
    for _ in range(warm):
        cpu_sample(threads, 1)
    steps = max(1, a.steps)
    t_prep = t_iter = 0.0
    t_begin = time.perf_counter()
    done = 0
    for _ in range(steps):
        tp, ti = cpu_sample(threads)
        t_prep += tp
        t_iter += ti
        done += 1
        if time.perf_counter() - t_begin > 200.0:      # keep the whole arm within a few minutes
            break
    t_prep /= done
    t_iter /= done
    value = cpu_value(t_prep, t_iter)
    line = {
        'impl': 'reference', 'metric': 'raft_sample_iters_per_sec', 'value': value, 'unit': 'sample-iterations/s',
        'n_gpus': a.gpus, 'steps': done, 'warmup': warm, 'ms_per_step': 1e3 * (t_prep + ITERS * t_iter),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(1, 1),
        'cpu_baseline': {'value': value, 'unit': 'sample-iterations/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{done} x (pre-loop work + 4 of {ITERS} RAFT iterations) on B=1, N={N_POINTS}, scaled to '
                                   f'{ITERS} iterations: t_prepare={t_prep:.2f} s, t_iteration={t_iter:.3f} s; torch CPU ops, '
                                   f'{threads} of {os.cpu_count()} host threads (fastest of a thread sweep)'},
        'e2e': {'value': value, 'unit': 'sample-iterations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    return line


def workload_config(batch_per_gpu, world):
    return {'workload': f'RSF.forward: N={N_POINTS} pts x2 clouds, truncate_k={TRUNC_K}, corr_levels={LEVELS}, '
                        f'iters={ITERS}, batch {batch_per_gpu}/GPU, fp32 (BASELINE.json metric config; batch from configs[2])',
            'global_batch': batch_per_gpu * world, 'points': N_POINTS, 'truncate_k': TRUNC_K, 'iters': ITERS,
            'parallelism': f'batch-shard x{world} (no data-path collective)',
            'l2_policy': 'per-iteration candidate state (B*N*K*8 B = 268 MB/GPU) exceeds the 126 MB L2; no explicit flush'}


# --------------------------------------------------------------------------------------------------
# native arm
# --------------------------------------------------------------------------------------------------
def run_native(a):
    from pvraft_b200 import RSF, ops
    from pvraft_b200 import dist as D
    rank, world, local = D.init_from_env()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    B = a.batch
    torch.manual_seed(0)
    model = RSF(make_args()).to(dev).eval()
    pc1_h, pc2_h = synthetic_clouds(B, N_POINTS, 1234 + rank)
    pc1_h, pc2_h = pc1_h.pin_memory(), pc2_h.pin_memory()
    pc1, pc2 = pc1_h.to(dev), pc2_h.to(dev)
    out_h = torch.empty(B, N_POINTS, 3).pin_memory()

    def step_resident():
        with torch.no_grad():
            return model([pc1, pc2], ITERS)[-1]

    def step_e2e():
        with torch.no_grad():
            flows = model([pc1_h.to(dev, non_blocking=True), pc2_h.to(dev, non_blocking=True)], ITERS)
            out_h.copy_(flows[-1], non_blocking=True)
        return flows[-1]

    def timed(fn, steps, sample_clocks=False):
        D.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local) if sample_clocks and rank == 0 else None
        if sampler:
            sampler.start()
        l0 = ops.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        D.barrier()
        ms = D.max_over_ranks(e0.elapsed_time(e1), dev)
        clocks = sampler.stop() if sampler else None
        return ms, ops.launch_count - l0, clocks

    for _ in range(max(a.warmup, 3)):
        step_resident()
    ms, launches, clocks = timed(step_resident, a.steps, sample_clocks=True)
    if clocks and set(clocks['reasons']) & {'hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown'}:
        ms, launches, clocks = timed(step_resident, a.steps, sample_clocks=True)     # re-measure once
    step_e2e()
    ms_e2e, _, _ = timed(step_e2e, a.steps)
    gb = B * world
    value = gb * ITERS * a.steps / (ms * 1e-3)
    e2e = gb * ITERS * a.steps / (ms_e2e * 1e-3)

    # ---- dominant kernel: the fused correlation lookup, timed in situ with CUDA events ----------------
    lk_ms = []
    orig = ops.corr_lookup

    def hooked(*args, **kw):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig(*args, **kw)
        e.record()
        lk_ms.append((s, e))
        return r

    ops.corr_lookup = hooked
    for _ in range(2):
        step_resident()
    torch.cuda.synchronize()
    ops.corr_lookup = orig
    durs = [s.elapsed_time(e) for s, e in lk_ms]
    lookup_ms = statistics.mean(durs)
    peaks, peak_kind = measured_peaks()
    alg = alg_bytes_lookup(N_POINTS, TRUNC_K) * B
    achieved = alg / (lookup_ms * 1e-3) / 1e9
    roofline = {'kernel': 'k_corr_lookup (pvraft_corr_lookup_fwd)', 'bound': 'hbm', 'achieved': achieved,
                'peak': peaks['hbm_gbs'], 'peak_kind': peak_kind + ' (MEASURED_PEAKS.json hbm_gbs)' if peak_kind == 'measured' else 'fallback',
                'unit': 'GB/s', 'frac': achieved / peaks['hbm_gbs'], 'traffic': None,
                'alg_bytes_per_launch': alg, 'avg_launch_ms': lookup_ms, 'launches_timed': len(durs),
                'share_of_step': lookup_ms * ITERS / (ms / a.steps)}
    traffic_file = os.path.join(ROOT, 'profiles', 'lookup_dram_bytes.json')
    if os.path.exists(traffic_file):
        try:
            roofline['traffic'] = json.load(open(traffic_file)).get('dram_bytes_per_launch')
        except Exception:   # noqa: BLE001
            pass

    if rank != 0:
        return None
    cpu = None
    if world == 1 and not a.no_cpu:
        threads = cpu_pick_threads()
        tp, ti = cpu_sample(threads)
        cpu = {'value': cpu_value(tp, ti), 'unit': 'sample-iterations/s', 'cores': threads, 'kind': 'port',
               'sample': f'pre-loop work + 4 of {ITERS} RAFT iterations on B=1, N={N_POINTS}, scaled to {ITERS} iterations '
                         f'(t_prepare={tp:.2f} s, t_iteration={ti:.3f} s; oracle port of the reference, torch CPU ops, '
                         f'{threads} of {os.cpu_count()} host threads)'}
    line = {
        'metric': 'raft_sample_iters_per_sec', 'value': value, 'unit': 'sample-iterations/s', 'n_gpus': world,
        'steps': a.steps, 'warmup': max(a.warmup, 3), 'ms_per_step': ms / a.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': workload_config(B, world),
        'e2e': {'value': e2e, 'unit': 'sample-iterations/s', 'h2d_bytes_per_step': 2 * B * N_POINTS * 3 * 4 * world,
                'd2h_bytes_per_step': B * N_POINTS * 3 * 4 * world, 'ms_per_step': ms_e2e / a.steps},
        'gpu_launches': launches, 'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
    }
    return line


class _QuietStdout:
    """Libraries (NCCL's version banner, warnings) write to fd 1; the contract is ONE JSON line on stdout.
    Route fd 1 to stderr while the benchmark runs and restore it for the final print."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=40)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', choices=['native', 'reference'], default='native')
    ap.add_argument('--batch', type=int, default=BATCH_PER_GPU, help='samples per GPU')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    a = ap.parse_args()
    with _QuietStdout():
        line = run_reference(a) if a.impl == 'reference' else run_native(a)
    if line is not None:
        print(json.dumps(line), flush=True)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == '__main__':
    main()
