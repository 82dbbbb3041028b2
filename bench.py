#!/usr/bin/env python
"""Benchmark of the SMPLify hot path: fits/sec (100 + 100 Adam iterations per fit).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

One "step" = one SMPLify.__call__ over a batch of `--batch` synthetic samples per GPU
(default 4096, the batch BASELINE.json's target is quoted on).  For N > 1 the driver launches
this file under torchrun; every rank fits its own shard (no inner-loop traffic, weak scaling)
and the step ends with one NCCL all-gather of the packed results (pose, betas, camera,
reprojection loss), as in BASELINE config 4.

--impl reference times the reference's CPU implementation of the same path (the oracle port,
oracle/port.py, all host threads) on a bounded sample of the workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION, set on some boxes) goes there too
if os.environ.get('NCCL_DEBUG', '').upper() in ('', 'VERSION'):
    os.environ['NCCL_DEBUG'] = 'WARN'

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

NUM_ITERS = 100
ALG_GFLOP_PER_FIT = 5.45          # reference formulation, dense regressors
EXEC_MFLOP_PER_FIT = 72.7         # FLOPs the fit kernel actually executes per fit
PACKED = 72 + 10 + 3 + 49         # floats per sample gathered at the end of a sharded step
# dram__bytes_read.sum + dram__bytes_write.sum of one fit-kernel launch at B=4096 (ncu --set full, profiles/)
FIT_KERNEL_DRAM_BYTES_NCU = 5817856   # profiles/fit_kernel_r1.md (dram read 5.74 MB + write 0.08 MB per launch at B = 4096)


def measured_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {'hbm_gbs': d.get('hbm_gbs', 6650.0), 'bf16_tflops': d.get('bf16_tflops', 1590.0),
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', 1400.0), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, universal_newlines=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(',')]
                if len(parts) >= 7:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = [float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit()]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': reasons, 'samples': len(self.samples)}


def time_oracle_cpu(batch_sample, steps, warmup):
    """Reference CPU path (oracle port, eager torch, all host threads)."""
    import torch
    from inbed_pose_estimation_b200 import synthetic
    from oracle import port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle = port.build_oracle(seed=0, num_iters=NUM_ITERS)
    inp = synthetic.make_fit_inputs(batch_sample, seed=7)
    args = lambda: [torch.from_numpy(inp[k].copy()) for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')]
    for _ in range(warmup):
        oracle(*args())
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle(*args())
    dt = time.perf_counter() - t0
    return batch_sample * steps / dt, dt / steps, cores


def time_oracle_cpu_smpl_forward(pose, betas, reps=5):
    """CPU baseline of BASELINE config 1 (SMPL forward, batch 32): the oracle port on all host threads -> (ms per call, cores).
    Lives here because bench.py's cpu_baseline leg is the one place outside tests/ that may execute oracle/."""
    import torch
    from oracle import port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle = port.build_oracle(seed=0, num_iters=1)
    with torch.no_grad():
        oracle.smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
        t0 = time.perf_counter()
        for _ in range(reps):
            oracle.smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    return (time.perf_counter() - t0) / reps * 1e3, cores


def run_reference(a):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    sample = a.ref_batch
    steps, warmup = max(1, min(a.steps, 3)), min(a.warmup, 1)
    fits_s, s_per_step, cores = time_oracle_cpu(sample, steps, warmup)
    line = {
        'impl': 'reference', 'metric': 'smplify_fits_per_sec', 'value': fits_s, 'unit': 'fits/s', 'n_gpus': a.gpus,
        'steps': steps, 'warmup': warmup, 'ms_per_step': 1e3 * s_per_step, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'SMPLify.__call__ 100+100 Adam iterations, batch %d per GPU (CPU sample: %d fits per step)' % (a.batch, sample),
                   'batch_per_gpu': a.batch, 'num_iters': NUM_ITERS},
        'cpu_baseline': {'value': fits_s, 'unit': 'fits/s', 'cores': cores, 'kind': 'port',
                         'sample': '%d fits (100+100 iterations) per step, %d steps, oracle/port.py eager torch fp32' % (sample, steps)},
        'e2e': {'value': fits_s, 'unit': 'fits/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def run_ours(a):
    import torch
    import torch.distributed as dist
    from inbed_pose_estimation_b200 import _native, synthetic

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B = a.batch
    fitter = synthetic.build_smplify(dev, num_iters=NUM_ITERS, seed=0)
    inp = synthetic.make_fit_inputs(B, seed=100 + rank)
    keys = ('pose', 'betas', 'cam_t', 'center', 'keypoints')
    d_in = [torch.from_numpy(inp[k]).to(dev) for k in keys]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2
    gathered = torch.empty((world, B, PACKED), device=dev) if world > 1 else None
    lib = _native.lib()

    def step():
        kp = d_in[4].clone()
        v, j, pose, betas, cam, reproj = fitter(d_in[0], d_in[1], d_in[2], d_in[3], kp)
        if world > 1:
            packed = torch.cat([pose, betas, cam, reproj], dim=1)
            dist.all_gather_into_tensor(gathered.view(world * B, PACKED), packed)
        return reproj

    for _ in range(a.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    lib.smplb200_launch_count(1)
    evs = []
    torch.cuda.synchronize()
    for _ in range(a.steps):
        flush.zero_()                                   # L2 flush between timed iterations (untimed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    launches = int(lib.smplb200_launch_count(0))
    if world > 1:
        dist.barrier()
    total_ms = sum(e0.elapsed_time(e1) for e0, e1 in evs)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())

    # ---- kernel-only timing of the dominant kernel (fit kernel without the vertex pass) ---------------
    ws = torch.empty(lib.smplb200_fit_workspace_bytes(B), dtype=torch.uint8, device=dev)
    outs = [torch.empty((B, n), device=dev) for n in (147, 72, 10, 3, 49)]
    st = torch.cuda.current_stream(dev).cuda_stream
    handle = fitter.smpl.native(dev).handle

    def fit_only():
        kp = d_in[4].clone()
        _native.check(lib.smplb200_smplify_fit(handle, B, NUM_ITERS, 1e-2, 5000., _native.ptr(d_in[0]), _native.ptr(d_in[1]),
                                               _native.ptr(d_in[2]), _native.ptr(d_in[3]), _native.ptr(kp), None,
                                               _native.ptr(outs[0]), _native.ptr(outs[1]), _native.ptr(outs[2]), _native.ptr(outs[3]),
                                               _native.ptr(outs[4]), None, ws.data_ptr(), ws.numel(), st))
    fit_only()
    k_ms = []
    for _ in range(max(3, min(a.steps, 10))):
        flush.zero_()
        kp = d_in[4].clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _native.check(lib.smplb200_smplify_fit(handle, B, NUM_ITERS, 1e-2, 5000., _native.ptr(d_in[0]), _native.ptr(d_in[1]),
                                               _native.ptr(d_in[2]), _native.ptr(d_in[3]), _native.ptr(kp), None,
                                               _native.ptr(outs[0]), _native.ptr(outs[1]), _native.ptr(outs[2]), _native.ptr(outs[3]),
                                               _native.ptr(outs[4]), None, ws.data_ptr(), ws.numel(), st))
        e1.record()
        torch.cuda.synchronize()
        k_ms.append(e0.elapsed_time(e1))
    kernel_ms = float(np.mean(k_ms))

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, H2D + D2H inside) ----------
    h_in = [torch.from_numpy(inp[k].copy()).pin_memory() for k in keys]
    h_out = [torch.empty((B, n), dtype=torch.float32).pin_memory() for n in (147, 72, 10, 3, 49)]
    kp_host = h_in[4].clone().pin_memory()

    def ctypes_ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def e2e_step():
        kp_host.copy_(h_in[4])
        _native.check(lib.smplb200_smplify_fit_host(
            handle, B, NUM_ITERS, 1e-2, 5000., ctypes_ptr(h_in[0]), ctypes_ptr(h_in[1]), ctypes_ptr(h_in[2]), ctypes_ptr(h_in[3]),
            ctypes_ptr(kp_host), None, ctypes_ptr(h_out[0]), ctypes_ptr(h_out[1]), ctypes_ptr(h_out[2]), ctypes_ptr(h_out[3]),
            ctypes_ptr(h_out[4])))

    for _ in range(max(1, a.warmup)):
        e2e_step()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        e2e_step()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    clocks = sampler.stop()

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        fits_s, s_per_step, cores = time_oracle_cpu(a.ref_batch, 1, 1)
        cpu = {'value': fits_s, 'unit': 'fits/s', 'cores': cores, 'kind': 'port',
               'sample': '%d fits (100+100 iterations), 1 warm-up + 1 timed call, oracle/port.py eager torch fp32' % a.ref_batch}

    if rank == 0:
        # which fit kernel ran
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        n16, small, n_small = _native.fit_tile_plan(B, sms)
        if n16 and n_small:
            fit_kernel = 'smplify_fit_mixed_kernel<16,%d> (%d x 16 + %d x %d samples)' % (small, n16, n_small, small)
        else:
            fit_kernel = 'smplify_fit_kernel<%d>' % (16 if n16 else small)
        peaks = measured_peaks()
        fits = world * B * a.steps
        alg_tflops = ALG_GFLOP_PER_FIT * B / (kernel_ms * 1e-3) / 1e3
        exec_tflops = EXEC_MFLOP_PER_FIT * 1e-6 * B / (kernel_ms * 1e-3)
        fp32_peak = ctypes.c_double(0.0)
        if lib.smplb200_probe_fp32_peak(1, ctypes.byref(fp32_peak)) != 0:
            fp32_peak = ctypes.c_double(float('nan'))
        line = {
            'metric': 'smplify_fits_per_sec', 'value': fits / (total_ms * 1e-3), 'unit': 'fits/s', 'n_gpus': world,
            'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': total_ms / a.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'SMPLify.__call__ 100+100 Adam iterations, batch %d per GPU, 49 keypoints, '
                                   'synthetic SMPL-shaped model (6890 verts, 24 joints, 10 betas)' % B,
                       'batch_per_gpu': B, 'num_iters': NUM_ITERS, 'l2': 'flushed between timed steps (256 MiB write)',
                       'gather': 'NCCL all_gather of [B,134] per step' if world > 1 else 'none (1 GPU)'},
            'e2e': {'value': fits / e2e_s, 'unit': 'fits/s', 'h2d_bytes_per_step': B * 234 * 4,
                    'd2h_bytes_per_step': B * (147 + 72 + 10 + 3 + 49 + 147) * 4,
                    'note': 'smplb200_smplify_fit_host: pinned host buffers, H2D of the 5 inputs + fit + vertex kernel + D2H of '
                            'joints/pose/betas/cam/reprojection/keypoints inside the timed region; vertices stay in HBM as in the reference'},
            'gpu_launches': launches,
            'clocks': clocks,
            'roofline': {'bound': 'tensor', 'kernel': fit_kernel, 'achieved': alg_tflops,
                         'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                         'frac': alg_tflops / peaks['bf16_tflops_sustained'], 'traffic': FIT_KERNEL_DRAM_BYTES_NCU,
                         'kernel_ms': kernel_ms, 'peak_source': peaks['source'] + ' bf16 sustained (MEASURED_PEAKS.json)',
                         'note': 'achieved = ALGORITHMIC 5.45 GFLOP/fit (reference formulation, SURVEY 8d) x %d fits per launch / '
                                 'measured kernel time. The kernel runs the constant-folded joint model (declared algebraic '
                                 'saving, DESIGN.md 2), so this can exceed 1; see roofline_executed for executed FLOPs vs the '
                                 'fp32 pipe it actually runs on' % B},
            'roofline_executed': {'bound': 'fp32', 'kernel': fit_kernel, 'achieved': exec_tflops,
                                  'peak': fp32_peak.value, 'unit': 'TFLOP/s', 'frac': exec_tflops / fp32_peak.value,
                                  'peak_source': 'measured live: smplb200_probe_fp32_peak (packed FFMA2)',
                                  'executed_mflop_per_fit': EXEC_MFLOP_PER_FIT},
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=4096, help='samples per GPU per step')
    ap.add_argument('--ref-batch', type=int, default=32, help='bounded CPU sample (fits per reference step)')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
