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
EXEC_MFLOP_PER_FIT = 72.7         # useful FLOPs of the folded formulation per fit: what the CUDA-core kernel executes
# tcgen05 MMAs the pair kernel issues per pair of CTAs (32 samples): forward call 3 tiles x (28 k-steps folded + 9 k-steps prior)
# x 3 (3xTF32) = 333, backward call 88 k-steps x 3 = 264; each is M=256 N=32 K=8 (padded rows / samples included)
PAIR_MMAS_FWD_CALL, PAIR_MMAS_BWD_CALL, PAIR_MMA_FLOP = 333, 264, 2 * 256 * 32 * 8
PACKED = 72 + 10 + 3 + 49         # floats per sample gathered at the end of a sharded step
BULK = 65536                      # BASELINE config 4: samples of the bulk refit, split over the ranks
# SURVEY.md §8d, per LBS sample: forward 15.85 MFLOP (GEMM-able 14.26), fwd+bwd 31.7 (GEMM-able 28.5); compulsory HBM bytes
LBS_FWD_MFLOP, LBS_FWDBWD_MFLOP, LBS_GEMM_FWDBWD_MFLOP = 15.85, 31.7, 28.5
LBS_FWD_BYTES, LBS_FWDBWD_BYTES = 83596, 167192


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


def ncu_dram_bytes_per_sample():
    """DRAM bytes per LBS sample (forward + backward) from the committed ncu --set full summary, if there is one
    (profiles/lbs_tc_kernels_r*.json written by tools/ncu_summary.py); None otherwise - never a constant made up here."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'lbs_tc_kernels_r*.json'))):
        try:
            with open(path) as f:
                d = json.load(f)
            best = {'bytes_per_sample': d['dram_bytes_per_sample'], 'batch': d['batch'], 'source': os.path.relpath(path, ROOT)}
        except Exception:
            pass
    return best


def ncu_fit_kernel_traffic(batch):
    """dram__bytes_read.sum + dram__bytes_write.sum of one fit-kernel launch at `batch` from the committed ncu summary
    (profiles/fit_kernel_r*.json), or None."""
    import glob
    best = None
    for path in sorted(glob.glob(os.path.join(ROOT, 'profiles', 'fit_kernel_r*.json'))):
        try:
            with open(path) as f:
                d = json.load(f)
            if int(d.get('batch', -1)) == int(batch):
                best = float(d['dram_bytes_per_launch'])
        except Exception:
            pass
    return best


def time_torch_gpu_reference(dev, batches=(256, 4096)):
    """The oracle port - the reference's eager-PyTorch op sequence (smplify/smplify.py:40-136 on torch autograd + Adam) - run ON
    THE GPU: BASELINE.json's ">= 100x the reference's single-GPU PyTorch SMPLify" denominator.  One warm-up call at the
    smallest batch, then one timed call per batch (about 1 + 3 + 30 s).  Baseline leg: the oracle is the thing being timed
    as the REFERENCE here, never as part of the product path."""
    import torch
    from inbed_pose_estimation_b200 import synthetic
    from oracle import port
    oracle = port.build_oracle(seed=0, num_iters=NUM_ITERS)
    oracle.smpl = oracle.smpl.to(dev)
    oracle.pose_prior = oracle.pose_prior.to(dev)
    keys = ('pose', 'betas', 'cam_t', 'center', 'keypoints')
    rows, warm = [], False
    for B in batches:
        inp = synthetic.make_fit_inputs(B, seed=7)
        args = lambda: [torch.from_numpy(inp[k].copy()).to(dev) for k in keys]
        try:
            if not warm:
                small = synthetic.make_fit_inputs(32, seed=7)
                oracle.num_iters = 3
                oracle(*[torch.from_numpy(small[k].copy()).to(dev) for k in keys])     # cuBLAS handles, allocator
                oracle.num_iters = NUM_ITERS
                warm = True
            torch.cuda.synchronize(dev)
            torch.cuda.reset_peak_memory_stats(dev)
            t0 = time.perf_counter()
            oracle(*args())
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t0
            rows.append({'batch': B, 'seconds_per_call': dt, 'fits_per_sec': B / dt,
                         'peak_mem_gb': torch.cuda.max_memory_allocated(dev) / 2 ** 30})
        except RuntimeError as e:
            rows.append({'batch': B, 'error': str(e)[:160]})
    del oracle
    torch.cuda.empty_cache()
    return rows


def time_lbs(dev, fitter, batches, flush, tf32_peak, hbm_gbs):
    """BASELINE config 5 rows (SMPL LBS forward and forward + backward, axis-angle mode, dverts / djoints ~ N(0,1)) through
    smplb200_smpl_forward / smplb200_smpl_backward: CUDA events, L2 flushed between timed calls, median of 5."""
    import torch
    from inbed_pose_estimation_b200 import _native
    lib = _native.lib()
    handle = fitter.smpl.native(dev).handle
    st = torch.cuda.current_stream(dev).cuda_stream
    P = _native.ptr
    ncu = ncu_dram_bytes_per_sample()
    rows = []
    for B in batches:
        g = torch.Generator(device='cpu').manual_seed(B)
        pose = (0.2 * torch.randn(B, 72, generator=g)).to(dev)
        betas = (0.5 * torch.randn(B, 10, generator=g)).to(dev)
        verts = torch.empty(B, 6890, 3, device=dev)
        vposed = torch.empty(B, _native.VPOSED_PITCH, device=dev)
        joints = torch.empty(B, 49, 3, device=dev)
        dverts = torch.randn(B, 6890, 3, device=dev)
        djoints = torch.randn(B, 49, 3, device=dev)
        dpose, dbetas = torch.empty(B, 72, device=dev), torch.empty(B, 10, device=dev)
        ws = torch.empty(lib.smplb200_smpl_workspace_bytes(B), dtype=torch.uint8, device=dev)

        def fwd_nograd():
            _native.check(lib.smplb200_smpl_forward(handle, B, 0, P(pose), P(betas), P(verts), P(joints), None, ws.data_ptr(), ws.numel(), st))

        def both():
            _native.check(lib.smplb200_smpl_forward(handle, B, 0, P(pose), P(betas), P(verts), P(joints), P(vposed), ws.data_ptr(), ws.numel(), st))
            _native.check(lib.smplb200_smpl_backward(handle, B, 0, P(pose), P(betas), P(vposed), P(dverts), P(djoints), P(dpose), P(dbetas),
                                                     ws.data_ptr(), ws.numel(), st))

        def timed(fn):
            ms = []
            for _ in range(5):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                torch.cuda.synchronize(dev)
                ms.append(e0.elapsed_time(e1))
            return float(np.median(ms))

        for _ in range(2):
            both()
            fwd_nograd()
        torch.cuda.synchronize(dev)
        t_f, t_fb = timed(fwd_nograd), timed(both)
        gemm_tflops = 3.0 * LBS_GEMM_FWDBWD_MFLOP * B / t_fb / 1e3          # 3xTF32: three MMAs per fp32-accurate product
        rows.append({
            'batch': B, 'fwd_ms': t_f, 'fwd_bwd_ms': t_fb,
            'fwd_samples_per_s': B / t_f * 1e3, 'fwd_bwd_samples_per_s': B / t_fb * 1e3,
            'fwd_hbm_frac': B * LBS_FWD_BYTES / t_f / 1e6 / hbm_gbs, 'fwd_bwd_hbm_frac': B * LBS_FWDBWD_BYTES / t_fb / 1e6 / hbm_gbs,
            'fwd_bwd_alg_tflops': B * LBS_FWDBWD_MFLOP / t_fb / 1e3,
            'fwd_bwd_tf32_mma_tflops': gemm_tflops, 'fwd_bwd_tensor_frac': (gemm_tflops / tf32_peak) if tf32_peak else None,
            'dram_bytes_per_sample_ncu': ncu,
        })
        del verts, vposed, dverts, ws
        torch.cuda.empty_cache()
    return rows


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
    from inbed_pose_estimation_b200 import _native, sharded, synthetic

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
    packed = torch.empty((B, PACKED), device=dev) if world > 1 else None   # written by the fit kernel itself
    gathered = torch.empty((world * B, PACKED), device=dev) if world > 1 else None
    lib = _native.lib()

    def step():
        kp = d_in[4].clone()
        out = fitter(d_in[0], d_in[1], d_in[2], d_in[3], kp, packed_out=packed)
        if world > 1:
            dist.all_gather_into_tensor(gathered, packed)
        return out[5]

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

    def fit_only(kp):
        _native.check(lib.smplb200_smplify_fit(handle, B, NUM_ITERS, 1e-2, 5000., _native.ptr(d_in[0]), _native.ptr(d_in[1]),
                                               _native.ptr(d_in[2]), _native.ptr(d_in[3]), _native.ptr(kp), None,
                                               _native.ptr(outs[0]), _native.ptr(outs[1]), _native.ptr(outs[2]), _native.ptr(outs[3]),
                                               _native.ptr(outs[4]), None, None, ws.data_ptr(), ws.numel(), st))
    fit_only(d_in[4].clone())
    k_ms = []
    for _ in range(max(3, min(a.steps, 10))):
        flush.zero_()
        kp = d_in[4].clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fit_only(kp)
        e1.record()
        torch.cuda.synchronize()
        k_ms.append(e0.elapsed_time(e1))
    kernel_ms = float(np.mean(k_ms))

    # ---- end to end from pinned HOST buffers, host <-> device copies inside the timed region ------------------------------
    # N = 1: the host-buffer C-ABI call smplb200_smplify_fit_host (H2D, fit, vertex kernels, D2H on two streams).
    # N > 1: the same work per rank through the public Python API - H2D of the rank's inputs, SMPLify.__call__, NCCL
    #        all-gather of the packed rows, D2H of the GATHERED [N*B,134] rows and of the rank's joints - so that the gather
    #        is inside the number.  Vertices stay in HBM in both (the reference keeps them on the device too).
    h_in = [torch.from_numpy(inp[k].copy()).pin_memory() for k in keys]
    h_out = [torch.empty((B, n), dtype=torch.float32).pin_memory() for n in (147, 72, 10, 3, 49)]
    kp_host = h_in[4].clone().pin_memory()
    h_gathered = torch.empty((world * B, PACKED), dtype=torch.float32).pin_memory() if world > 1 else None

    def ctypes_ptr(t):
        return ctypes.c_void_p(t.data_ptr())

    def e2e_step():
        if world == 1:
            kp_host.copy_(h_in[4])
            _native.check(lib.smplb200_smplify_fit_host(
                handle, B, NUM_ITERS, 1e-2, 5000., ctypes_ptr(h_in[0]), ctypes_ptr(h_in[1]), ctypes_ptr(h_in[2]), ctypes_ptr(h_in[3]),
                ctypes_ptr(kp_host), None, ctypes_ptr(h_out[0]), ctypes_ptr(h_out[1]), ctypes_ptr(h_out[2]), ctypes_ptr(h_out[3]),
                ctypes_ptr(h_out[4])))
        else:
            d = [h.to(dev, non_blocking=True) for h in h_in]
            out = fitter(d[0], d[1], d[2], d[3], d[4], packed_out=packed)
            dist.all_gather_into_tensor(gathered, packed)
            h_gathered.copy_(gathered, non_blocking=True)
            h_out[0].copy_(out[1].view(B, 147), non_blocking=True)
            kp_host.copy_(d[4], non_blocking=True)                   # the in-place confidence zeroing reaches the host copy
            torch.cuda.synchronize()

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

    # ---- BASELINE config 4 at this N: bulk refit of 65 536 samples split over the ranks (strong scaling) -------------------
    config4 = None
    if not a.no_config4:
        n_bulk = BULK
        lo, hi = sharded.shard_bounds(n_bulk, world, rank)
        bulk = synthetic.make_fit_inputs(hi - lo, seed=4000 + rank)
        fits_l = torch.from_numpy(np.concatenate([bulk['pose'], bulk['betas']], axis=1)).to(dev)
        cam_l, cen_l, kp_l = (torch.from_numpy(bulk[k]).to(dev) for k in ('cam_t', 'center', 'keypoints'))
        loss_l = fitter.get_fitting_loss(fits_l[:, :72].contiguous(), fits_l[:, 72:].contiguous(), cam_l, cen_l, kp_l.clone()).mean(dim=-1)
        packed_l = torch.empty((hi - lo, PACKED), device=dev)

        def bulk_step():
            fitter(fits_l[:, :72].contiguous(), fits_l[:, 72:].contiguous(), cam_l, cen_l, kp_l.clone(), packed_out=packed_l)
            allp = sharded.gather_rows(packed_l, n_bulk)
            pose_g, betas_g, cam_g, reproj_g = sharded.unpack_results(allp)
            old = sharded.gather_rows(loss_l[:, None], n_bulk)[:, 0] if world > 1 else loss_l
            oldf = sharded.gather_rows(fits_l, n_bulk) if world > 1 else fits_l
            return sharded.keep_if_better(oldf, old, torch.cat([pose_g, betas_g], dim=1), reproj_g.mean(dim=-1))

        bulk_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms4 = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            bulk_step()
            e1.record()
            torch.cuda.synchronize()
            ms4.append(e0.elapsed_time(e1))
        t4 = torch.tensor([float(np.median(ms4))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t4, op=dist.ReduceOp.MAX)
        config4 = {'workload': 'bulk refit of %d samples split over %d GPU(s): SMPLify 100+100, NCCL all-gather of [N,134], keep-if-better'
                               % (n_bulk, world), 'samples': n_bulk, 'samples_per_gpu': hi - lo, 'ms': float(t4.item()),
                   'fits_per_sec': n_bulk / float(t4.item()) * 1e3, 'scaling': 'strong',
                   'tile_plan': _native.fit_tile_plan(hi - lo, torch.cuda.get_device_properties(dev).multi_processor_count)}
        del fits_l, cam_l, cen_l, kp_l, packed_l, bulk
        torch.cuda.empty_cache()

    # ---- BASELINE configs 2 and 3: the reference's own operating points (--batch_size 32; batch 256), device-timed ----------
    small_rows = None
    if rank == 0 and world == 1 and not a.no_small:
        small_rows = {'note': 'SMPLify.__call__ 100+100 on a resident batch, median of 7 calls after 3 warm-ups, L2 flushed between calls; '
                         'clusters of `cluster_size` CTAs per 4-sample tile (csrc/fit_split.cuh)', 'rows': []}
        sms_here = torch.cuda.get_device_properties(dev).multi_processor_count
        for b in (32, 256):
            si = synthetic.make_fit_inputs(b, seed=9000 + b)
            sa = [torch.from_numpy(si[k]).to(dev) for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')]
            ms_s = []
            for i in range(10):
                flush.zero_()
                kp = sa[4].clone()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fitter(sa[0], sa[1], sa[2], sa[3], kp)
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ms_s.append(e0.elapsed_time(e1))
            small_rows['rows'].append({'batch': b, 'ms': float(np.median(ms_s)), 'fits_per_sec': b / float(np.median(ms_s)) * 1e3,
                                  'cluster_size': _native.fit_split_plan(b)})

    cpu = gpu_ref = lbs = None
    tf32_peak = None
    if rank == 0 and world == 1:
        pk = ctypes.c_double(0.0)
        if lib.smplb200_probe_tf32_peak(ctypes.byref(pk)) == 0:
            tf32_peak = pk.value
        if not a.no_lbs:
            lbs = {'workload': 'BASELINE config 5 rows: SMPL LBS forward / forward+backward, axis-angle mode, dverts and djoints ~ N(0,1)',
                   'tf32_peak_tflops': tf32_peak,
                   'tf32_peak_source': 'measured live: smplb200_probe_tf32_peak (tcgen05 kind::tf32 M=128 N=256 K=8 back to back on every SM)',
                   'algorithmic_per_sample': {'fwd_mflop': LBS_FWD_MFLOP, 'fwd_bwd_mflop': LBS_FWDBWD_MFLOP,
                                              'gemm_fwd_bwd_mflop': LBS_GEMM_FWDBWD_MFLOP, 'fwd_bytes': LBS_FWD_BYTES,
                                              'fwd_bwd_bytes': LBS_FWDBWD_BYTES},
                   'rows': time_lbs(dev, fitter, (4096, 16384), flush, tf32_peak, measured_peaks()['hbm_gbs'])}
        if not a.no_cpu_baseline:
            fits_s, s_per_step, cores = time_oracle_cpu(a.ref_batch, 1, 1)
            cpu = {'value': fits_s, 'unit': 'fits/s', 'cores': cores, 'kind': 'port',
                   'sample': '%d fits (100+100 iterations), 1 warm-up + 1 timed call, oracle/port.py eager torch fp32' % a.ref_batch}
        if not a.no_gpu_reference:
            rows = time_torch_gpu_reference(dev)
            gpu_ref = {'what': "the reference's eager-PyTorch SMPLify (oracle/port.py: the same op sequence as smplify/smplify.py:40-136, "
                               'torch autograd + torch.optim.Adam) run on THIS GPU, one timed call per batch',
                       'rows': rows}

    if rank == 0:
        # which fit kernel ran
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        n16, small, n_small = _native.fit_tile_plan(B, sms)
        pairs, p16, p12 = _native.fit_pair_plan(B, sms)
        if pairs:
            fit_kernel = 'smplify_fit_pair_kernel<16,12> (%d pairs of 2 x 16 + %d pairs of 2 x 12 samples, 2-CTA clusters)' % (p16, p12)
        elif n16 and n_small:
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
        value = fits / (total_ms * 1e-3)
        if pairs:
            # executed tensor-pipe work per launch: every pair makes NUM_ITERS + 2 forward calls and NUM_ITERS backward calls
            mma_flop = (p16 + p12) * ((NUM_ITERS + 2) * PAIR_MMAS_FWD_CALL + NUM_ITERS * PAIR_MMAS_BWD_CALL) * PAIR_MMA_FLOP
            tensor_tflops = mma_flop / (kernel_ms * 1e-3) / 1e12
            pk = ctypes.c_double(0.0)
            tf32 = pk.value if lib.smplb200_probe_tf32_peak(ctypes.byref(pk)) == 0 else None
            roofline = {'bound': 'tensor', 'kernel': fit_kernel, 'achieved': tensor_tflops, 'peak': tf32, 'unit': 'TFLOP/s',
                        'frac': tensor_tflops / tf32 if tf32 else None, 'traffic': ncu_fit_kernel_traffic(B), 'kernel_ms': kernel_ms,
                        'kernel_share_of_step': kernel_ms / (total_ms / a.steps),
                        'peak_source': 'measured live: smplb200_probe_tf32_peak (tcgen05 kind::tf32, M=128 N=256 K=8 back to back)',
                        'executed_mma_gflop_per_launch': mma_flop / 1e9,
                        'note': 'EXECUTED tcgen05 kind::tf32 MMA FLOPs (3xTF32, M=256 N=32 K=8 per MMA, padded rows and the 2 x 12-sample '
                                'pairs of the last wave included) over the kernel time.  The per-iteration GEMMs are 60 % of the kernel; '
                                'they are bound by the hand-shake latency of the operand ring (one tcgen05.commit per 32-k chunk), not by '
                                'the tensor pipe or L2: see profiles/fit_pair_kernel_r2.md.  roofline_fp32_equivalent gives the figure '
                                'comparable with round 1.'}
        else:
            roofline = None
        line = {
            'metric': 'smplify_fits_per_sec', 'value': value, 'unit': 'fits/s', 'n_gpus': world,
            'steps': a.steps, 'warmup': a.warmup, 'ms_per_step': total_ms / a.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': 'SMPLify.__call__ 100+100 Adam iterations, batch %d per GPU, 49 keypoints, '
                                   'synthetic SMPL-shaped model (6890 verts, 24 joints, 10 betas)' % B,
                       'batch_per_gpu': B, 'num_iters': NUM_ITERS, 'l2': 'flushed between timed steps (256 MiB write)',
                       'gather': 'NCCL all_gather of [B,134] per step (rows written by the fit kernel)' if world > 1 else 'none (1 GPU)',
                       'vertices': 'computed every step, stay in HBM'},
            'e2e': {'value': fits / e2e_s, 'unit': 'fits/s', 'h2d_bytes_per_step': B * 234 * 4,
                    'd2h_bytes_per_step': (B * (147 + 72 + 10 + 3 + 49 + 147) * 4) if world == 1 else (world * B * PACKED + B * 294) * 4,
                    'note': ('smplb200_smplify_fit_host: pinned host buffers, H2D of the 5 inputs + fit + vertex kernels + D2H of '
                             'joints/pose/betas/cam/reprojection/keypoints inside the timed region; vertices stay in HBM as in the reference')
                    if world == 1 else
                            ('per rank: H2D of the 5 inputs from pinned memory, SMPLify.__call__, NCCL all-gather of the packed rows, D2H of '
                             'the gathered [N*B,134] rows + joints + keypoints, all inside the timed region (max over ranks); vertices stay in HBM')},
            'gpu_launches': launches,
            'clocks': clocks,
            'roofline': roofline if roofline is not None else
                        {'bound': 'fp32', 'kernel': fit_kernel, 'achieved': exec_tflops, 'peak': fp32_peak.value, 'unit': 'TFLOP/s',
                         'frac': exec_tflops / fp32_peak.value, 'traffic': ncu_fit_kernel_traffic(B), 'kernel_ms': kernel_ms,
                         'kernel_share_of_step': kernel_ms / (total_ms / a.steps),
                         'peak_source': 'measured live: smplb200_probe_fp32_peak (packed FFMA2 on every SM)',
                         'executed_mflop_per_fit': EXEC_MFLOP_PER_FIT,
                         'note': 'the CUDA-core fit kernel is bound by the fp32 FMA pipe (tensor pipe 0 %%, HBM < 0.1 %% - '
                                 'profiles/): achieved = the %.1f MFLOP it EXECUTES per fit x %d fits / kernel time' % (EXEC_MFLOP_PER_FIT, B)},
            'roofline_fp32_equivalent': {'bound': 'fp32', 'achieved': exec_tflops, 'peak': fp32_peak.value, 'unit': 'TFLOP/s',
                                         'frac': exec_tflops / fp32_peak.value,
                                         'note': 'useful work of the folded formulation (%.1f MFLOP per fit, one multiply-add per product, no '
                                                 'padding) over the kernel time, against the measured fp32 FFMA2 peak: the round-1 figure of '
                                                 'record (0.42), comparable across kernels' % EXEC_MFLOP_PER_FIT},
            'roofline_algorithmic': {'bound': 'tensor', 'achieved': alg_tflops, 'peak': peaks['bf16_tflops_sustained'], 'unit': 'TFLOP/s',
                                     'frac': alg_tflops / peaks['bf16_tflops_sustained'],
                                     'peak_source': peaks['source'] + ' bf16 sustained (MEASURED_PEAKS.json)',
                                     'note': 'SURVEY 8d algorithmic count (5.45 GFLOP per fit, the reference formulation: full LBS in each '
                                             'of the 201 forwards) over the kernel time.  It exceeds 1 because of a declared 75x ALGEBRAIC '
                                             'saving (DESIGN.md 2: the joint regressors are folded through the blend basis and skinning '
                                             'weights once, in float64, so an iteration costs 0.72 instead of 31.7 MFLOP) - a statement '
                                             'about the algebra, not about kernel quality; `roofline` is the figure of record'},
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        if gpu_ref is not None:
            for r in gpu_ref['rows']:
                if 'fits_per_sec' in r and r['batch'] == B:
                    gpu_ref['speedup_at_batch_%d' % B] = value / r['fits_per_sec']
            line['gpu_reference'] = gpu_ref
        if lbs is not None:
            line['lbs'] = lbs
        if config4 is not None:
            line['config4'] = config4
        if small_rows is not None:
            line['small_batch'] = small_rows
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
    ap.add_argument('--no-gpu-reference', action='store_true', help='skip the eager-PyTorch-on-GPU denominator (about 40 s)')
    ap.add_argument('--no-lbs', action='store_true', help='skip the BASELINE config 5 rows')
    ap.add_argument('--no-config4', action='store_true', help='skip the 65 536-sample bulk refit')
    ap.add_argument('--no-small', action='store_true', help='skip the batch-32 / batch-256 rows (BASELINE configs 2, 3)')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)


if __name__ == '__main__':
    main()
