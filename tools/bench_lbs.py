#!/usr/bin/env python
"""BASELINE.json config 5: SMPL LBS forward + backward throughput sweep (axis-angle mode, upstream
gradients dverts ~ N(0,1), djoints ~ N(0,1)), batch 1K..64K, against the HBM and tensor-pipe rooflines.

    python tools/bench_lbs.py [--batches 1024,4096,...] [--reps 5] [--json out.json]

Every number is CUDA-event time around the C-ABI calls smplb200_smpl_forward / smplb200_smpl_backward
(inputs resident in HBM, L2 flushed between timed calls).  Algorithmic work per sample:
forward 15.85 MFLOP (GEMM-able 14.26), fwd+bwd 31.7 MFLOP; compulsory HBM bytes fwd+bwd 167.2 KB
(forward alone: write vertices 82 680 + joints 588, read pose/betas 328 = 83.6 KB).
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inbed_pose_estimation_b200 import _native, synthetic  # noqa: E402
from inbed_pose_estimation_b200.smpl import SMPL  # noqa: E402

FWD_MFLOP, BWD_MFLOP = 15.85, 15.85
FWD_BYTES = 82680 + 588 + 328
BWD_BYTES = 82680 + 588 + 328                      # read dverts + djoints, write dpose + dbetas
SAVED_BYTES = 2 * 82680                            # v_posed written by forward, re-read by backward (implementation traffic)


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    d = json.load(open(p)) if os.path.exists(p) else {}
    return d.get('hbm_gbs', 6650.0), d.get('bf16_tflops', 1590.0), ('measured' if d else 'fallback')


def timed(fn, flush, reps):
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batches', default='1024,2048,4096,8192,16384,32768,65536')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--json', default=None)
    a = ap.parse_args()
    dev = torch.device('cuda', 0)
    lib = _native.lib()
    smpl = SMPL(model_arrays=synthetic.model_arrays(0), j_regressor_extra=synthetic.make_extra_regressor(1)).to(dev)
    handle = smpl.native(dev).handle
    st = torch.cuda.current_stream(dev).cuda_stream
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hbm, bf16, src = peaks()
    rows = []
    for B in [int(b) for b in a.batches.split(',')]:
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
        P = _native.ptr

        def fwd_nograd():
            _native.check(lib.smplb200_smpl_forward(handle, B, 0, P(pose), P(betas), P(verts), P(joints), None, ws.data_ptr(), ws.numel(), st))

        def fwd():
            _native.check(lib.smplb200_smpl_forward(handle, B, 0, P(pose), P(betas), P(verts), P(joints), P(vposed), ws.data_ptr(), ws.numel(), st))

        def bwd():
            _native.check(lib.smplb200_smpl_backward(handle, B, 0, P(pose), P(betas), P(vposed), P(dverts), P(djoints), P(dpose), P(dbetas),
                                                     ws.data_ptr(), ws.numel(), st))

        def both():
            fwd()
            bwd()

        for _ in range(3):
            both()
            fwd_nograd()
        torch.cuda.synchronize()
        t_f0, t_f, t_b, t_fb = timed(fwd_nograd, flush, a.reps), timed(fwd, flush, a.reps), timed(bwd, flush, a.reps), timed(both, flush, a.reps)
        row = {
            'batch': B, 'fwd_nograd_ms': t_f0, 'fwd_ms': t_f, 'bwd_ms': t_b, 'fwd_bwd_ms': t_fb,
            'fwd_nograd_samples_per_s': B / t_f0 * 1e3, 'fwd_bwd_samples_per_s': B / t_fb * 1e3,
            'fwd_nograd_hbm_gbs': B * FWD_BYTES / t_f0 / 1e6, 'fwd_nograd_hbm_frac': B * FWD_BYTES / t_f0 / 1e6 / hbm,
            'fwd_bwd_hbm_gbs': B * (FWD_BYTES + BWD_BYTES) / t_fb / 1e6, 'fwd_bwd_hbm_frac': B * (FWD_BYTES + BWD_BYTES) / t_fb / 1e6 / hbm,
            'fwd_bwd_hbm_gbs_incl_saved_vposed': B * (FWD_BYTES + BWD_BYTES + SAVED_BYTES) / t_fb / 1e6,
            'fwd_nograd_alg_tflops': B * FWD_MFLOP / t_f0 / 1e3, 'fwd_bwd_alg_tflops': B * (FWD_MFLOP + BWD_MFLOP) / t_fb / 1e3,
            'fwd_bwd_tensor_frac_of_bf16': B * (FWD_MFLOP + BWD_MFLOP) / t_fb / 1e3 / bf16,
        }
        rows.append(row)
        print('B=%6d  fwd(no grad) %8.3f ms  fwd %8.3f ms  bwd %8.3f ms  fwd+bwd %8.3f ms | %9.0f samples/s fwd+bwd | '
              'HBM %.0f GB/s (%.1f%% of %s %.0f) | %.1f alg TFLOP/s' %
              (B, t_f0, t_f, t_b, t_fb, row['fwd_bwd_samples_per_s'], row['fwd_bwd_hbm_gbs'], 100 * row['fwd_bwd_hbm_frac'], src, hbm,
               row['fwd_bwd_alg_tflops']), flush=True)
        del verts, vposed, dverts, ws
        torch.cuda.empty_cache()
    out = {'workload': 'SMPL LBS fwd+bwd sweep (config 5)', 'vertex_kernels': 'tcgen05 3xTF32 (blend GEMM, skinning, dx GEMM, dA GEMM)',
           'hbm_peak_gbs': hbm, 'bf16_peak_tflops': bf16, 'peak_source': src, 'rows': rows}
    if a.json:
        with open(a.json, 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
