#!/usr/bin/env python
"""Development aid: the pair fit kernel (tcgen05 GEMMs shared by a 2-CTA cluster) against the CUDA-core tile kernel on the same
inputs.  Runs itself twice (SMPLB200_FIT_VARIANT is read once per process) and compares loss traces and results.

    python tools/pair_debug.py [--batch 64] [--iters 3]
"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run_one(a):
    import torch
    from inbed_pose_estimation_b200 import synthetic
    fitter = synthetic.build_smplify('cuda', num_iters=a.iters, seed=0)
    inp = synthetic.make_fit_inputs(a.batch, seed=a.seed)
    args = [torch.from_numpy(inp[k].copy()).cuda() for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')]
    out = fitter(*args, return_loss_trace=True)
    torch.cuda.synchronize()
    trace = fitter.last_loss_trace.cpu().numpy()
    ms = None
    if a.time:
        for _ in range(2):
            fitter(*[t.clone() for t in args])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fitter(*[t.clone() for t in args])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
    np.savez(a.out, trace=trace, verts=out[0][:, ::16].cpu().numpy(), joints=out[1].cpu().numpy(),
             pose=out[2].cpu().numpy(), betas=out[3].cpu().numpy(), cam=out[4].cpu().numpy(), reproj=out[5].cpu().numpy(),
             ms=np.float64(ms if ms is not None else -1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=64)
    ap.add_argument('--iters', type=int, default=3)
    ap.add_argument('--seed', type=int, default=5)
    ap.add_argument('--out', default=None)
    ap.add_argument('--time', action='store_true')
    ap.add_argument('--variant', default='10', help="SMPLB200_FIT_VARIANT of the second run: 10 = pair kernel, 12 = small-batch cluster kernel")
    a = ap.parse_args()
    if a.out:
        run_one(a)
        return
    tmp = tempfile.mkdtemp()
    res = {}
    for name, variant in (('tile', '11'), ('pair', a.variant)):
        out = os.path.join(tmp, name + '.npz')
        env = dict(os.environ, SMPLB200_FIT_VARIANT=variant)
        cmd = [sys.executable, os.path.abspath(__file__), '--batch', str(a.batch), '--iters', str(a.iters), '--seed', str(a.seed), '--out', out]
        if a.time:
            cmd.append('--time')
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=600)
        if r.returncode != 0:
            print('%s kernel FAILED (rc %d):\n%s' % (name, r.returncode, r.stdout[-3000:]))
            return 1
        res[name] = np.load(out)
    t, p = res['tile'], res['pair']
    print('batch %d, %d + %d iterations; tile kernel %.3f ms, variant-%s kernel %.3f ms' % (a.batch, a.iters, a.iters, float(t['ms']), a.variant, float(p['ms'])))
    tr_t, tr_p = t['trace'], p['trace']
    rel = np.abs(tr_p - tr_t) / np.maximum(np.abs(tr_t), 1e-30)
    for i in range(tr_t.shape[0]):
        if 8 < i < tr_t.shape[0] - 4 and i % 25:
            continue
        print('  iteration %3d (%s): max rel loss diff %.3e   (loss sum %.6e vs %.6e)%s' %
              (i, 'camera' if i < a.iters else 'body', rel[i].max(), tr_p[i].astype(np.float64).sum(), tr_t[i].astype(np.float64).sum(),
               '' if np.all(np.isfinite(tr_p[i])) else '  NON-FINITE'))
    for k in ('pose', 'betas', 'cam', 'joints', 'verts', 'reproj'):
        d = np.abs(p[k] - t[k])
        print('  %-7s max abs diff %.3e (row of the max: %d)' % (k, d.max(), int(np.unravel_index(d.argmax(), d.shape)[0])))
    return 0


if __name__ == '__main__':
    sys.exit(main() or 0)
