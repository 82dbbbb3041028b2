"""Per-phase cycle counts of the fit kernel's stage-2 loop (CTA 0, thread 0), from a profiling build.

    python tools/phase_clocks.py build     # here (no GPU): compiles build/libsmplify_b200_clk.so with -DSMPLB200_PHASE_CLOCKS
    python tools/phase_clocks.py run [--batch 4096] [--iters 100]     # on the GPU box
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'inbed_pose_estimation_b200', 'libsmplify_b200_clk.so')
NAMES = ['prior GEMM', 'prior select + angle/shape priors', 'Rodrigues + pose features + rest joints', '(mark)',
         'folded GEMM forward || chain forward', '49 output joints', 'projection + GMoF (+ trace)', 'source grads + joint backward',
         'picked-vertex backward', '(mark)', 'folded GEMM backward || chain backward', 'Rodrigues backward + Adam']

if sys.argv[1] == 'build':
    from inbed_pose_estimation_b200 import _native
    print(_native.build(force=True, extra_flags=['-DSMPLB200_PHASE_CLOCKS'], out=OUT))
else:
    os.environ['SMPLB200_LIB'] = OUT
    import argparse
    import torch
    from inbed_pose_estimation_b200 import _native, synthetic
    ap = argparse.ArgumentParser()
    ap.add_argument('cmd')
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--iters', type=int, default=100)
    a = ap.parse_args()
    lib = _native.lib()
    fitter = synthetic.build_smplify('cuda', num_iters=a.iters, seed=0)
    inp = synthetic.make_fit_inputs(a.batch, seed=1)
    args = [torch.from_numpy(inp[k]).cuda() for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')]
    buf = (ctypes.c_ulonglong * 32)()
    for rep in range(3):
        lib.smplb200_debug_phase_clocks(buf, 1)
        fitter(args[0], args[1], args[2], args[3], args[4].clone())
        torch.cuda.synchronize()
    lib.smplb200_debug_phase_clocks(buf, 0)
    tot = sum(buf[i] for i in range(12))
    print('stage-2 cycles per iteration (CTA 0): %.0f' % (tot / a.iters))
    for i, n in enumerate(NAMES):
        print('%-42s %9.0f clk/iter  %5.1f%%' % (n, buf[i] / a.iters, 100.0 * buf[i] / tot))
