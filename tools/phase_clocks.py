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
    print(_native.build(force=True, extra_flags=['-DSMPLB200_PHASE_CLOCKS'] + sys.argv[2:], out=OUT))
else:
    os.environ['SMPLB200_LIB'] = OUT
    import argparse
    import torch
    from inbed_pose_estimation_b200 import _native, synthetic
    ap = argparse.ArgumentParser()
    ap.add_argument('cmd')
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--iters', type=int, default=100)
    ap.add_argument('--pair', action='store_true', help='slot names of the pair kernel (fit_pair.cuh)')
    ap.add_argument('--split', action='store_true', help='slot names of the small-batch cluster kernel (fit_split.cuh)')
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
    if a.split:
        names = ['pose features + rest joints', 'prior + forward GEMM rows (thread 0), reduce, broadcast', 'whole forward (slots 0, 1, 12, 13 + waiting for the chain sweep)',
                 '49 output joints', 'projection + GMoF (+ trace)', 'joint backward', 'picked-vertex backward',
                 'backward GEMM rows, reduce, broadcast', 'cluster barrier', 'landing copy + Rodrigues backward + Adam',
                 '  [beside the forward GEMMs] chain forward sweep (first chain thread)', '  [beside the backward GEMM] chain backward sweep',
                 '  cluster barrier after the forward GEMMs (thread 0)', '  prior selection on the GEMM threads']
        tot = sum(buf[i] for i in range(10))
        print('cluster kernel, CTA 0: %.0f cycles per stage-2 iteration (slots 0-2 include the 2 extra forward calls)' % (tot / a.iters))
        for i, n in enumerate(names):
            print('%-75s %9.0f clk/iter' % (n, buf[i] / (a.iters + 2 if i < 3 or i in (10, 12) else a.iters)))
    elif a.pair:
        names = {0: 'pose features + rest joints + B operands', 1: 'cluster barrier before the forward call', 2: 'forward call: generator warp 0 busy',
                 3: 'forward call: wait for the end barrier (gen warp 0)', 4: '49 output joints', 5: 'projection + GMoF', 6: 'joint backward',
                 7: 'picked-vertex backward', 8: 'cluster barrier before the backward call', 9: 'backward call: generator warp 0 busy',
                 10: 'backward call: wait for the end barrier (gen warp 0)', 11: 'Rodrigues backward + Adam',
                 16: '  [forward call] generator warp 4', 17: '  [forward call] epilogue warp 8 (incl. prior select)', 18: '  [forward call] MMA warp',
                 19: '  [forward call] chain warp', 20: '  [backward call] generator warp 4', 21: '  [backward call] epilogue warp 8',
                 22: '  [backward call] MMA warp', 23: '  [backward call] chain warp',
                 24: '  [forward call] generator warp 0: waiting for a ring slot', 25: '  [forward call] generator warp 0: TMEM store + publish',
                 26: '  [forward call] MMA warp: waiting for chunks', 28: '  [backward call] generator warp 0: waiting for a ring slot',
                 29: '  [backward call] generator warp 0: TMEM store + publish', 30: '  [backward call] MMA warp: waiting for chunks'}
        n_fwd = a.iters + 2                                   # the stage-1 prologue and the final forward also make a forward call
        tot = sum(buf[i] for i in range(12))
        print('pair kernel, CTA 0: %.0f cycles per stage-2 iteration (thread 0; forward-call slots include the 2 extra calls)' % (tot / a.iters))
        for i in sorted(names):
            per = a.iters if (i in (4, 5, 6, 7, 11) or 8 <= i <= 10 or 20 <= i <= 23 or i >= 28) else n_fwd
            print('%-58s %9.0f clk/call' % (names[i], buf[i] / per))
    else:
        tot = sum(buf[i] for i in range(12))
        print('stage-2 cycles per iteration (CTA 0): %.0f' % (tot / a.iters))
        for i, n in enumerate(NAMES):
            print('%-42s %9.0f clk/iter  %5.1f%%' % (n, buf[i] / a.iters, 100.0 * buf[i] / tot))
