"""Event timeline of one forward GEMM call of the pair kernel (CTA 0), from a profiling build with -DPG_TRACE.
    python tools/pg_trace.py build ; python tools/pg_trace.py run [--batch 2368]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'inbed_pose_estimation_b200', 'libsmplify_b200_trace.so')
if sys.argv[1] == 'build':
    from inbed_pose_estimation_b200 import _native
    print(_native.build(force=True, extra_flags=['-DPG_TRACE'] + sys.argv[2:], out=OUT))
else:
    os.environ['SMPLB200_LIB'] = OUT
    import torch
    from inbed_pose_estimation_b200 import _native, synthetic
    batch = int(sys.argv[sys.argv.index('--batch') + 1]) if '--batch' in sys.argv else 2368
    lib = _native.lib()
    fitter = synthetic.build_smplify('cuda', num_iters=12, seed=0)
    inp = synthetic.make_fit_inputs(batch, seed=1)
    args = [torch.from_numpy(inp[k]).cuda() for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')]
    for _ in range(2):
        fitter(args[0], args[1], args[2], args[3], args[4].clone())
    torch.cuda.synchronize()
    buf = (ctypes.c_longlong * 512)()
    lib.smplb200_debug_pg_trace(buf)
    ev = [[buf[e * 64 + c] for c in range(64)] for e in range(8)]
    t0 = min(v for v in ev[0][:30] + ev[3][:30] if v > 0)
    rel = lambda v: (v - t0) if v > 0 else -1
    print('chunk | generator: start  slot-free  published | MMA warp: wait-begin  chunk-ready  issued')
    for c in range(30):
        print('%5d | %9d %9d %9d | %9d %9d %9d' % (c, rel(ev[0][c]), rel(ev[1][c]), rel(ev[2][c]), rel(ev[3][c]), rel(ev[4][c]), rel(ev[5][c])))
    names = {0: 'prior tile 0 ready', 1: 'prior tile 1 ready', 2: 'prior tile 2 ready', 8: 'Pd written', 9: 'pair sync 1 done', 10: 'select done',
             11: 'pair sync 2 done', 3: 'fwd tile 0 ready', 4: 'fwd tile 1 ready', 5: 'fwd tile 2 ready', 12: 'Q written'}
    for k in (0, 1, 2, 8, 9, 10, 11, 3, 4, 5, 12):
        print('epilogue warp 8: %-20s %9d' % (names[k], rel(ev[6][k])))
