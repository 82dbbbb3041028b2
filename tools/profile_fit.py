"""Runs the fit a few times at a given batch (for ncu / timing)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inbed_pose_estimation_b200 import synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--batch', type=int, default=4096)
ap.add_argument('--iters', type=int, default=100)
ap.add_argument('--reps', type=int, default=3)
a = ap.parse_args()
fitter = synthetic.build_smplify('cuda', num_iters=a.iters, seed=0)
inp = synthetic.make_fit_inputs(a.batch, seed=1)
args = [torch.from_numpy(inp[k]).cuda() for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')]
for r in range(a.reps):
    kp = args[4].clone()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fitter(args[0], args[1], args[2], args[3], kp)
    e1.record()
    torch.cuda.synchronize()
    print('rep %d: %.3f ms  (%.0f fits/s)' % (r, e0.elapsed_time(e1), a.batch / e0.elapsed_time(e1) * 1e3))
