"""Times the oracle port (the reference's eager-PyTorch op sequence) ON THE GPU - the
"reference single-GPU PyTorch SMPLify" denominator of BASELINE.json's >= 100x target.
TEST/BASELINE INFRASTRUCTURE ONLY; run by hand:  python -m oracle.time_torch_gpu --batch 32 256 4096
"""
import argparse
import json
import time

import torch

from inbed_pose_estimation_b200 import synthetic
from oracle import port


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, nargs='+', default=[32, 256, 4096])
    ap.add_argument('--iters', type=int, default=100)
    a = ap.parse_args()
    dev = torch.device('cuda')
    oracle = port.build_oracle(seed=0, num_iters=a.iters)
    oracle.smpl = oracle.smpl.to(dev)
    oracle.pose_prior = oracle.pose_prior.to(dev)
    for B in a.batch:
        inp = synthetic.make_fit_inputs(B, seed=7)
        args = lambda: [torch.from_numpy(inp[k].copy()).to(dev) for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')]
        try:
            oracle(*args())                      # warm-up (cuBLAS handles, allocator)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            oracle(*args())
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(json.dumps({'impl': 'oracle port, eager torch on cuda', 'batch': B, 'iters': a.iters,
                              'seconds_per_call': dt, 'fits_per_sec': B / dt,
                              'peak_mem_gb': torch.cuda.max_memory_allocated() / 2 ** 30}))
        except RuntimeError as e:
            print(json.dumps({'batch': B, 'error': str(e)[:200]}))


if __name__ == '__main__':
    main()
