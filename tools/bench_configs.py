#!/usr/bin/env python
"""The BASELINE.json configurations other than the headline one (bench.py) and the LBS sweep (tools/bench_lbs.py):

  1  SMPL forward (LBS), batch 32, fp32, synthetic SMPL-shaped model: CPU oracle beside the CUDA path
  2  SMPLify 100-iteration fit, batch 32 (train.py --batch_size 32 --run_smplify)
  3  cascade refinement: 3 SMPLify calls at batch 256 + 3 rotation-matrix-mode SMPL forward/backward passes
  4  bulk refit of 65 536 samples sharded over the ranks, keep-if-better against a stored fits array, NCCL gather

    python tools/bench_configs.py                          # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_configs.py --configs 4

One JSON line per configuration (rank 0).  CUDA-event timing, max over ranks, after warm-up.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from inbed_pose_estimation_b200 import geometry, sharded, synthetic  # noqa: E402


def cuda_time(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--configs', default='1,2,3,4')
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--bulk', type=int, default=65536)
    ap.add_argument('--no-cpu', action='store_true')
    a = ap.parse_args()
    world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    want = [int(c) for c in a.configs.split(',')]
    fitter = synthetic.build_smplify(dev, num_iters=100, seed=0)
    smpl = fitter.smpl
    out = []

    def inputs(B, seed):
        inp = synthetic.make_fit_inputs(B, seed=seed)
        return [torch.from_numpy(inp[k]).to(dev) for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')]

    if 1 in want and rank == 0:
        pose, betas = inputs(32, 1)[:2]
        with torch.no_grad():
            ms = cuda_time(lambda: smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas), a.reps)
        line = {'config': 1, 'workload': 'SMPL forward (LBS) batch 32 fp32', 'gpu_ms': ms, 'gpu_bodies_per_s': 32 / ms * 1e3}
        if not a.no_cpu:
            import bench                                   # the CPU baseline leg lives in bench.py
            cpu_ms, cores = bench.time_oracle_cpu_smpl_forward(pose.cpu(), betas.cpu())
            line.update({'cpu_ms': cpu_ms, 'cpu_bodies_per_s': 32 / cpu_ms * 1e3, 'cpu_cores': cores, 'cpu_kind': 'port'})
        out.append(line)

    if 2 in want and rank == 0:
        x = inputs(32, 2)
        ms = cuda_time(lambda: fitter(x[0], x[1], x[2], x[3], x[4].clone()), a.reps)
        out.append({'config': 2, 'workload': 'SMPLify 100+100 iterations, batch 32', 'gpu_ms': ms, 'fits_per_s': 32 / ms * 1e3})

    if 3 in want and rank == 0:
        xs = [inputs(256, 30 + i) for i in range(3)]
        rots = [geometry.batch_rodrigues(x[0].reshape(-1, 3)).view(256, 24, 3, 3) for x in xs]

        def cascade():
            for x, R in zip(xs, rots):
                Rg = R.clone().requires_grad_(True)
                o = smpl(global_orient=Rg[:, :1], body_pose=Rg[:, 1:], betas=x[1], pose2rot=False)     # trainer.py:597, with autograd
                (o.vertices.square().sum() + o.joints.square().sum()).backward()
                fitter(x[0], x[1], x[2], x[3], x[4].clone())
        ms = cuda_time(cascade, a.reps)
        out.append({'config': 3, 'workload': '3 x (rotmat-mode SMPL fwd+bwd + SMPLify 100+100) at batch 256', 'gpu_ms': ms,
                    'fits_per_s': 3 * 256 / ms * 1e3})

    if 4 in want:
        N = a.bulk
        inp = synthetic.make_fit_inputs(N, seed=4)
        fits = torch.from_numpy(np.concatenate([inp['pose'], inp['betas']], axis=1)).to(dev)
        cam, cen, kp = (torch.from_numpy(inp[k]).to(dev) for k in ('cam_t', 'center', 'keypoints'))
        old_loss = torch.full((N,), 1e9, device=dev)
        refit = sharded.ShardedRefit(smplify=fitter, device=dev)
        res = {}

        def bulk():
            res['out'] = refit(fits, cam, cen, kp, old_loss)
        for _ in range(2):
            bulk()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = []
        for _ in range(max(2, a.reps // 2)):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            bulk()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ms))], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        f, loss, upd, _ = res['out']
        if rank == 0:
            out.append({'config': 4, 'workload': 'bulk refit of %d samples, keep-if-better, gather of [N,134]' % N, 'n_gpus': world,
                        'ms': float(t.item()), 'fits_per_s': N / float(t.item()) * 1e3, 'updated': int(upd.sum()),
                        'mean_loss_after': float(loss.mean())})
    if rank == 0:
        for line in out:
            print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
