"""Host-side logic of the sharded refit (shard bounds, ragged all-gather, keep-if-better) with the
gloo backend, world size 2, on CPU.  The per-rank fit is a stand-in (2 oracle iterations) - this test
is about the plumbing; the CUDA fit itself is covered by test_gpu_smplify.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from inbed_pose_estimation_b200 import sharded, synthetic


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            spans = [sharded.shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_keep_if_better():
    old = torch.arange(12.).view(4, 3)
    new = -old
    fits, loss, upd = sharded.keep_if_better(old, torch.tensor([1., 5., 2., 9.]), new, torch.tensor([2., 4., 2., 1.]))
    assert upd.tolist() == [False, True, False, True]
    assert torch.equal(fits[1], new[1]) and torch.equal(fits[0], old[0]) and loss.tolist() == [1., 4., 2., 1.]


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    from oracle import port as oracle_port
    oracle = oracle_port.build_oracle(seed=0, num_iters=2)
    inp = synthetic.make_fit_inputs(n, seed=9)
    fits = torch.from_numpy(np.concatenate([inp['pose'], inp['betas']], axis=1))
    refit = sharded.ShardedRefit(fit_fn=oracle, device=torch.device('cpu'))
    old_loss = torch.tensor([1e9 if i % 2 == 0 else 0. for i in range(n)])
    out = refit(fits, torch.from_numpy(inp['cam_t']), torch.from_numpy(inp['center']),
                torch.from_numpy(inp['keypoints']), old_loss)
    ret[rank] = [t.numpy() for t in out]
    dist.destroy_process_group()


def test_two_rank_refit_equals_single_process():
    n = 7                                             # ragged: 4 + 3
    from oracle import port as oracle_port
    oracle = oracle_port.build_oracle(seed=0, num_iters=2)
    inp = synthetic.make_fit_inputs(n, seed=9)
    fits = torch.from_numpy(np.concatenate([inp['pose'], inp['betas']], axis=1))
    single = sharded.ShardedRefit(fit_fn=oracle, device=torch.device('cpu'))(
        fits, torch.from_numpy(inp['cam_t']), torch.from_numpy(inp['center']), torch.from_numpy(inp['keypoints']),
        torch.tensor([1e9 if i % 2 == 0 else 0. for i in range(n)]))
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), n, ret), nprocs=2, join=True)
    for rank in (0, 1):
        for a, b in zip(ret[rank], single):
            np.testing.assert_allclose(a, b.numpy(), rtol=1e-6, atol=1e-6)
    assert ret[0][2].dtype == bool and ret[0][2].tolist() == [i % 2 == 0 for i in range(n)]
