"""BASELINE.json's full sizes, checked through size-independent properties (the oracle cannot run them in seconds):

config 4  bulk refit of 65 536 samples: keep-if-better never raises a loss, a second refit from the refitted parameters is
          idempotent in its bookkeeping, rows equal the same rows fitted alone, the store update is a pure masked scatter;
config 5  SMPL LBS forward + backward at batch 65 536: rows equal the same rows computed in a small batch, gradients of a
          linear functional are linear in the upstream gradient."""
import numpy as np
import pytest
import torch

from inbed_pose_estimation_b200 import sharded, synthetic

pytestmark = pytest.mark.gpu
N = 65536


def test_bulk_refit_65536_properties():
    fitter = synthetic.build_smplify('cuda', num_iters=3, seed=0)          # 3 + 3 iterations keep the test short
    inp = synthetic.make_fit_inputs(N, seed=4)
    dev = torch.device('cuda')
    fits = torch.from_numpy(np.concatenate([inp['pose'], inp['betas']], axis=1)).to(dev)
    cam, cen, kp = (torch.from_numpy(inp[k]).to(dev) for k in ('cam_t', 'center', 'keypoints'))
    refit = sharded.ShardedRefit(smplify=fitter, device=dev)
    loss0 = fitter.get_fitting_loss(fits[:, :72], fits[:, 72:], cam, cen, kp.clone()).mean(dim=-1)
    f1, l1, up1, cam1 = refit(fits, cam, cen, kp, loss0)
    assert torch.isfinite(f1).all() and torch.isfinite(l1).all()
    assert (l1 <= loss0).all() and torch.equal(l1 < loss0, up1)            # keep-if-better: never worse, update <=> strictly better
    assert torch.equal(f1[~up1], fits[~up1])                               # rows that were not updated are bit-identical
    assert int(up1.sum()) > N // 2                                         # and most fits do improve
    # the same rows fitted alone (other tile sizes, other tile positions) agree
    for lo, hi in ((0, 40), (N - 53, N)):
        sub = fitter(fits[lo:hi, :72].contiguous(), fits[lo:hi, 72:].contiguous(), cam[lo:hi].contiguous(), cen[lo:hi].contiguous(),
                     kp[lo:hi].clone())
        new_loss = sub[5].mean(dim=-1)
        better = new_loss < loss0[lo:hi]
        assert torch.equal(better, up1[lo:hi])
        np.testing.assert_allclose(f1[lo:hi][better, :72].cpu().numpy(), sub[2][better].cpu().numpy(), atol=1e-5)
        np.testing.assert_allclose(l1[lo:hi][better].cpu().numpy(), new_loss[better].cpu().numpy(), rtol=1e-5)
    # deterministic: the whole 65 536-sample refit twice gives identical bits
    f2, l2, up2, _ = refit(fits, cam, cen, kp, loss0)
    assert torch.equal(f1, f2) and torch.equal(l1, l2) and torch.equal(up1, up2)


def test_lbs_65536_rows_and_linearity():
    fitter = synthetic.build_smplify('cuda', num_iters=1, seed=0)
    smpl = fitter.smpl
    inp = synthetic.make_fit_inputs(N, seed=9)
    pose = torch.from_numpy(inp['pose']).cuda()
    betas = torch.from_numpy(inp['betas']).cuda()
    with torch.no_grad():
        big = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
        assert torch.isfinite(big.vertices).all() and torch.isfinite(big.joints).all()
        for lo, hi in ((0, 19), (N - 33, N), (31000, 31017)):
            small = smpl(global_orient=pose[lo:hi, :3], body_pose=pose[lo:hi, 3:], betas=betas[lo:hi])
            # vertices: the tensor-core path gives a sample's row the same bits in any batch; joints: the folded GEMM of
            # small tiles splits its reduction differently (fp32 rounding only)
            assert torch.equal(big.vertices[lo:hi], small.vertices)
            np.testing.assert_allclose(big.joints[lo:hi].cpu().numpy(), small.joints.cpu().numpy(), rtol=0, atol=2e-6)
        del big
    # backward at full size: d/dpose of <g, vertices> is linear in g  (g1, g2 drawn once; grad(g1 + 2 g2) = grad(g1) + 2 grad(g2))
    gen = torch.Generator(device='cuda').manual_seed(5)
    g1 = torch.randn(N, 6890, 3, device='cuda', generator=gen)
    g2 = torch.randn(N, 6890, 3, device='cuda', generator=gen)

    def grads(g):
        p = pose.clone().requires_grad_(True)
        b = betas.clone().requires_grad_(True)
        out = smpl(global_orient=p[:, :3], body_pose=p[:, 3:], betas=b)
        out.vertices.backward(g)
        return p.grad, b.grad
    pa, ba = grads(g1)
    pb, bb = grads(g2)
    g1.add_(g2, alpha=2.)
    pc, bc = grads(g1)
    scale = float(pc.abs().max())
    np.testing.assert_allclose(pc.cpu().numpy(), (pa + 2. * pb).cpu().numpy(), rtol=2e-4, atol=2e-5 * scale)
    np.testing.assert_allclose(bc.cpu().numpy(), (ba + 2. * bb).cpu().numpy(), rtol=2e-4, atol=2e-5 * float(bc.abs().max()))
    assert torch.isfinite(pc).all() and torch.isfinite(bc).all()
