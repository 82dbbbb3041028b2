"""GPU parity of SMPL.forward / backward (C ABI) against the oracle and golden vectors."""
import numpy as np
import pytest
import torch

from conftest import golden
from inbed_pose_estimation_b200 import synthetic
from inbed_pose_estimation_b200.smpl import SMPL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def smpl():
    return SMPL(model_arrays=synthetic.model_arrays(0), j_regressor_extra=synthetic.make_extra_regressor(1)).cuda()


@pytest.fixture(scope='module')
def oracle64():
    from oracle import port
    return port.build_oracle(seed=0, dtype=torch.float64).smpl


def test_forward_matches_reference_golden(smpl):
    g = golden('smpl_forward.npz')
    pose, betas = torch.from_numpy(g['pose']).cuda(), torch.from_numpy(g['betas']).cuda()
    out = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    np.testing.assert_allclose(out.vertices.cpu().numpy()[:, ::8], g['vertices_sub'], atol=5e-6)
    np.testing.assert_allclose(out.joints.cpu().numpy(), g['joints'], atol=5e-6)
    np.testing.assert_allclose(out.vertices.double().sum(dim=1).cpu().numpy(), g['vertices_checksum'], atol=2e-3)
    R = torch.from_numpy(g['rotmats']).cuda()
    out2 = smpl(global_orient=R[:, :1], body_pose=R[:, 1:], betas=betas, pose2rot=False, return_full_pose=True)
    np.testing.assert_allclose(out2.vertices.cpu().numpy()[:, ::8], g['vertices_rotmat_sub'], atol=5e-6)
    np.testing.assert_allclose(out2.joints.cpu().numpy(), g['joints_rotmat'], atol=5e-6)
    assert out2.full_pose.shape == (4, 24, 3, 3) and out.full_pose is None


def test_backward_matches_reference_golden(smpl):
    g = golden('smpl_forward.npz')
    rs = np.random.RandomState(int(g['grad_seed']))
    rs.randn(64, 3); rs.randn(64, 3, 3); rs.randn(8, 49, 3); rs.randn(8, 3); rs.randn(8, 3); rs.randn(8, 2); rs.randn(8, 49, 2)
    gv = torch.tensor(rs.randn(4, 6890, 3).astype(np.float32)).cuda()
    gj = torch.tensor(rs.randn(4, 49, 3).astype(np.float32)).cuda()
    pose = torch.from_numpy(g['pose']).cuda().requires_grad_(True)
    betas = torch.from_numpy(g['betas']).cuda().requires_grad_(True)
    out = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    ((out.vertices * gv).sum() + (out.joints * gj).sum()).backward()
    np.testing.assert_allclose(pose.grad.cpu().numpy(), g['grad_pose'], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(betas.grad.cpu().numpy(), g['grad_betas'], rtol=1e-4, atol=2e-4)
    R = torch.from_numpy(g['rotmats']).cuda().requires_grad_(True)
    b2 = torch.from_numpy(g['betas']).cuda().requires_grad_(True)
    out2 = smpl(global_orient=R[:, :1], body_pose=R[:, 1:], betas=b2, pose2rot=False)
    ((out2.vertices * gv).sum() + (out2.joints * gj).sum()).backward()
    np.testing.assert_allclose(R.grad.cpu().numpy(), g['grad_rotmats'], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(b2.grad.cpu().numpy(), g['grad_betas_rotmat'], rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize('batch', [1, 37, 300])
def test_forward_backward_vs_oracle_ragged(smpl, oracle64, batch):
    rs = np.random.RandomState(batch)
    inp = synthetic.make_fit_inputs(batch, seed=batch)
    gv = rs.randn(batch, 6890, 3)
    gj = rs.randn(batch, 49, 3)
    p64 = torch.tensor(inp['pose'], dtype=torch.float64, requires_grad=True)
    b64 = torch.tensor(inp['betas'], dtype=torch.float64, requires_grad=True)
    o = oracle64(global_orient=p64[:, :3], body_pose=p64[:, 3:], betas=b64)
    ((o.vertices * torch.tensor(gv)).sum() + (o.joints * torch.tensor(gj)).sum()).backward()
    pose = torch.from_numpy(inp['pose']).cuda().requires_grad_(True)
    betas = torch.from_numpy(inp['betas']).cuda().requires_grad_(True)
    out = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    ((out.vertices * torch.tensor(gv, dtype=torch.float32).cuda()).sum() +
     (out.joints * torch.tensor(gj, dtype=torch.float32).cuda()).sum()).backward()
    np.testing.assert_allclose(out.vertices.detach().cpu().numpy(), o.vertices.detach().numpy(), atol=1e-5)
    np.testing.assert_allclose(out.joints.detach().cpu().numpy(), o.joints.detach().numpy(), atol=1e-5)
    scale = np.abs(p64.grad.numpy()).max()
    np.testing.assert_allclose(pose.grad.cpu().numpy(), p64.grad.numpy(), rtol=1e-4, atol=1e-5 * scale)
    np.testing.assert_allclose(betas.grad.cpu().numpy(), b64.grad.numpy(), rtol=1e-4, atol=1e-5 * np.abs(b64.grad.numpy()).max())


def test_joints_only_grad_and_no_grad(smpl, oracle64):
    inp = synthetic.make_fit_inputs(9, seed=2)
    pose = torch.from_numpy(inp['pose']).cuda().requires_grad_(True)
    betas = torch.from_numpy(inp['betas']).cuda()
    out = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    out.joints.square().sum().backward()
    p64 = torch.tensor(inp['pose'], dtype=torch.float64, requires_grad=True)
    o = oracle64(global_orient=p64[:, :3], body_pose=p64[:, 3:], betas=torch.tensor(inp['betas'], dtype=torch.float64))
    o.joints.square().sum().backward()
    np.testing.assert_allclose(pose.grad.cpu().numpy(), p64.grad.numpy(), rtol=1e-4, atol=1e-4)
    # a joints-only loss must not run the tcgen05 vertex backward on a zero gradient: one launch (the pose backward kernel)
    from inbed_pose_estimation_b200 import _native
    p2 = torch.from_numpy(inp['pose']).cuda().requires_grad_(True)
    out = smpl(global_orient=p2[:, :3], body_pose=p2[:, 3:], betas=betas)
    loss = out.joints.square().sum()
    _native.lib().smplb200_launch_count(1)
    loss.backward()
    assert _native.lib().smplb200_launch_count(0) == 1
    assert torch.equal(p2.grad, pose.grad)
    # neither output used: zero gradient, no launch at all
    p3 = torch.from_numpy(inp['pose']).cuda().requires_grad_(True)
    out = smpl(global_orient=p3[:, :3], body_pose=p3[:, 3:], betas=betas)
    (out.joints.sum() * 0 + p3.sum()).backward()
    with torch.no_grad():
        out = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    assert not out.vertices.requires_grad


def test_zero_pose_and_errors(smpl):
    betas = torch.from_numpy(synthetic.make_fit_inputs(3, seed=1)['betas']).cuda()
    out = smpl(global_orient=torch.zeros(3, 3).cuda(), body_pose=torch.zeros(3, 69).cuda(), betas=betas)
    arrays = synthetic.model_arrays(0)
    v_shaped = arrays['v_template'][None] + np.einsum('bl,vcl->bvc', betas.cpu().numpy(), arrays['shapedirs'])
    np.testing.assert_allclose(out.vertices.cpu().numpy(), v_shaped, atol=2e-6)
    out0 = smpl(global_orient=torch.zeros(0, 3).cuda(), body_pose=torch.zeros(0, 69).cuda(), betas=torch.zeros(0, 10).cuda())
    assert out0.vertices.shape == (0, 6890, 3) and out0.joints.shape == (0, 49, 3)
    with pytest.raises(RuntimeError, match='CUDA'):
        smpl(global_orient=torch.zeros(1, 3), body_pose=torch.zeros(1, 69), betas=torch.zeros(1, 10))


def test_large_batch_rows_equal_small_batch_rows(smpl):
    """Batch independence at a size that crosses the 16-sample pose tiles, several 128-sample MMA tiles and K splits:
    the first rows of a 2100-sample call equal the same rows computed alone (vertices bit for bit - a sample's MMA row
    does not depend on its neighbours - joints and gradients up to the different split-K partitions)."""
    inp = synthetic.make_fit_inputs(2100, seed=21)
    gen = torch.Generator().manual_seed(3)
    gv = torch.randn(2100, 6890, 3, generator=gen).cuda()
    outs = []
    for n in (2100, 37):
        pose = torch.from_numpy(inp['pose'][:n]).cuda().requires_grad_(True)
        betas = torch.from_numpy(inp['betas'][:n]).cuda().requires_grad_(True)
        o = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
        (o.vertices * gv[:n]).sum().backward()
        outs.append((o.vertices.detach()[:37].cpu(), o.joints.detach()[:37].cpu(), pose.grad[:37].cpu(), betas.grad[:37].cpu()))
    assert torch.equal(outs[0][0], outs[1][0])
    np.testing.assert_allclose(outs[0][1].numpy(), outs[1][1].numpy(), rtol=0, atol=2e-6)    # joints: small tiles split the folded GEMM's reduction
    np.testing.assert_allclose(outs[0][2].numpy(), outs[1][2].numpy(), rtol=2e-5, atol=2e-5 * float(outs[1][2].abs().max()))
    np.testing.assert_allclose(outs[0][3].numpy(), outs[1][3].numpy(), rtol=2e-5, atol=2e-5 * float(outs[1][3].abs().max()))
    assert torch.isfinite(outs[0][2]).all()
