"""GPU parity of batch_rodrigues / perspective_projection (C ABI) with the reference-generated vectors."""
import numpy as np
import pytest
import torch

from conftest import golden
from inbed_pose_estimation_b200 import geometry

pytestmark = pytest.mark.gpu


def test_known_answers():
    k = golden('kat.npz')
    p = geometry.perspective_projection(torch.tensor([[[0., 0., 0.], [1., 2., 0.]]]).cuda(), torch.eye(3)[None].cuda(),
                                        torch.tensor([[0., 0., 10.]]).cuda(), 5000., torch.tensor([[112., 112.]]).cuda())
    assert np.array_equal(p.cpu().numpy(), k['proj_simple'])
    assert np.array_equal(geometry.batch_rodrigues(torch.zeros(2, 3).cuda()).cpu().numpy(), k['rodrigues_zero'])


def test_rodrigues_forward_backward():
    g = golden('geometry.npz')
    th = torch.from_numpy(g['theta']).cuda().requires_grad_(True)
    R = geometry.batch_rodrigues(th)
    np.testing.assert_allclose(R.detach().cpu().numpy(), g['rotmat'], atol=5e-7)
    (R * torch.from_numpy(g['grad_rotmat']).cuda()).sum().backward()
    np.testing.assert_allclose(th.grad.cpu().numpy(), g['grad_theta'], rtol=1e-5, atol=2e-6)


def test_projection_forward_backward():
    g = golden('geometry.npz')
    pts = torch.from_numpy(g['points']).cuda().requires_grad_(True)
    rot = torch.from_numpy(g['rotation']).cuda().requires_grad_(True)
    tr = torch.from_numpy(g['translation']).cuda().requires_grad_(True)
    cen = torch.from_numpy(g['center']).cuda()
    pr = geometry.perspective_projection(pts, rot, tr, 5000., cen)
    np.testing.assert_allclose(pr.detach().cpu().numpy(), g['projected'], rtol=2e-6, atol=2e-4)
    (pr * torch.from_numpy(g['grad_projected']).cuda()).sum().backward()
    np.testing.assert_allclose(pts.grad.cpu().numpy(), g['grad_points'], rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(rot.grad.cpu().numpy(), g['grad_rotation'], rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(tr.grad.cpu().numpy(), g['grad_translation'], rtol=1e-4, atol=1e-2)
    # per-batch focal length
    f = torch.full((8,), 5000.).cuda()
    pr2 = geometry.perspective_projection(pts.detach(), rot.detach(), tr.detach(), f, cen)
    assert torch.equal(pr2, pr.detach())


@pytest.mark.parametrize('tag', ['verts', 'joints'])
def test_projection_out_3d_forward_backward(tag):
    """utils/geometry.py:108-114 (out_3d=True; train/trainer.py:621-626 calls it on the 6890 vertices): third channel = depth."""
    g = golden('geometry_out3d.npz')
    pts = torch.from_numpy(g[tag + '_points']).cuda().requires_grad_(True)
    rot = torch.from_numpy(g[tag + '_rotation']).cuda().requires_grad_(True)
    tr = torch.from_numpy(g[tag + '_translation']).cuda().requires_grad_(True)
    cen = torch.from_numpy(g[tag + '_center']).cuda()
    pr = geometry.perspective_projection(pts, rot, tr, 5000., cen, out_3d=True)
    assert pr.shape == g[tag + '_projected'].shape and pr.shape[-1] == 3
    np.testing.assert_allclose(pr.detach().cpu().numpy(), g[tag + '_projected'], rtol=2e-6, atol=2e-4)
    # the first two channels are the 2-D projection, bit for bit
    assert torch.equal(pr.detach()[..., :2], geometry.perspective_projection(pts.detach(), rot.detach(), tr.detach(), 5000., cen))
    (pr * torch.from_numpy(g[tag + '_grad_projected']).cuda()).sum().backward()
    np.testing.assert_allclose(pts.grad.cpu().numpy(), g[tag + '_grad_points'], rtol=1e-4, atol=1e-3)
    sc = np.abs(g[tag + '_grad_rotation']).max()
    np.testing.assert_allclose(rot.grad.cpu().numpy(), g[tag + '_grad_rotation'], rtol=1e-4, atol=1e-5 * sc)
    sc = np.abs(g[tag + '_grad_translation']).max()
    np.testing.assert_allclose(tr.grad.cpu().numpy(), g[tag + '_grad_translation'], rtol=1e-4, atol=1e-5 * sc)


def test_projection_broadcast_rotation_and_rejected_grads():
    """A rotation given as [3,3] / [1,3,3] (the reference's einsum broadcasts it) gets the gradient summed over the batch;
    intrinsics that require grad are refused rather than silently ignored."""
    g = golden('geometry.npz')
    pts = torch.from_numpy(g['points']).cuda()
    tr = torch.from_numpy(g['translation']).cuda()
    cen = torch.from_numpy(g['center']).cuda()
    gp = torch.from_numpy(g['grad_projected']).cuda()
    R0 = torch.from_numpy(g['rotation'][0]).cuda()
    full = R0[None].expand(8, 3, 3).contiguous().requires_grad_(True)
    (geometry.perspective_projection(pts, full, tr, 5000., cen) * gp).sum().backward()
    for shape in ((3, 3), (1, 3, 3)):
        r = R0.reshape(shape).clone().requires_grad_(True)
        (geometry.perspective_projection(pts, r, tr, 5000., cen) * gp).sum().backward()
        assert r.grad.shape == shape
        np.testing.assert_allclose(r.grad.reshape(3, 3).cpu().numpy(), full.grad.sum(0).cpu().numpy(), rtol=1e-5, atol=1e-3)
    with pytest.raises(NotImplementedError):
        geometry.perspective_projection(pts, full, tr, torch.full((8,), 5000., device='cuda', requires_grad=True), cen)
    with pytest.raises(NotImplementedError):
        geometry.perspective_projection(pts, full, tr, 5000., cen.clone().requires_grad_(True))
