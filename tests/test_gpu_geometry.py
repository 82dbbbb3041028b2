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
