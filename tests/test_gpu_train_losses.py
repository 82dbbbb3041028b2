"""GPU parity of the train-step losses and bookkeeping - through the C ABI - with the vectors the
reference's own train/trainer.py source produced (tests/golden/train_losses.npz) and with the CPU restatement on larger
seeded inputs.  Tolerances: losses 1e-6 relative (fp32 values, double accumulation here vs torch's fp32 tree), gradients
1e-6 relative; masks, overwrites and valid_fit bit-exact."""
import numpy as np
import pytest
import torch

from conftest import golden
from inbed_pose_estimation_b200 import train_losses as TL
from oracle import adjacent

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def g():
    return golden('train_losses.npz')


def C(a):
    return torch.from_numpy(np.asarray(a)).cuda()


def test_keypoint_loss(g):
    for tag, (ow, gw) in (('a', (0., 1.)), ('b', (0.5, 2.))):
        p = C(g['kp_pred']).requires_grad_(True)
        loss = TL.keypoint_loss(p, C(g['kp_gt']), ow, gw)
        assert loss.dim() == 0
        (2. * loss).backward()
        np.testing.assert_allclose(loss.item(), float(g['kp_loss_' + tag]), rtol=1e-6)
        np.testing.assert_allclose(p.grad.cpu().numpy(), 2. * g['kp_grad_' + tag], rtol=2e-6, atol=1e-10)


def test_keypoint_3d_loss(g):
    p = C(g['k3_pred']).requires_grad_(True)
    loss = TL.keypoint_3d_loss(p, C(g['k3_gt']), C(g['k3_has']))
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g['k3_loss']), rtol=1e-6)
    np.testing.assert_allclose(p.grad.cpu().numpy(), g['k3_grad'], rtol=1e-5, atol=1e-9)
    assert (p.grad[:, :25] == 0).all()
    p2 = C(g['k3_pred']).requires_grad_(True)
    none = TL.keypoint_3d_loss(p2, C(g['k3_gt']), torch.zeros(24, dtype=torch.uint8).cuda())
    none.backward()
    assert none.item() == 0. and (p2.grad == 0).all()


def test_shape_loss(g):
    pv, gv = [C(a) for a in adjacent.golden_vertex_pair(int(g['sh_seed']), 24)]
    p = pv.clone().requires_grad_(True)
    loss = TL.shape_loss(p, gv, C(g['sh_valid']))
    loss.backward()
    np.testing.assert_allclose(loss.item(), float(g['sh_loss']), rtol=1e-6)
    grad = p.grad.cpu()
    np.testing.assert_allclose(grad.numpy()[:, ::689], g['sh_grad_probe'], rtol=1e-6)
    np.testing.assert_allclose(float(grad.abs().double().sum()), float(g['sh_grad_abs_sum']), rtol=1e-7)
    assert (grad[0, :100] == 0).all()                                  # sign(0) = 0 at the kink
    with torch.no_grad():
        assert TL.shape_loss(pv, gv, C(g['sh_valid'])).item() == loss.item()   # the no-gradient path reads the same sums


def test_smpl_losses(g):
    pr, pb = C(g['sl_pred_rotmat']).requires_grad_(True), C(g['sl_pred_betas']).requires_grad_(True)
    lp, lb = TL.smpl_losses(pr, pb, C(g['sl_gt_pose']), C(g['sl_gt_betas']), C(g['sh_valid']))
    (lp + 3. * lb).backward()
    np.testing.assert_allclose(lp.item(), float(g['sl_loss_pose']), rtol=1e-6)
    np.testing.assert_allclose(lb.item(), float(g['sl_loss_betas']), rtol=1e-6)
    np.testing.assert_allclose(pr.grad.cpu().numpy(), g['sl_grad_rotmat'], rtol=1e-5, atol=2e-9)
    np.testing.assert_allclose(pb.grad.cpu().numpy(), g['sl_grad_betas_x3'], rtol=1e-5, atol=1e-9)
    # only the pose loss is used: the betas prediction gets no gradient object at all
    pr2, pb2 = C(g['sl_pred_rotmat']).requires_grad_(True), C(g['sl_pred_betas']).requires_grad_(True)
    TL.smpl_losses(pr2, pb2, C(g['sl_gt_pose']).view(24, 24, 3), C(g['sl_gt_betas']), C(g['sh_valid']))[0].backward()
    assert pb2.grad is None and torch.allclose(pr2.grad, pr.grad)


def test_finalize_fits(g):
    ov, gv = [C(a) for a in adjacent.golden_vertex_pair(int(g['fin_vertex_seed']), 6)]
    i = lambda k: C(g['fin_in_' + k])
    pose, betas, cam, joints, verts = i('opt_pose'), i('opt_betas'), i('opt_cam_t'), i('opt_joints'), ov.clone()
    valid = TL.finalize_fits_(pose, betas, cam, joints, verts, i('opt_joint_loss'), i('has_smpl'),
                              i('gt_pose'), i('gt_betas'), i('gt_cam_t'), i('gt_model_joints'), gv, smplify_threshold=100.)
    for k, v in (('opt_pose', pose), ('opt_betas', betas), ('opt_cam_t', cam), ('opt_joints', joints)):
        assert np.array_equal(v.cpu().numpy(), g['fin_out_' + k]), k
    assert valid.dtype == torch.bool and np.array_equal(valid.cpu().numpy().astype(np.uint8), g['fin_out_valid_fit'])
    for r in range(6):
        assert bool(torch.equal(verts[r], gv[r])) == bool(g['fin_out_vertices_from_gt'][r])
        assert bool(torch.equal(verts[r], ov[r])) == bool(g['fin_out_vertices_kept'][r])
    # without vertices; empty batch
    pose2, betas2, cam2, joints2 = i('opt_pose'), i('opt_betas'), i('opt_cam_t'), i('opt_joints')
    v2 = TL.finalize_fits_(pose2, betas2, cam2, joints2, None, i('opt_joint_loss'), i('has_smpl'),
                           i('gt_pose'), i('gt_betas'), i('gt_cam_t'), i('gt_model_joints'), None)
    assert torch.equal(v2, valid) and torch.equal(betas2, betas)
    z = lambda *s: torch.zeros(*s).cuda()
    assert TL.finalize_fits_(z(0, 72), z(0, 10), z(0, 3), z(0, 49, 3), None, z(0), z(0), z(0, 72), z(0, 10), z(0, 3),
                             z(0, 49, 3), None).shape == (0,)


@pytest.mark.parametrize('B', [1, 333, 4096])
def test_against_restatement_at_size(B):
    gen = torch.Generator().manual_seed(B)
    r = lambda *s: torch.randn(*s, generator=gen)
    mask = (torch.rand(B, generator=gen) < 0.6).to(torch.uint8)
    if B == 1:
        mask[:] = 1
    # smpl_losses
    pr, pb, gp, gb = r(B, 24, 3, 3), r(B, 10), 0.5 * r(B, 72), r(B, 10)
    a, b = pr.clone().requires_grad_(True), pb.clone().requires_grad_(True)
    lp, lb = adjacent.smpl_losses(a, b, gp, gb, mask)
    (lp + lb).backward()
    ac, bc = pr.cuda().requires_grad_(True), pb.cuda().requires_grad_(True)
    lpc, lbc = TL.smpl_losses(ac, bc, gp.cuda(), gb.cuda(), mask.cuda())
    (lpc + lbc).backward()
    np.testing.assert_allclose([lpc.item(), lbc.item()], [lp.item(), lb.item()], rtol=2e-6)
    np.testing.assert_allclose(ac.grad.cpu().numpy(), a.grad.numpy(), rtol=2e-5, atol=1e-6 * float(a.grad.abs().max()))
    np.testing.assert_allclose(bc.grad.cpu().numpy(), b.grad.numpy(), rtol=2e-5, atol=1e-6 * float(b.grad.abs().max()))
    # keypoint losses
    kp, gk = r(B, 49, 2), torch.cat([r(B, 49, 2), torch.rand(B, 49, 1, generator=gen)], dim=-1)
    a = kp.clone().requires_grad_(True)
    l = adjacent.keypoint_loss(a, gk, 0.3, 1.0)
    l.backward()
    ac = kp.cuda().requires_grad_(True)
    lc = TL.keypoint_loss(ac, gk.cuda(), 0.3, 1.0)
    lc.backward()
    np.testing.assert_allclose(lc.item(), l.item(), rtol=2e-6)
    np.testing.assert_allclose(ac.grad.cpu().numpy(), a.grad.numpy(), rtol=2e-5, atol=1e-6 * float(a.grad.abs().max()))
    pj, gj = r(B, 49, 3), torch.cat([r(B, 24, 3), torch.rand(B, 24, 1, generator=gen)], dim=-1)
    a = pj.clone().requires_grad_(True)
    l = adjacent.keypoint_3d_loss(a, gj, mask)
    l.backward()
    ac = pj.cuda().requires_grad_(True)
    lc = TL.keypoint_3d_loss(ac, gj.cuda(), mask.cuda())
    lc.backward()
    np.testing.assert_allclose(lc.item(), l.item(), rtol=2e-6)
    np.testing.assert_allclose(ac.grad.cpu().numpy(), a.grad.numpy(), rtol=2e-5, atol=1e-6 * float(a.grad.abs().max()))
    # shape loss (bounded: 4096 x 83 KB x 2 on the CPU side is fine, but keep the restatement quick)
    Bs = min(B, 512)
    pv, gv = r(Bs, 6890, 3), r(Bs, 6890, 3)
    a = pv.clone().requires_grad_(True)
    l = adjacent.shape_loss(a, gv, mask[:Bs])
    l.backward()
    ac = pv.cuda().requires_grad_(True)
    lc = TL.shape_loss(ac, gv.cuda(), mask[:Bs].cuda())
    lc.backward()
    np.testing.assert_allclose(lc.item(), l.item(), rtol=2e-6)
    np.testing.assert_allclose(ac.grad.cpu().numpy(), a.grad.numpy(), rtol=1e-5)
    # finalize
    opt = [r(B, 72), 1.7 * r(B, 10), r(B, 3), r(B, 49, 3)]
    gts = [r(B, 72), r(B, 10), r(B, 3), r(B, 49, 3)]
    loss = 200. * torch.rand(B, generator=gen)
    ref = adjacent.finalize_fits(opt[0], opt[1], opt[2], opt[3], None, loss, mask, gts[0], gts[1], gts[2], gts[3], None, 100.)
    dev = [t.cuda() for t in opt]
    valid = TL.finalize_fits_(dev[0], dev[1], dev[2], dev[3], None, loss.cuda(), mask.cuda(), gts[0].cuda(), gts[1].cuda(),
                              gts[2].cuda(), gts[3].cuda(), None, 100.)
    for d, e in zip(dev, ref[:4]):
        assert torch.equal(d.cpu(), e)
    assert torch.equal(valid.cpu(), ref[5])


def test_errors():
    with pytest.raises(RuntimeError, match='CUDA'):
        TL.keypoint_loss(torch.zeros(2, 49, 2), torch.zeros(2, 49, 3), 0., 1.)
    with pytest.raises(ValueError):
        TL.shape_loss(torch.zeros(2, 6890, 3).cuda(), torch.zeros(3, 6890, 3).cuda(), torch.ones(2).cuda())
    with pytest.raises(ValueError):
        TL.smpl_losses(torch.zeros(2, 24, 3, 3).cuda(), torch.zeros(2, 10).cuda(), torch.zeros(2, 72).cuda(), torch.zeros(2, 10).cuda(),
                       torch.ones(3).cuda())


def test_weak_perspective_projection(g):
    """trainer.py:187-199 in one kernel: values bit-close to the reference's eager ops, gradients to 1e-5."""
    from inbed_pose_estimation_b200 import geometry
    cam, joints = C(g['wp_cam']).requires_grad_(True), C(g['wp_joints']).requires_grad_(True)
    kp, cam_t = geometry.weak_perspective_projection(joints, cam, float(g['wp_focal']), float(g['wp_img_res']))
    ((kp * C(g['wp_g_kp'])).sum() + (cam_t * C(g['wp_g_cam_t'])).sum()).backward()
    np.testing.assert_allclose(kp.detach().cpu().numpy(), g['wp_kp'], rtol=1e-6, atol=1e-6)
    assert np.array_equal(cam_t.detach().cpu().numpy(), g['wp_cam_t'])
    np.testing.assert_allclose(cam.grad.cpu().numpy(), g['wp_grad_cam'], rtol=1e-5, atol=1e-5 * float(np.abs(g['wp_grad_cam']).max()))
    np.testing.assert_allclose(joints.grad.cpu().numpy(), g['wp_grad_joints'], rtol=1e-5, atol=1e-6)
    # only the keypoints are used downstream; cam_t feeds SMPLify detached
    cam2, joints2 = C(g['wp_cam']).requires_grad_(True), C(g['wp_joints']).requires_grad_(True)
    kp2, t2 = geometry.weak_perspective_projection(joints2, cam2, float(g['wp_focal']), float(g['wp_img_res']))
    kp2.sum().backward()
    assert torch.isfinite(cam2.grad).all() and not t2.detach().requires_grad
    big = torch.randn(4096, 49, 3, generator=torch.Generator().manual_seed(1))
    bc = torch.cat([0.5 + torch.rand(4096, 1, generator=torch.Generator().manual_seed(2)), 0.1 * torch.randn(4096, 2, generator=torch.Generator().manual_seed(3))], dim=1)
    k_ref, t_ref = adjacent.weak_perspective_projection(big, bc)
    k_gpu, t_gpu = geometry.weak_perspective_projection(big.cuda(), bc.cuda())
    np.testing.assert_allclose(k_gpu.cpu().numpy(), k_ref.numpy(), rtol=2e-6, atol=2e-6)
    assert torch.equal(t_gpu.cpu(), t_ref)
    assert geometry.weak_perspective_projection(torch.zeros(0, 49, 3).cuda(), torch.zeros(0, 3).cuda())[0].shape == (0, 49, 2)
