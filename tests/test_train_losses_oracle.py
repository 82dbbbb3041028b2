"""The CPU restatement of the train-step losses / bookkeeping (oracle/adjacent.py) against the vectors
the reference's own train/trainer.py source produced (tests/golden/train_losses.npz, oracle/run_reference_trainer.py)."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import adjacent


@pytest.fixture(scope='module')
def g():
    return golden('train_losses.npz')


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_keypoint_loss(g):
    for tag, (ow, gw) in (('a', (0., 1.)), ('b', (0.5, 2.))):
        p = T(g['kp_pred']).clone().requires_grad_(True)
        loss = adjacent.keypoint_loss(p, T(g['kp_gt']), ow, gw)
        loss.backward()
        assert np.array_equal(loss.detach().numpy(), g['kp_loss_' + tag])
        np.testing.assert_allclose(p.grad.numpy(), g['kp_grad_' + tag], rtol=1e-6, atol=1e-10)


def test_keypoint_3d_loss(g):
    p = T(g['k3_pred']).clone().requires_grad_(True)
    loss = adjacent.keypoint_3d_loss(p, T(g['k3_gt']), T(g['k3_has']))
    loss.backward()
    np.testing.assert_allclose(loss.detach().numpy(), g['k3_loss'], rtol=1e-6)
    np.testing.assert_allclose(p.grad.numpy(), g['k3_grad'], rtol=1e-5, atol=1e-9)
    assert float(adjacent.keypoint_3d_loss(T(g['k3_pred']), T(g['k3_gt']), torch.zeros(24, dtype=torch.uint8))) == 0.
    assert g['k3_loss_none'].shape == (1,) and float(g['k3_loss_none'][0]) == 0.     # the reference's shape-[1] zero


def test_shape_loss(g):
    pv, gv = [T(a) for a in adjacent.golden_vertex_pair(int(g['sh_seed']), 24)]
    assert np.array_equal(pv.numpy()[:, ::689], g['sh_pred_probe']) and np.array_equal(gv.numpy()[:, ::689], g['sh_gt_probe'])
    p = pv.clone().requires_grad_(True)
    loss = adjacent.shape_loss(p, gv, T(g['sh_valid']))
    loss.backward()
    np.testing.assert_allclose(loss.detach().numpy(), g['sh_loss'], rtol=1e-6)
    np.testing.assert_allclose(p.grad.numpy()[:, ::689], g['sh_grad_probe'], rtol=1e-6)
    np.testing.assert_allclose(float(p.grad.abs().double().sum()), float(g['sh_grad_abs_sum']), rtol=1e-9)


def test_smpl_losses(g):
    pr, pb = T(g['sl_pred_rotmat']).clone().requires_grad_(True), T(g['sl_pred_betas']).clone().requires_grad_(True)
    lp, lb = adjacent.smpl_losses(pr, pb, T(g['sl_gt_pose']), T(g['sl_gt_betas']), T(g['sh_valid']))
    (lp + 3. * lb).backward()
    np.testing.assert_allclose(lp.detach().numpy(), g['sl_loss_pose'], rtol=1e-6)
    np.testing.assert_allclose(lb.detach().numpy(), g['sl_loss_betas'], rtol=1e-6)
    np.testing.assert_allclose(pr.grad.numpy(), g['sl_grad_rotmat'], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(pb.grad.numpy(), g['sl_grad_betas_x3'], rtol=1e-5, atol=1e-9)


def test_finalize_fits(g):
    ov, gv = [T(a) for a in adjacent.golden_vertex_pair(int(g['fin_vertex_seed']), 6)]
    assert np.array_equal(ov.numpy()[:, ::689], g['fin_in_opt_vertices_probe'])
    i = lambda k: T(g['fin_in_' + k])
    pose, betas, cam, joints, verts, valid = adjacent.finalize_fits(
        i('opt_pose'), i('opt_betas'), i('opt_cam_t'), i('opt_joints'), ov, i('opt_joint_loss'), i('has_smpl'),
        i('gt_pose'), i('gt_betas'), i('gt_cam_t'), i('gt_model_joints'), gv, 100.)
    for k, v in (('opt_pose', pose), ('opt_betas', betas), ('opt_cam_t', cam), ('opt_joints', joints)):
        assert np.array_equal(v.numpy(), g['fin_out_' + k]), k
    assert np.array_equal(valid.numpy().astype(np.uint8), g['fin_out_valid_fit'])
    for r in range(6):
        assert bool(torch.equal(verts[r], gv[r])) == bool(g['fin_out_vertices_from_gt'][r])
        assert bool(torch.equal(verts[r], ov[r])) == bool(g['fin_out_vertices_kept'][r])
    assert (betas[0] == 0).all() and torch.equal(betas[1], i('gt_betas')[1])       # extreme betas: zeroed / ground truth


def test_weak_perspective_projection(g):
    cam, joints = T(g['wp_cam']).clone().requires_grad_(True), T(g['wp_joints']).clone().requires_grad_(True)
    kp, cam_t = adjacent.weak_perspective_projection(joints, cam, float(g['wp_focal']), float(g['wp_img_res']))
    ((kp * T(g['wp_g_kp'])).sum() + (cam_t * T(g['wp_g_cam_t'])).sum()).backward()
    np.testing.assert_allclose(kp.detach().numpy(), g['wp_kp'], rtol=1e-6, atol=1e-6)
    assert np.array_equal(cam_t.detach().numpy(), g['wp_cam_t'])
    np.testing.assert_allclose(cam.grad.numpy(), g['wp_grad_cam'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(joints.grad.numpy(), g['wp_grad_joints'], rtol=1e-5, atol=1e-6)
