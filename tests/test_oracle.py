"""The CPU oracle (oracle/port.py) replayed against vectors produced by the reference's
own files (oracle/run_reference.py -> tests/golden/).  CPU only."""
import numpy as np
import torch

from conftest import golden
from inbed_pose_estimation_b200 import constants as C
from oracle import port


def _t(a):
    return torch.from_numpy(np.array(a, copy=True))


def test_integer_tables_bit_exact():
    g = golden('tables.npz')
    assert [C.JOINT_MAP[n] for n in C.JOINT_NAMES] == g['joint_map'].tolist()
    assert C.SMPLIFY_IGNORED_JOINTS == g['ign_joints'].tolist() == [1, 9, 12, 27, 28]
    assert C.SMPL_POSE_FLIP_PERM == g['pose_flip_perm'].tolist()
    assert C.J49_FLIP_PERM == g['j49_flip_perm'].tolist()
    assert C.JOINT_NAMES == g['joint_names'].tolist()
    assert C.CAMERA_OP_JOINTS == [9, 12, 2, 5] and C.CAMERA_GT_JOINTS == [27, 28, 33, 34]


def test_known_answers():
    g = golden('kat.npz')
    assert float(port.gmof(torch.tensor(100.), 100)) == float(g['gmof_100_100']) == 5000.0
    assert float(port.angle_prior(torch.zeros(1, 69)).sum() * 15.2 ** 2) == float(g['angle_prior_zero'])
    p = port.project_points(torch.tensor([[[0., 0., 0.], [1., 2., 0.]]]), torch.eye(3)[None],
                            torch.tensor([[0., 0., 10.]]), 5000., torch.tensor([[112., 112.]]))
    assert np.array_equal(p.numpy(), g['proj_simple'])
    assert np.array_equal(p.numpy()[0], np.array([[112., 112.], [612., 1112.]], dtype=np.float32))
    assert np.array_equal(port.quaternion_rodrigues(torch.zeros(2, 3)).numpy(), g['rodrigues_zero'])
    assert np.array_equal(port.exp_map_rodrigues(torch.zeros(2, 3)).numpy(), np.tile(np.eye(3, dtype=np.float32), (2, 1, 1)))


def test_geometry_matches_reference_files():
    g = golden('geometry.npz')
    th = _t(g['theta']).requires_grad_(True)
    R = port.quaternion_rodrigues(th)
    assert np.array_equal(R.detach().numpy(), g['rotmat'])
    (R * _t(g['grad_rotmat'])).sum().backward()
    np.testing.assert_allclose(th.grad.numpy(), g['grad_theta'], rtol=0, atol=1e-6)
    # the two Rodrigues forms agree to rounding
    assert (port.exp_map_rodrigues(th.detach()) - R.detach()).abs().max() < 2e-6
    pr = port.project_points(_t(g['points']), _t(g['rotation']), _t(g['translation']), 5000., _t(g['center']))
    assert np.array_equal(pr.numpy(), g['projected'])


def test_smpl_forward_matches_reference_files(oracle_fp32):
    g = golden('smpl_forward.npz')
    smpl = oracle_fp32.smpl
    pose, betas = _t(g['pose']).requires_grad_(True), _t(g['betas']).requires_grad_(True)
    out = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    s = int(g['vertex_stride'])
    assert np.array_equal(out.vertices.detach().numpy()[:, ::s], g['vertices_sub'])
    assert np.array_equal(out.joints.detach().numpy(), g['joints'])
    rs = np.random.RandomState(int(g['grad_seed']))
    # regenerate the upstream gradients exactly as run_reference drew them
    rs.randn(64, 3); rs.randn(64, 3, 3); rs.randn(8, 49, 3); rs.randn(8, 3); rs.randn(8, 3); rs.randn(8, 2)
    rs.randn(8, 49, 2)
    gv = torch.tensor(rs.randn(4, 6890, 3).astype(np.float32))
    gj = torch.tensor(rs.randn(4, 49, 3).astype(np.float32))
    ((out.vertices * gv).sum() + (out.joints * gj).sum()).backward()
    np.testing.assert_allclose(pose.grad.numpy(), g['grad_pose'], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(betas.grad.numpy(), g['grad_betas'], rtol=1e-5, atol=1e-5)
    out2 = smpl(global_orient=_t(g['rotmats'])[:, :1], body_pose=_t(g['rotmats'])[:, 1:], betas=_t(g['betas']),
                pose2rot=False)
    assert np.array_equal(out2.joints.numpy(), g['joints_rotmat'])
    # zero pose => vertices are the shaped template
    z = smpl(global_orient=torch.zeros(2, 3), body_pose=torch.zeros(2, 69), betas=_t(g['betas'])[:2])
    v_shaped = smpl.v_template + torch.einsum('bl,mkl->bmk', _t(g['betas'])[:2], smpl.shapedirs)
    assert (z.vertices - v_shaped).abs().max() < 1e-6


def test_prior_matches_reference_files(oracle_fp32):
    g = golden('prior.npz')
    prior = oracle_fp32.pose_prior
    assert np.array_equal(prior.nll_weights.numpy(), g['nll_weights'])
    bp = _t(g['body_pose']).requires_grad_(True)
    nll = prior(bp)
    assert np.array_equal(nll.detach().numpy(), g['nll'])
    nll.sum().backward()
    np.testing.assert_allclose(bp.grad.numpy(), g['grad'], rtol=1e-6, atol=1e-6)


def test_fit_matches_reference_files(oracle_fp32):
    """100+100 iterations, B=4: outputs and every per-iteration loss equal the reference run."""
    g = golden('smplify_default.npz')
    kp = _t(g['keypoints'])
    trace = []
    v, j, pose, betas, cam, reproj = oracle_fp32(_t(g['pose']), _t(g['betas']), _t(g['cam_t']),
                                                 _t(g['center']), kp, trace=trace)
    sums = np.array([float(t.sum()) for t in trace])
    np.testing.assert_allclose(sums, g['loss_trace'], rtol=2e-6)
    np.testing.assert_allclose(pose.numpy(), g['out_pose'], atol=2e-5)
    np.testing.assert_allclose(betas.numpy(), g['out_betas'], atol=2e-5)
    np.testing.assert_allclose(cam.detach().numpy(), g['out_cam_t'], atol=2e-5)
    np.testing.assert_allclose(j.numpy(), g['out_joints'], atol=2e-5)
    np.testing.assert_allclose(v.numpy()[:, ::8], g['out_vertices_sub'], atol=2e-5)
    np.testing.assert_allclose(reproj.numpy(), g['out_reproj'], rtol=1e-4, atol=1e-3)
    assert np.all(kp.numpy()[:, C.SMPLIFY_IGNORED_JOINTS, 2] == 0)     # in-place side effect
    kp2 = _t(g['keypoints'])
    fl = oracle_fp32.get_fitting_loss(_t(g['pose']), _t(g['betas']), _t(g['cam_t']), _t(g['center']), kp2)
    np.testing.assert_allclose(fl.numpy(), g['init_fitting_loss'], rtol=1e-6)
    assert np.array_equal(kp2.numpy(), g['keypoints_after_loss'])
