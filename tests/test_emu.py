"""Host emulation of the CUDA tile code (tests/emu) against the oracle - CPU only.

This checks the arithmetic the GPU kernels execute (the phases are the same source,
compiled for the host): folded joint model, hand-derived backward pass, priors, Adam,
two-stage control flow.  The GPU parity tests proper are in test_gpu_*.py."""
import numpy as np
import pytest
import torch

from conftest import golden
from inbed_pose_estimation_b200 import constants as C, synthetic
from emu import emu as emu_mod

pytestmark = pytest.mark.skipif(not emu_mod.available(), reason='nvcc needed to build the emulation library')


@pytest.fixture(scope='module')
def emu():
    return emu_mod.Emu(seed=0)


def test_forward_joints_and_transforms(emu, oracle_fp32):
    g = golden('smpl_forward.npz')
    joints, A, x = emu.pose_forward(g['pose'], g['betas'])
    np.testing.assert_allclose(joints, g['joints'], atol=2e-6)
    joints_r, A_r, _ = emu.pose_forward(g['rotmats'].reshape(4, 216), g['betas'], rotmat_mode=True)
    np.testing.assert_allclose(joints_r, g['joints_rotmat'], atol=2e-6)
    # A and x reproduce the oracle's vertices through plain numpy skinning
    arrays = emu_mod.model_arrays(0)
    basis = emu.array('basis').reshape(224, -1)[:218, :20670]
    vp = (x[:, :218].astype(np.float64) @ basis.astype(np.float64)).reshape(4, 6890, 3)
    T = np.einsum('vj,bje->bve', arrays['weights'].astype(np.float64), A.reshape(4, 24, 12).astype(np.float64)).reshape(4, 6890, 3, 4)
    verts = np.einsum('bvrc,bvc->bvr', T[..., :3], vp) + T[..., 3]
    np.testing.assert_allclose(verts[:, ::8], g['vertices_sub'], atol=3e-6)


def test_backward_joints_matches_autograd(emu, oracle_fp32):
    rs = np.random.RandomState(5)
    inp = synthetic.make_fit_inputs(6, seed=11)
    dj = rs.randn(6, 49, 3).astype(np.float32)
    pose = torch.tensor(inp['pose'], dtype=torch.float64, requires_grad=True)
    betas = torch.tensor(inp['betas'], dtype=torch.float64, requires_grad=True)
    smpl64 = _oracle64().smpl
    out = smpl64(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    (out.joints * torch.tensor(dj, dtype=torch.float64)).sum().backward()
    d_pose, d_betas = emu.pose_backward(inp['pose'], inp['betas'], d_joints=dj)
    np.testing.assert_allclose(d_pose, pose.grad.numpy(), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(d_betas, betas.grad.numpy(), rtol=2e-4, atol=2e-5)
    # rotation-matrix mode
    from oracle import port
    R = port.exp_map_rodrigues(torch.tensor(inp['pose'], dtype=torch.float64).reshape(-1, 3)).view(6, 24, 3, 3)
    R = R.clone().requires_grad_(True)
    b2 = torch.tensor(inp['betas'], dtype=torch.float64, requires_grad=True)
    out = smpl64(global_orient=R[:, :1], body_pose=R[:, 1:], betas=b2, pose2rot=False)
    (out.joints * torch.tensor(dj, dtype=torch.float64)).sum().backward()
    d_R, d_b = emu.pose_backward(R.detach().numpy().reshape(6, 216).astype(np.float32), inp['betas'], d_joints=dj, rotmat_mode=True)
    np.testing.assert_allclose(d_R, R.grad.numpy().reshape(6, 216), rtol=2e-4, atol=2e-5)
    np.testing.assert_allclose(d_b, b2.grad.numpy(), rtol=2e-4, atol=2e-5)


_O64 = []


def _oracle64():
    if not _O64:
        from oracle import port
        _O64.append(port.build_oracle(seed=0, dtype=torch.float64))
    return _O64[0]


def test_backward_vertex_path_matches_autograd(emu):
    """dL/dA and dL/dx fed in from the vertex kernels reach pose and betas correctly."""
    rs = np.random.RandomState(6)
    inp = synthetic.make_fit_inputs(4, seed=12)
    smpl64 = _oracle64().smpl
    pose = torch.tensor(inp['pose'], dtype=torch.float64, requires_grad=True)
    betas = torch.tensor(inp['betas'], dtype=torch.float64, requires_grad=True)
    out = smpl64(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    gv = rs.randn(4, 6890, 3)
    (out.vertices * torch.tensor(gv)).sum().backward()
    # numpy restatement of what lbs_vertex_backward_kernel produces
    arrays = emu_mod.model_arrays(0)
    _, A, x = emu.pose_forward(inp['pose'], inp['betas'])
    basis = emu.array('basis').reshape(224, -1)[:, :20670].astype(np.float64)
    W = arrays['weights'].astype(np.float64)
    vp = (x.astype(np.float64) @ basis).reshape(4, 6890, 3)
    T = np.einsum('vj,bje->bve', W, A.reshape(4, 24, 12).astype(np.float64)).reshape(4, 6890, 3, 4)
    dvp = np.einsum('bvrc,bvr->bvc', T[..., :3], gv)
    dT = np.concatenate([gv[..., None] * vp[:, :, None, :], gv[..., None]], axis=-1)      # [B,V,3,4]
    dA = np.einsum('vj,bve->bje', W, dT.reshape(4, 6890, 12)).astype(np.float32)
    dx = (dvp.reshape(4, -1) @ basis.T).astype(np.float32)
    d_pose, d_betas = emu.pose_backward(inp['pose'], inp['betas'], dA=dA, dx=dx)
    np.testing.assert_allclose(d_pose, pose.grad.numpy(), rtol=3e-4, atol=3e-4)
    np.testing.assert_allclose(d_betas, betas.grad.numpy(), rtol=3e-4, atol=3e-4)


def test_fitting_loss_and_side_effect(emu):
    g = golden('smplify_default.npz')
    inp = {k: g[k] for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')}
    out = emu.fit(inp, loss_only=True)
    np.testing.assert_allclose(out['reproj'], g['init_fitting_loss'], rtol=2e-5, atol=1e-3)
    assert np.array_equal(out['keypoints'], g['keypoints_after_loss'])


@pytest.mark.parametrize('variant', ['default', 'trainer', 'slp'])
def test_fit_matches_reference_run(emu, variant):
    """100 + 100 iterations against the vectors the reference's own files produced."""
    g = golden('smplify_%s.npz' % variant)
    inp = {k: g[k] for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')}
    out = emu.fit(inp, num_iters=100)
    sums = out['trace'].astype(np.float64).sum(axis=1)
    np.testing.assert_allclose(sums, g['loss_trace'], rtol=1e-5)
    np.testing.assert_allclose(out['pose'], g['out_pose'], atol=1e-4)
    np.testing.assert_allclose(out['betas'], g['out_betas'], atol=1e-4)
    np.testing.assert_allclose(out['cam_t'], g['out_cam_t'], atol=1e-4)
    np.testing.assert_allclose(out['joints'], g['out_joints'], atol=1e-4)
    np.testing.assert_allclose(out['reproj'], g['out_reproj'], rtol=1e-4, atol=1e-2)
    assert np.all(out['keypoints'][:, C.SMPLIFY_IGNORED_JOINTS, 2] == 0)
    keep = [j for j in range(49) if j not in C.SMPLIFY_IGNORED_JOINTS]
    assert np.array_equal(out['keypoints'][:, keep], g['keypoints'][:, keep])


def test_fit_beyond_the_on_chip_adam_table(emu):
    """Steps >= 256 of a stage take their Adam bias corrections from adam_scalars() on the fly (the table in shared memory
    holds 256): a 270 + 270 iteration fit must track the oracle to the end, and its first 256 camera-stage losses must be
    those of a 256-iteration fit (same table)."""
    from oracle import port
    inp = synthetic.make_fit_inputs(2, seed=31)
    out = emu.fit(inp, num_iters=270)
    short = emu.fit(inp, num_iters=256)
    assert np.array_equal(out['trace'][:256], short['trace'][:256])
    oracle = port.build_oracle(seed=0, num_iters=270)
    trace = []
    ref = oracle(*[torch.from_numpy(inp[k].copy()) for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')], trace=trace)
    tr_o = torch.stack(trace).double().sum(1).numpy()
    np.testing.assert_allclose(out['trace'].astype(np.float64).sum(1), tr_o, rtol=1e-5)
    np.testing.assert_allclose(out['pose'], ref[2].numpy(), atol=1e-4)
    np.testing.assert_allclose(out['cam_t'], ref[4].numpy(), atol=1e-4)


def _check_prior_terms(out, g):
    """Prior terms of one evaluation against the reference-run vectors of tests/golden/prior_near_tie.npz."""
    assert np.array_equal(out['argmin'], g['argmin'])                    # same arg-min although the two best are ~1e-6 apart
    np.testing.assert_allclose(out['components'], g['components'], rtol=1e-6)
    np.testing.assert_allclose(out['terms'][:, 0], 4.78 ** 2 * g['nll'], rtol=2e-6)
    np.testing.assert_allclose(out['terms'][:, 1], 15.2 ** 2 * g['angle'], rtol=2e-6)
    np.testing.assert_allclose(out['terms'][:, 2], 25. * g['shape'], rtol=2e-6)
    scale = np.abs(g['grad_body_pose']).max()
    np.testing.assert_allclose(out['grad_body_pose'], g['grad_body_pose'], rtol=1e-5, atol=2e-6 * scale)
    np.testing.assert_allclose(out['grad_betas'], g['grad_betas'], rtol=1e-6)


def test_prior_terms_and_near_tie_argmin(emu):
    """Rows a15 / a16 (prior part) / a19 in isolation: max-mixture prior, angle prior, shape prior and their gradients at
    poses where two mixture components are within 3e-6 .. 1e-5 (relative) of each other."""
    g = golden('prior_near_tie.npz')
    assert g['rel_gap'].min() < 4e-6
    pose = np.concatenate([np.zeros((len(g['body_pose']), 3), np.float32), g['body_pose']], axis=1)
    _check_prior_terms(emu.prior_terms(pose, g['betas']), g)
    # the round-1 vectors (random poses, no tie): value and gradient of the GMM term alone
    g1 = golden('prior.npz')
    pose1 = np.concatenate([np.zeros((16, 3), np.float32), g1['body_pose']], axis=1)
    out = emu.prior_terms(pose1, np.zeros((16, 10), np.float32))
    np.testing.assert_allclose(out['terms'][:, 0], 4.78 ** 2 * g1['nll'], rtol=2e-6)


@pytest.mark.parametrize('variant', ['default', 'slp'])
def test_fit_on_sparse_structured_model(variant):
    """The folded constants (Cf, wkj, sigma_src, J0/JS) built from a SPARSE model - <= 10 vertices per regressor row, <= 4
    skinning weights per vertex, exact zeros, extra-regressor rows that do not sum to 1 - against the reference's own run."""
    g = golden('smplify_sparse_%s.npz' % variant)
    e = emu_mod.Emu(seed=int(g['model_seed']), structure='sparse')
    inp = {k: g[k] for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')}
    out = e.fit(inp, num_iters=100)
    np.testing.assert_allclose(out['trace'].astype(np.float64).sum(axis=1), g['loss_trace'], rtol=1e-5)
    np.testing.assert_allclose(out['pose'], g['out_pose'], atol=1e-4)
    np.testing.assert_allclose(out['betas'], g['out_betas'], atol=1e-4)
    np.testing.assert_allclose(out['cam_t'], g['out_cam_t'], atol=1e-4)
    np.testing.assert_allclose(out['joints'], g['out_joints'], atol=1e-4)
    fs = golden('smpl_forward_sparse.npz')
    joints, A, x = e.pose_forward(fs['pose'], fs['betas'])
    np.testing.assert_allclose(joints, fs['joints'], atol=2e-6)
    d_pose, d_betas = e.pose_backward(fs['pose'], fs['betas'], d_joints=fs['grad_joints'])
    assert np.all(np.isfinite(d_pose)) and np.all(np.isfinite(d_betas))
