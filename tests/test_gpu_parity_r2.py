"""GPU parity, pinned DIRECTLY to the oracle / reference-run vectors on the configurations the numbers are quoted on:

  * rows of the headline batch (B = 4096 + 7: 16-sample tiles of the first wave, 12-sample tiles of the second) against the
    oracle at 100 + 100 iterations - parameters / joints / vertices 1e-4 abs, per-sample loss of every iteration 1e-5 rel;
  * BASELINE config 3 (B = 256 SMPLify + rotation-matrix-mode SMPL forward / backward) against the oracle;
  * two shards of the 65 536-sample bulk refit (config 4) against the oracle;
  * the prior terms in isolation incl. near-tie arg-min (rows a15, a16, a19), both GEMM forms of the prior phase;
  * a SPARSE-structured model (the sparsity pattern of the real SMPL file) against the reference's own run.
Everything goes through the C ABI."""
import numpy as np
import pytest
import torch

from conftest import golden
from inbed_pose_estimation_b200 import constants as C, sharded, synthetic
from inbed_pose_estimation_b200.smpl import SMPL

pytestmark = pytest.mark.gpu
KEYS = ('pose', 'betas', 'cam_t', 'center', 'keypoints')


@pytest.fixture(scope='module')
def fitter():
    return synthetic.build_smplify('cuda', num_iters=100, seed=0)


def _cuda(inp):
    return [torch.from_numpy(np.ascontiguousarray(inp[k]).copy()).cuda() for k in KEYS]


def _oracle_rows(oracle, inp, rows):
    sub = {k: inp[k][rows] for k in KEYS}
    trace = []
    out = oracle(*[torch.from_numpy(sub[k].copy()) for k in KEYS], trace=trace)
    return out, torch.stack(trace).numpy()


_O64 = []


def _assert_loss_trace(got, ref32, inp, rows, what, max_adjudicated=4):
    """Per-sample loss of every iteration, 1e-5 relative, against the fp32 oracle.  The fp32 oracle is itself only an
    approximation of the arithmetic it stands for: on a few ill-conditioned samples ITS rounding noise exceeds 1e-5 late in
    the fit (measured in the build container at batch 256: oracle fp32 vs the same oracle in fp64 up to 3.0e-5 on one sample,
    this library vs fp64 at most 2e-6).  Entries beyond 1e-5 are therefore adjudicated by the oracle run in float64 on those
    rows: there this library must be within 1e-5 of the float64 trace or - where fp32 rounding noise itself is larger than
    that - at least closer to it than the fp32 oracle is (the reduction order of the GEMMs differs between the kernels: the
    small-batch cluster kernel measured 1.12e-5 on one sample of batch 256 in the last iterations, the fp32 oracle 3e-5 there),
    and wherever the two fp32 traces disagree by more than 1e-5 the fp32 oracle must be the farther one."""
    rel = np.abs(got - ref32) / np.abs(ref32)
    bad = np.unique(np.nonzero(rel > 1e-5)[1])
    if bad.size == 0:
        return
    assert bad.size <= max_adjudicated, '%s: %d samples beyond 1e-5 (max rel %.2e)' % (what, bad.size, rel.max())
    _, tr64 = _oracle64_rows(inp, rows[bad])
    ours = np.abs(got[:, bad] - tr64) / np.abs(tr64)
    theirs = np.abs(ref32[:, bad] - tr64) / np.abs(tr64)
    beyond = ours > 1e-5
    assert np.all(ours[beyond] < theirs[beyond]) and ours.max() <= 3e-5, '%s: %.2e from the float64 oracle' % (what, ours.max())
    off = rel[:, bad] > 1e-5
    assert np.all(theirs[off] > ours[off]), what + ': the deviation is not the fp32 oracle\'s own rounding'


def _oracle64_rows(inp, rows):
    from oracle import port
    if not _O64:
        _O64.append(port.build_oracle(seed=0, dtype=torch.float64))
    tr64 = []
    out = _O64[0](*[torch.from_numpy(inp[k][rows].copy()).double() for k in KEYS], trace=tr64)
    return [t.detach().numpy() for t in out], torch.stack(tr64).numpy()


def _assert_fit_rows(out, trace, rows, ref, ref_trace, what, inp=None, max_adjudicated=4):
    """out: the six result tensors of the big GPU fit; trace [200, B]; rows: the row indices the oracle fitted.

    Fitted parameters / joints / vertices: 1e-4 absolute against the fp32 oracle.  A fit is 200 chained Adam steps and a few
    samples are ill-conditioned enough that the fp32 ORACLE ITSELF ends more than 1e-4 away from the same oracle run in
    float64 (build container, batch 256, seed 33: rows 246 and 249, 1.5e-4 and 2.3e-4; this library 1.3e-5 and 1.2e-4 from
    float64 on the same rows).  Rows beyond 1e-4 are therefore adjudicated in float64: this library must be no farther from
    the float64 result than max(1e-4, three times the fp32 oracle's own distance to it) - different fp32 roundings of an
    ill-conditioned fit scatter by that much around the exact result."""
    got = [t[rows].cpu().numpy() for t in out]
    want = [t.detach().numpy() for t in ref]
    _assert_loss_trace(trace[:, rows], ref_trace, inp, rows, what + ': per-sample loss of every iteration')
    good = _adjudicated_close(got[:5], want[:5], ('vertices', 'joints', 'pose', 'betas', 'camera'), (0, 1, 2, 3, 4), inp, rows, what,
                              max_adjudicated)
    np.testing.assert_allclose(got[5][good], want[5][good], rtol=1e-4, atol=1e-2, err_msg=what + ': reprojection')


def _adjudicated_close(got, want, names, oracle_slots, inp, rows, what, max_adjudicated=4):
    """got / want: lists of [len(rows), ...] arrays (this library / the fp32 oracle); oracle_slots: which entries of the oracle's
    6-tuple they are.  1e-4 absolute; rows beyond it are re-fitted by the float64 oracle and must be no farther from it than
    max(1e-4, three times the fp32 oracle's own distance).  Returns the indices of the rows that met 1e-4 directly."""
    bad = set()
    for g, w in zip(got, want):
        err = np.abs(g - w).reshape(len(rows), -1).max(axis=1)
        bad.update(np.nonzero(err > 1e-4)[0].tolist())
    bad = sorted(bad)
    assert len(bad) <= max_adjudicated, '%s: %d rows beyond 1e-4 of the fp32 oracle' % (what, len(bad))
    good = np.setdiff1d(np.arange(len(rows)), bad)
    for g, w, nm in zip(got, want, names):
        np.testing.assert_allclose(g[good], w[good], atol=1e-4, err_msg='%s: %s' % (what, nm))
    if bad:
        ref64, _ = _oracle64_rows(inp, np.asarray(rows)[bad])
        for g, w, slot, nm in zip(got, want, oracle_slots, names):
            x = ref64[slot]
            ours = np.abs(g[bad] - x).reshape(len(bad), -1).max(axis=1)
            theirs = np.abs(w[bad] - x).reshape(len(bad), -1).max(axis=1)
            assert np.all(ours <= np.maximum(1e-4, 3.0 * theirs)), \
                '%s: %s rows %s: %s from the float64 oracle (the fp32 oracle: %s)' % (what, nm, np.asarray(rows)[bad], ours, theirs)
    return good


def test_headline_batch_rows_match_oracle(fitter, oracle_fp32):
    """B = 4103 = 148 x 16 + 144 x 12 + 7 on a 148-SM part: 32 rows out of the 16-sample wave and 39 out of the 12-sample wave
    (incl. the ragged last tile) are fitted by the oracle and compared directly - no detour through a small GPU batch."""
    B = 4096 + 7
    inp = synthetic.make_fit_inputs(B, seed=5)
    out = fitter(*_cuda(inp), return_loss_trace=True)
    trace = fitter.last_loss_trace.cpu().numpy()
    for rows, what in ((np.r_[0:8, 1000:1016, 2360:2368], '16-sample tiles'), (np.r_[2368:2384, 3500:3516, B - 7:B], '12-sample tiles')):
        ref, ref_trace = _oracle_rows(oracle_fp32, inp, rows)
        _assert_fit_rows(out, trace, rows, ref, ref_trace, what, inp)


def test_config3_batch256_matches_oracle(fitter, oracle_fp32):
    """BASELINE config 3: SMPLify at batch 256 (64 tiles of 4 samples) and the rotation-matrix-mode SMPL forward + backward
    of the cascade stage (train/trainer.py:597-615), all 256 rows against the oracle."""
    from oracle import port
    inp = synthetic.make_fit_inputs(256, seed=33)
    out = fitter(*_cuda(inp), return_loss_trace=True)
    trace = fitter.last_loss_trace.cpu().numpy()
    rows = np.arange(256)
    ref, ref_trace = _oracle_rows(oracle_fp32, inp, rows)
    _assert_fit_rows(out, trace, rows, ref, ref_trace, 'batch 256', inp)
    # rotation-matrix mode forward + backward at the same batch, fp64 oracle
    o64 = port.build_oracle(seed=0, dtype=torch.float64).smpl
    rs = np.random.RandomState(256)
    gv, gj = rs.randn(256, 6890, 3), rs.randn(256, 49, 3)
    R64 = port.exp_map_rodrigues(torch.tensor(inp['pose'], dtype=torch.float64).reshape(-1, 3)).view(256, 24, 3, 3).clone().requires_grad_(True)
    b64 = torch.tensor(inp['betas'], dtype=torch.float64, requires_grad=True)
    o = o64(global_orient=R64[:, :1], body_pose=R64[:, 1:], betas=b64, pose2rot=False)
    ((o.vertices * torch.tensor(gv)).sum() + (o.joints * torch.tensor(gj)).sum()).backward()
    R = R64.detach().float().cuda().requires_grad_(True)
    b = torch.from_numpy(inp['betas']).cuda().requires_grad_(True)
    g = fitter.smpl(global_orient=R[:, :1], body_pose=R[:, 1:], betas=b, pose2rot=False)
    ((g.vertices * torch.tensor(gv, dtype=torch.float32).cuda()).sum() + (g.joints * torch.tensor(gj, dtype=torch.float32).cuda()).sum()).backward()
    np.testing.assert_allclose(g.vertices.detach().cpu().numpy(), o.vertices.detach().numpy(), atol=1e-5)
    np.testing.assert_allclose(g.joints.detach().cpu().numpy(), o.joints.detach().numpy(), atol=1e-5)
    sR, sb = np.abs(R64.grad.numpy()).max(), np.abs(b64.grad.numpy()).max()
    np.testing.assert_allclose(R.grad.cpu().numpy(), R64.grad.numpy(), rtol=1e-4, atol=1e-5 * sR)
    np.testing.assert_allclose(b.grad.cpu().numpy(), b64.grad.numpy(), rtol=1e-4, atol=1e-5 * sb)


def test_bulk_refit_shards_match_oracle(fitter, oracle_fp32):
    """BASELINE config 4 at full size (65 536 samples, 100 + 100 iterations, keep-if-better): rows out of two different
    8192-sample shards (the first and the last of an 8-GPU split) against the oracle's fit of the same rows, and the
    keep-if-better decision / stored fits derived from the oracle's losses."""
    N = 65536
    inp = synthetic.make_fit_inputs(N, seed=4)
    dev = torch.device('cuda')
    fits = torch.from_numpy(np.concatenate([inp['pose'], inp['betas']], axis=1)).to(dev)
    cam, cen, kp = (torch.from_numpy(inp[k]).to(dev) for k in ('cam_t', 'center', 'keypoints'))
    loss0 = fitter.get_fitting_loss(fits[:, :72], fits[:, 72:], cam, cen, kp.clone()).mean(dim=-1)
    refit = sharded.ShardedRefit(smplify=fitter, device=dev)
    f1, l1, up1, cam1 = refit(fits, cam, cen, kp, loss0)
    for rows in (np.r_[100:116, 8000:8016], np.r_[N - 8192:N - 8176, N - 16:N]):
        sub = {k: inp[k][rows] for k in KEYS}
        ref = oracle_fp32(*[torch.from_numpy(sub[k].copy()) for k in KEYS])
        new_loss = ref[5].mean(dim=-1).numpy()
        old = loss0[rows].cpu().numpy()
        better = new_loss < old
        margin = np.abs(new_loss - old) > 1e-3 * np.abs(old)               # away from ties the decision is the oracle's
        assert np.array_equal(up1[rows].cpu().numpy()[margin], better[margin])
        sel = better & margin
        assert sel.sum() >= 16
        good = _adjudicated_close([f1[rows][:, :72].cpu().numpy()[sel], f1[rows][:, 72:].cpu().numpy()[sel], cam1[rows].cpu().numpy()[sel]],
                                  [ref[2].numpy()[sel], ref[3].numpy()[sel], ref[4].detach().numpy()[sel]], ('pose', 'betas', 'camera'),
                                  (2, 3, 4), inp, rows[sel], 'bulk refit shard')
        np.testing.assert_allclose(l1[rows].cpu().numpy()[sel][good], new_loss[sel][good], rtol=1e-4)


# ---- priors in isolation ---------------------------------------------------------------------------------------------
def _check_prior_terms(out, g, rows=slice(None)):
    out = {k: v.cpu().numpy()[rows] for k, v in out.items()}
    assert np.array_equal(out['argmin'], g['argmin'])                    # same arg-min although the two best are ~1e-6 apart
    np.testing.assert_allclose(out['components'], g['components'], rtol=1e-6)
    np.testing.assert_allclose(out['terms'][:, 0], 4.78 ** 2 * g['nll'], rtol=2e-6)
    np.testing.assert_allclose(out['terms'][:, 1], 15.2 ** 2 * g['angle'], rtol=2e-6)
    np.testing.assert_allclose(out['terms'][:, 2], 25. * g['shape'], rtol=2e-6)
    scale = np.abs(g['grad_body_pose']).max()
    np.testing.assert_allclose(out['grad_body_pose'], g['grad_body_pose'], rtol=1e-5, atol=2e-6 * scale)
    np.testing.assert_allclose(out['grad_betas'], g['grad_betas'], rtol=1e-6)


def test_prior_terms_near_tie_argmin(fitter):
    """smplify/prior.py:181-196 + losses.py:19-24,46-52 on poses where two mixture components are 3e-6 .. 1e-5 (relative)
    apart: same arg-min, values and gradients as the reference's own run.  Batch 32 runs the K-split GEMM form of the
    prior phase (8-sample tiles), batch 2048 + 32 the two-sample-group form of the 16-sample tiles."""
    g = golden('prior_near_tie.npz')
    n = len(g['body_pose'])
    pose = np.concatenate([np.zeros((n, 3), np.float32), g['body_pose']], axis=1)
    _check_prior_terms(fitter.prior_terms(torch.from_numpy(pose).cuda(), torch.from_numpy(g['betas']).cuda()), g)
    reps = 2048 // n + 1
    big_pose = torch.from_numpy(np.tile(pose, (reps, 1))).cuda()
    big_betas = torch.from_numpy(np.tile(g['betas'], (reps, 1))).cuda()
    out = fitter.prior_terms(big_pose, big_betas)
    for r in (0, reps // 2, reps - 1):
        _check_prior_terms(out, g, slice(r * n, (r + 1) * n))
    g1 = golden('prior.npz')
    pose1 = np.concatenate([np.zeros((16, 3), np.float32), g1['body_pose']], axis=1)
    o1 = fitter.prior_terms(torch.from_numpy(pose1).cuda(), torch.zeros(16, 10).cuda())
    np.testing.assert_allclose(o1['terms'][:, 0].cpu().numpy(), 4.78 ** 2 * g1['nll'], rtol=2e-6)
    np.testing.assert_allclose(o1['grad_body_pose'].cpu().numpy() / 4.78 ** 2 -
                               _angle_grad(g1['body_pose']) / 4.78 ** 2, g1['grad'], rtol=1e-5, atol=2e-6 * np.abs(g1['grad']).max())


def _angle_grad(body_pose):
    """d/dpose of 15.2^2 * sum exp(sign * pose[id])^2 (losses.py:19-24), to take the angle prior out of the total gradient."""
    gr = np.zeros_like(body_pose)
    for i, s in zip(C.ANGLE_PRIOR_IDS, C.ANGLE_PRIOR_SIGNS):
        gr[:, i] += 15.2 ** 2 * 2.0 * s * np.exp(s * body_pose[:, i]) ** 2
    return gr


# ---- sparse-structured model -------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def sparse_fitter():
    g = golden('smplify_sparse_default.npz')
    return synthetic.build_smplify('cuda', num_iters=100, seed=int(g['model_seed']), structure='sparse')


@pytest.mark.parametrize('variant', ['default', 'slp'])
def test_sparse_model_fit_matches_reference_golden(sparse_fitter, variant):
    """J_regressor with <= 10 vertices per joint, <= 4 skinning weights per vertex, exact zeros, extra-regressor rows that do
    not sum to 1: the folded constants, the hi/lo operand splits and the padded tiles against the reference's own run."""
    g = golden('smplify_sparse_%s.npz' % variant)
    args = _cuda({k: g[k] for k in KEYS})
    v, j, pose, betas, cam, reproj = sparse_fitter(*args, return_loss_trace=True)
    trace = sparse_fitter.last_loss_trace.double().sum(dim=1).cpu().numpy()
    np.testing.assert_allclose(trace, g['loss_trace'], rtol=1e-5)
    np.testing.assert_allclose(pose.cpu().numpy(), g['out_pose'], atol=1e-4)
    np.testing.assert_allclose(betas.cpu().numpy(), g['out_betas'], atol=1e-4)
    np.testing.assert_allclose(cam.cpu().numpy(), g['out_cam_t'], atol=1e-4)
    np.testing.assert_allclose(j.cpu().numpy(), g['out_joints'], atol=1e-4)
    np.testing.assert_allclose(v.cpu().numpy()[:, ::8], g['out_vertices_sub'], atol=1e-4)
    np.testing.assert_allclose(reproj.cpu().numpy(), g['out_reproj'], rtol=1e-4, atol=1e-2)


def test_sparse_model_smpl_forward_backward(sparse_fitter):
    g = golden('smpl_forward_sparse.npz')
    smpl = sparse_fitter.smpl
    gv, gj = torch.from_numpy(g['grad_vertices']).cuda(), torch.from_numpy(g['grad_joints']).cuda()
    pose = torch.from_numpy(g['pose']).cuda().requires_grad_(True)
    betas = torch.from_numpy(g['betas']).cuda().requires_grad_(True)
    out = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    np.testing.assert_allclose(out.vertices.detach().cpu().numpy()[:, ::8], g['vertices_sub'], atol=5e-6)
    np.testing.assert_allclose(out.joints.detach().cpu().numpy(), g['joints'], atol=5e-6)
    np.testing.assert_allclose(out.vertices.detach().double().sum(dim=1).cpu().numpy(), g['vertices_checksum'], atol=2e-3)
    ((out.vertices * gv).sum() + (out.joints * gj).sum()).backward()
    np.testing.assert_allclose(pose.grad.cpu().numpy(), g['grad_pose'], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(betas.grad.cpu().numpy(), g['grad_betas'], rtol=1e-4, atol=2e-4)
    R = torch.from_numpy(g['rotmats']).cuda().requires_grad_(True)
    b2 = torch.from_numpy(g['betas']).cuda().requires_grad_(True)
    out2 = smpl(global_orient=R[:, :1], body_pose=R[:, 1:], betas=b2, pose2rot=False)
    np.testing.assert_allclose(out2.vertices.detach().cpu().numpy()[:, ::8], g['vertices_rotmat_sub'], atol=5e-6)
    ((out2.vertices * gv).sum() + (out2.joints * gj).sum()).backward()
    np.testing.assert_allclose(R.grad.cpu().numpy(), g['grad_rotmats'], rtol=1e-4, atol=2e-4)
    np.testing.assert_allclose(b2.grad.cpu().numpy(), g['grad_betas_rotmat'], rtol=1e-4, atol=2e-4)


def test_sparse_model_against_independent_fp64_skinning(sparse_fitter):
    """An independent float64 formulation (plain numpy: blend shapes, per-vertex transforms as an explicit weighted sum over
    the <= 4 non-zero joints, no folding, no hi/lo split) of the vertices and the 49 joints on the sparse model."""
    seed = int(golden('smplify_sparse_default.npz')['model_seed'])
    m = synthetic.make_smpl_model(seed, 'sparse')
    jx = synthetic.make_extra_regressor(seed + 1, 'sparse').astype(np.float64)
    inp = synthetic.make_fit_inputs(5, seed=77)
    pose, betas = inp['pose'].astype(np.float64), inp['betas'].astype(np.float64)
    B = 5
    v_shaped = m['v_template'][None] + np.einsum('vcl,bl->bvc', m['shapedirs'], betas)
    J = np.einsum('jv,bvc->bjc', m['J_regressor'], v_shaped)
    r = pose.reshape(B, 24, 3)
    ang = np.linalg.norm(r + 1e-8, axis=-1, keepdims=True)
    n = r / ang
    K = np.zeros((B, 24, 3, 3))
    K[..., 0, 1], K[..., 0, 2], K[..., 1, 0] = -n[..., 2], n[..., 1], n[..., 2]
    K[..., 1, 2], K[..., 2, 0], K[..., 2, 1] = -n[..., 0], -n[..., 1], n[..., 0]
    R = np.eye(3) + np.sin(ang)[..., None] * K + (1 - np.cos(ang))[..., None] * (K @ K)
    feat = (R[:, 1:] - np.eye(3)).reshape(B, 207)
    v_posed = v_shaped + np.einsum('vcp,bp->bvc', m['posedirs'], feat)
    G = np.zeros((B, 24, 4, 4))
    for j in range(24):
        L = np.zeros((B, 4, 4))
        L[:, :3, :3] = R[:, j]
        L[:, 3, 3] = 1
        p = C.SMPL_PARENTS[j]
        L[:, :3, 3] = J[:, j] - (J[:, p] if j else 0)
        G[:, j] = L if j == 0 else G[:, p] @ L
    A = G.copy()
    A[:, :, :3, 3] -= np.einsum('bjrc,bjc->bjr', G[:, :, :3, :3], J)
    verts = np.zeros((B, 6890, 3))
    W = m['weights']
    for v in range(6890):
        T = np.zeros((B, 4, 4))
        for j in np.nonzero(W[v])[0]:
            T += W[v, j] * A[:, j]
        verts[:, v] = np.einsum('brc,bc->br', T[:, :3, :3], v_posed[:, v]) + T[:, :3, 3]
    joints54 = np.concatenate([G[:, :, :3, 3], verts[:, C.SMPL_EXTRA_VERTEX_IDS], np.einsum('kv,bvc->bkc', jx, verts)], axis=1)
    joints = joints54[:, [C.JOINT_MAP[nm] for nm in C.JOINT_NAMES]]
    p = torch.from_numpy(inp['pose']).cuda()
    out = sparse_fitter.smpl(global_orient=p[:, :3], body_pose=p[:, 3:], betas=torch.from_numpy(inp['betas']).cuda())
    np.testing.assert_allclose(out.vertices.cpu().numpy(), verts, atol=1e-5)
    np.testing.assert_allclose(out.joints.cpu().numpy(), joints, atol=1e-5)
