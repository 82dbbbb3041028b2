"""GPU parity of the fused SMPLify fit (through the C ABI) against the oracle and the
reference-generated golden vectors.  Tolerances are BASELINE.json's: fitted parameters,
joints and vertices within 1e-4 absolute, per-iteration losses within 1e-5 relative (fp32)."""
import os

import numpy as np
import pytest
import torch

from conftest import golden
from inbed_pose_estimation_b200 import _native, constants as C, synthetic

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def fitter():
    return synthetic.build_smplify('cuda', num_iters=100, seed=0)


def _cuda(inp):
    return [torch.from_numpy(inp[k].copy()).cuda() for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')]


@pytest.mark.parametrize('variant', ['default', 'trainer', 'slp'])
def test_fit_matches_reference_golden(fitter, variant):
    g = golden('smplify_%s.npz' % variant)
    args = _cuda({k: g[k] for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')})
    v, j, pose, betas, cam, reproj = fitter(*args, return_loss_trace=True)
    trace = fitter.last_loss_trace.double().sum(dim=1).cpu().numpy()
    np.testing.assert_allclose(trace, g['loss_trace'], rtol=1e-5)
    np.testing.assert_allclose(pose.cpu().numpy(), g['out_pose'], atol=1e-4)
    np.testing.assert_allclose(betas.cpu().numpy(), g['out_betas'], atol=1e-4)
    np.testing.assert_allclose(cam.cpu().numpy(), g['out_cam_t'], atol=1e-4)
    np.testing.assert_allclose(j.cpu().numpy(), g['out_joints'], atol=1e-4)
    np.testing.assert_allclose(v.cpu().numpy()[:, ::8], g['out_vertices_sub'], atol=1e-4)
    np.testing.assert_allclose(reproj.cpu().numpy(), g['out_reproj'], rtol=1e-4, atol=1e-2)
    kp = args[4].cpu().numpy()
    assert np.all(kp[:, C.SMPLIFY_IGNORED_JOINTS, 2] == 0)           # in-place side effect (smplify.py:105)
    keep = [k for k in range(49) if k not in C.SMPLIFY_IGNORED_JOINTS]
    assert np.array_equal(kp[:, keep], g['keypoints'][:, keep])
    for t in (v, j, pose, betas, cam, reproj):
        assert not t.requires_grad


def test_fit_matches_oracle_batch32(fitter, oracle_fp32):
    """BASELINE config 2 (B=32, 100+100 iterations), per-sample losses of every iteration."""
    inp = synthetic.make_fit_inputs(32, seed=21)
    trace_o = []
    torch.set_num_threads(max(1, torch.get_num_threads()))
    vo, jo, po, bo, co, ro = oracle_fp32(*[torch.from_numpy(inp[k].copy()) for k in
                                           ('pose', 'betas', 'cam_t', 'center', 'keypoints')], trace=trace_o)
    v, j, pose, betas, cam, reproj = fitter(*_cuda(inp), return_loss_trace=True)
    tr = fitter.last_loss_trace.cpu().numpy()
    tr_o = torch.stack(trace_o).numpy()
    np.testing.assert_allclose(tr.astype(np.float64).sum(1), tr_o.astype(np.float64).sum(1), rtol=1e-5)
    # per sample, every iteration (north_star: 1e-5 relative; entries beyond it are adjudicated by the float64 oracle)
    from test_gpu_parity_r2 import _assert_loss_trace
    _assert_loss_trace(tr, tr_o, inp, np.arange(32), 'batch 32')
    np.testing.assert_allclose(pose.cpu().numpy(), po.numpy(), atol=1e-4)
    np.testing.assert_allclose(betas.cpu().numpy(), bo.numpy(), atol=1e-4)
    np.testing.assert_allclose(cam.cpu().numpy(), co.detach().numpy(), atol=1e-4)
    np.testing.assert_allclose(j.cpu().numpy(), jo.numpy(), atol=1e-4)
    np.testing.assert_allclose(v.cpu().numpy(), vo.numpy(), atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize('cluster', [8, 4, 2])
def test_small_batch_cluster_kernel_matches_oracle(fitter, oracle_fp32, cluster):
    """Small batches (the reference's --batch_size 32, README.md:33-35) run on clusters of 8 / 4 / 2 CTAs per 4-sample tile
    (csrc/fit_split.cuh): for every cluster size the LARGEST batch this device plans it for (all clusters resident at once),
    made ragged, against the fp32 oracle - per-sample loss of every iteration, parameters, joints, vertices."""
    planned = [b for b in range(1, 600) if _native.fit_split_plan(b) == cluster]
    if not planned:
        pytest.skip('this device never plans clusters of %d CTAs' % cluster)
    B = max(planned) - 2
    assert _native.fit_split_plan(B) == cluster
    inp = synthetic.make_fit_inputs(B, seed=300 + B)
    rows = np.unique(np.concatenate([np.arange(4), np.arange(B - 6, B), np.random.RandomState(B).choice(B, 6, replace=False)]))
    from test_gpu_parity_r2 import _assert_fit_rows, _oracle_rows
    ref, ref_trace = _oracle_rows(oracle_fp32, inp, rows)
    out = fitter(*_cuda(inp), return_loss_trace=True)
    trace = fitter.last_loss_trace.clone()
    _assert_fit_rows(out, trace.cpu().numpy(), rows, ref, ref_trace, 'cluster kernel, batch %d' % B, inp=inp)
    # every CTA of a cluster repeats the per-sample phases and exchanges GEMM rows through distributed shared memory: a race
    # there would show as run-to-run differences (compute-sanitizer is not available on the pool) - three more runs, bit for bit
    for _ in range(3):
        again = fitter(*_cuda(inp), return_loss_trace=True)
        assert torch.equal(fitter.last_loss_trace, trace)
        for a, b in zip(again, out):
            assert torch.equal(a, b)


@pytest.mark.gpu
def test_small_batch_cluster_kernel_agrees_with_the_tile_kernel():
    """The same batch through the cluster kernel and through the 4-sample tile kernel (SMPLB200_FIT_VARIANT is read once per
    process, hence tools/pair_debug.py's two child processes).  The two sum the GEMMs' reduction dimension in different
    slices, so they agree to fp32 rounding carried through 200 Adam steps, not bit for bit."""
    import subprocess, sys, re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'pair_debug.py'), '--batch', '42', '--iters', '100', '--variant', '12'],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True, timeout=600)
    assert r.returncode == 0 and 'FAILED' not in r.stdout, r.stdout[-2000:]
    diffs = dict(re.findall(r'^\s+(pose|betas|cam|joints|verts)\s+max abs diff ([0-9.e+-]+)', r.stdout, flags=re.M))
    assert set(diffs) == {'pose', 'betas', 'cam', 'joints', 'verts'}, r.stdout[-2000:]
    for k, v in diffs.items():
        assert float(v) <= 2e-4, (k, v)
    rel = [float(x) for x in re.findall(r'max rel loss diff ([0-9.e+-]+)', r.stdout)]
    assert rel and max(rel) <= 3e-5 and 'NON-FINITE' not in r.stdout, r.stdout[-2000:]


def test_fitting_loss_matches_golden(fitter):
    g = golden('smplify_default.npz')
    args = _cuda({k: g[k] for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')})
    loss = fitter.get_fitting_loss(*args)
    np.testing.assert_allclose(loss.cpu().numpy(), g['init_fitting_loss'], rtol=2e-5, atol=1e-3)
    assert np.array_equal(args[4].cpu().numpy(), g['keypoints_after_loss'])


def test_non_contiguous_keypoints_side_effect(fitter):
    """The confidence zeroing must reach the caller's tensor even through a strided view."""
    inp = synthetic.make_fit_inputs(5, seed=3)
    big = torch.zeros(5, 49, 4, device='cuda')
    big[:, :, :3] = torch.from_numpy(inp['keypoints']).cuda()
    view = big[:, :, :3]
    fitter.get_fitting_loss(*(_cuda(inp)[:4] + [view]))
    assert torch.all(view[:, C.SMPLIFY_IGNORED_JOINTS, 2] == 0)
    assert torch.all(view[:, 0, 2] == 1)


def _rows_agree(a, b, atol=1e-4, frac=0.94, worst=2e-3):
    """Rows of two fp32 fits of the same samples: at least `frac` of the rows within atol, none beyond `worst`."""
    err = np.abs(a - b).reshape(a.shape[0], -1).max(axis=1)
    assert (err <= atol).mean() >= frac and err.max() <= worst, (np.sort(err)[-5:], float((err <= atol).mean()))


def test_batch_4096_properties(fitter):
    """Full-size batch: every sample's fit is independent of its neighbours (sharding premise),
    ragged tail tiles work, the loss decreases, repeated runs are bit-identical."""
    B = 4096 + 7
    inp = synthetic.make_fit_inputs(B, seed=5)
    out = fitter(*_cuda(inp), return_loss_trace=True)
    trace = fitter.last_loss_trace.cpu().numpy()
    assert np.all(np.isfinite(trace)) and all(torch.isfinite(t).all() for t in out)
    assert trace[199].sum() < trace[100].sum() and trace[99].sum() < trace[0].sum()
    out2 = fitter(*_cuda(inp))
    for a, b in zip(out, out2):
        assert torch.equal(a, b)                                      # deterministic
    # rows 1000:1050 sit in 16-sample tiles of the first wave, rows 4000: in the 12-sample tiles that fill the second
    # wave (4103 = 148 x 16 + 144 x 12 + 7: the last tile is ragged); alone they run in 4-sample tiles
    # The big batch runs on the pair kernel (tensor-core GEMMs), the same rows alone on the 4-sample tiles of the CUDA-core
    # kernel: two roundings of the same arithmetic.  After 200 chained Adam steps they agree to 1e-4 except on the few
    # ill-conditioned samples on which even the fp32 oracle is farther than that from its own float64 run (those rows are
    # adjudicated against the oracle in test_gpu_parity_r2.py::test_headline_batch_rows_match_oracle).
    for lo, hi in ((1000, 1050), (4000, B)):
        sub = {k: v[lo:hi] for k, v in inp.items()}
        out_sub = fitter(*_cuda(sub))
        for a, b in zip(out[:5], out_sub[:5]):
            _rows_agree(a[lo:hi].cpu().numpy(), b.cpu().numpy())
        np.testing.assert_allclose(out[5][lo:hi].cpu().numpy(), out_sub[5].cpu().numpy(), rtol=2e-3, atol=1e-2)
    ro = fitter.get_fitting_loss(out[2], out[3], out[4], torch.from_numpy(inp['center']).cuda(),
                                 torch.from_numpy(inp['keypoints'].copy()).cuda())
    # get_fitting_loss runs its forward on the CUDA cores, the fit's final forward on the tensor cores (3xTF32): joints agree to
    # ~1e-6 m, which is up to 1e-4 of a per-joint loss whose residual is a few pixels
    np.testing.assert_allclose(ro.cpu().numpy(), out[5].cpu().numpy(), rtol=2e-4, atol=5e-3)


def test_empty_batch_and_errors(fitter):
    z = lambda *s: torch.zeros(*s, device='cuda')
    out = fitter(z(0, 72), z(0, 10), z(0, 3), z(0, 2), z(0, 49, 3))
    assert out[0].shape == (0, 6890, 3) and out[5].shape == (0, 49)
    with pytest.raises(ValueError):
        fitter(z(2, 72), z(2, 10), z(2, 3), z(2, 2), z(2, 48, 3))
    with pytest.raises(RuntimeError, match='CUDA'):
        fitter(torch.zeros(2, 72), z(2, 10), z(2, 3), z(2, 2), z(2, 49, 3))


@pytest.mark.parametrize('B,small', [(16 * 148 + 30, 4), (16 * 148 + 700, 8), (16 * 148 + 1300, 12), (12 * 148 - 5, 12), (8 * 148, 8)])
def test_every_tile_mix_gives_the_same_fits(B, small):
    """The tile plan (16-sample waves + one wave of 4-, 8- or 12-sample tiles) only changes where a sample is computed:
    rows of the remainder wave and of the first wave equal the same rows fitted alone (4-sample tiles)."""
    from inbed_pose_estimation_b200 import _native
    n16, s, n_small = _native.fit_tile_plan(B, torch.cuda.get_device_properties(0).multi_processor_count)
    if torch.cuda.get_device_properties(0).multi_processor_count == 148:
        assert s == small
    fitter5 = synthetic.build_smplify('cuda', num_iters=5, seed=0)
    inp = synthetic.make_fit_inputs(B, seed=B)
    out = fitter5(*_cuda(inp))
    assert all(torch.isfinite(t).all() for t in out)
    for lo, hi in ((3, 29), (B - 37, B)):
        sub = {k: v[lo:hi] for k, v in inp.items()}
        out_sub = fitter5(*_cuda(sub))
        for a, b in zip(out[:5], out_sub[:5]):
            np.testing.assert_allclose(a[lo:hi].cpu().numpy(), b.cpu().numpy(), rtol=1e-4, atol=1e-4)
        # per-joint reprojection losses: a residual near zero turns a 1e-6 relative joint difference into 1e-4 of its loss
        np.testing.assert_allclose(out[5][lo:hi].cpu().numpy(), out_sub[5].cpu().numpy(), rtol=1e-4, atol=1e-2)


def test_more_iterations_than_the_on_chip_adam_table():
    """num_iters is unbounded in the reference; here the Adam scalars of steps >= 256 are computed on the fly.  A 300 + 300
    iteration fit must continue the 256-iteration trajectory (same first 256 losses per stage) and match the oracle at the end."""
    from oracle import port
    fit300 = synthetic.build_smplify('cuda', num_iters=300, seed=0)
    fit256 = synthetic.build_smplify('cuda', num_iters=256, seed=0)
    inp = synthetic.make_fit_inputs(6, seed=77)
    out = fit300(*_cuda(inp), return_loss_trace=True)
    tr300 = fit300.last_loss_trace.cpu().numpy()
    fit256(*_cuda(inp), return_loss_trace=True)
    tr256 = fit256.last_loss_trace.cpu().numpy()
    assert np.array_equal(tr300[:256], tr256[:256])                        # stage 1: identical prefix
    assert np.all(np.isfinite(tr300))
    oracle = port.build_oracle(seed=0, num_iters=300)
    trace = []
    ref = oracle(*[torch.from_numpy(inp[k].copy()) for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')], trace=trace)
    tr_o = torch.stack(trace).double().sum(1).numpy()
    np.testing.assert_allclose(tr300.astype(np.float64).sum(1), tr_o, rtol=1e-5)
    np.testing.assert_allclose(out[2].cpu().numpy(), ref[2].numpy(), atol=1e-4)


def test_host_buffer_entry_point_equals_device_path(fitter):
    """smplb200_smplify_fit_host (what bench.py times as e2e: H2D, fit, vertex kernels, D2H on two streams) returns the
    same bits as the device-pointer path, zeroes the ignored confidences in the caller's host buffer and fills vertices."""
    import ctypes
    from inbed_pose_estimation_b200 import _native
    B = 70
    inp = synthetic.make_fit_inputs(B, seed=123)
    dev = fitter(*_cuda(inp))
    lib = _native.lib()
    h = {k: np.ascontiguousarray(inp[k], dtype=np.float32).copy() for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')}
    out = {'verts': np.empty((B, 6890, 3), np.float32), 'joints': np.empty((B, 49, 3), np.float32), 'pose': np.empty((B, 72), np.float32),
           'betas': np.empty((B, 10), np.float32), 'cam': np.empty((B, 3), np.float32), 'reproj': np.empty((B, 49), np.float32)}
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    handle = fitter.smpl.native(torch.device('cuda', torch.cuda.current_device())).handle
    for with_verts in (True, False):
        kp = h['keypoints'].copy()
        _native.check(lib.smplb200_smplify_fit_host(handle, B, 100, 1e-2, 5000., p(h['pose']), p(h['betas']), p(h['cam_t']), p(h['center']),
                                                    p(kp), p(out['verts']) if with_verts else None, p(out['joints']), p(out['pose']),
                                                    p(out['betas']), p(out['cam']), p(out['reproj'])))
        for got, ref in ((out['joints'], dev[1]), (out['pose'], dev[2]), (out['betas'], dev[3]), (out['cam'], dev[4]), (out['reproj'], dev[5])):
            assert np.array_equal(got, ref.cpu().numpy())
        assert np.all(kp[:, C.SMPLIFY_IGNORED_JOINTS, 2] == 0)
    assert np.array_equal(out['verts'], dev[0].cpu().numpy())


def test_packed_result_rows_written_by_the_kernel(fitter):
    """packed_out [B,134] = pose | betas | camera | reprojection, the row a sharded refit all-gathers, bit for bit."""
    inp = synthetic.make_fit_inputs(37, seed=9)
    packed = torch.full((37, 134), float('nan'), device='cuda')
    v, j, pose, betas, cam, reproj = fitter(*_cuda(inp), packed_out=packed)
    assert torch.equal(packed, torch.cat([pose, betas, cam, reproj], dim=1))
    with pytest.raises(ValueError):
        fitter(*_cuda(inp), packed_out=torch.empty(37, 133, device='cuda'))
