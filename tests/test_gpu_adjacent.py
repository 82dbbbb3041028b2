"""GPU parity of the steps either side of SMPLify - through the C ABI - with the vectors the reference's
own files produced (tests/golden/adjacent.npz) and with the CPU restatement on larger seeded inputs."""
import numpy as np
import pytest
import torch

from conftest import golden
from inbed_pose_estimation_b200 import constants, fits_dict, geometry, sharded
from oracle import adjacent

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def g():
    return golden('adjacent.npz')


def test_rot6d(g):
    out = geometry.rot6d_to_rotmat(torch.from_numpy(g['rot6d_in']).cuda())
    np.testing.assert_allclose(out.cpu().numpy(), g['rot6d_out'], atol=2e-6)      # fp32 Gram-Schmidt: FMA contraction differs from eager torch
    x = torch.randn(5000, 6, generator=torch.Generator().manual_seed(1))
    np.testing.assert_allclose(geometry.rot6d_to_rotmat(x.cuda()).cpu().numpy(), adjacent.rot6d_to_rotmat(x).numpy(), atol=5e-6)
    assert geometry.rot6d_to_rotmat(torch.zeros(0, 6).cuda()).shape == (0, 3, 3)
    assert torch.equal(geometry.rotmat_to_rot6d(out), out[:, :2, :].reshape(-1, 6))


def test_rotmat_to_axis_angle(g):
    aa = geometry.rotation_matrix_to_angle_axis(torch.from_numpy(g['rotmat_in']).cuda(), scrub_nan=True)
    np.testing.assert_allclose(aa.cpu().numpy(), g['axis_angle_out'], atol=2e-6)
    # the trainer's shape: [B*24, 3, 4] with the homogeneous column appended
    R = torch.from_numpy(g['rotmat_in']).cuda()
    hom = torch.cat([R, torch.tensor([0., 0., 1.]).view(1, 3, 1).expand(R.shape[0], -1, -1).cuda()], dim=-1)
    assert torch.equal(geometry.rotation_matrix_to_angle_axis(hom, scrub_nan=True), aa)
    # round trip with batch_rodrigues on a large batch, and the scrub on a matrix that makes torchgeometry emit NaN
    th = 0.7 * torch.randn(20000, 3, generator=torch.Generator().manual_seed(2))
    th = th[th.norm(dim=1) < 3.0]                                 # beyond pi the axis-angle wraps to the equivalent rotation
    back = geometry.rotation_matrix_to_angle_axis(geometry.batch_rodrigues(th.cuda()))
    np.testing.assert_allclose(back.cpu().numpy(), th.numpy(), atol=5e-5)
    bad = torch.diag(torch.tensor([-1., -1., -1.])).view(1, 3, 3).cuda()          # not a rotation: sqrt of a negative number
    ref = adjacent.rotmat_to_axis_angle(bad.cpu(), scrub_nan=False)
    out = geometry.rotation_matrix_to_angle_axis(bad, scrub_nan=False).cpu()
    assert torch.equal(torch.isnan(ref), torch.isnan(out))
    if torch.isnan(ref).any():
        scrubbed = geometry.rotation_matrix_to_angle_axis(bad, scrub_nan=True).cpu()
        assert not torch.isnan(scrubbed).any() and torch.all(scrubbed[torch.isnan(ref)] == 0)


def test_estimate_translation(g):
    t = geometry.estimate_translation(torch.from_numpy(g['et_S']).cuda(), torch.from_numpy(g['et_kp']).cuda())
    np.testing.assert_allclose(t.cpu().numpy(), g['et_out'], rtol=2e-6, atol=2e-6)
    gen = torch.Generator().manual_seed(3)
    S = 0.4 * torch.randn(3000, 49, 3, generator=gen)
    kp = torch.cat([224 * torch.rand(3000, 49, 2, generator=gen), torch.rand(3000, 49, 1, generator=gen)], dim=-1)
    np.testing.assert_allclose(geometry.estimate_translation(S.cuda(), kp.cuda()).cpu().numpy(),
                               adjacent.estimate_translation(S, kp).numpy(), rtol=1e-5, atol=1e-5)


def test_fits_get_set_against_reference_vectors(g):
    store = torch.from_numpy(g['fits_store'].copy()).cuda()
    idx, rot, fl = torch.from_numpy(g['fits_index']), torch.from_numpy(g['fits_rot']), torch.from_numpy(g['fits_flipped'])
    pose, betas = fits_dict.fits_get(store, idx, rot, fl)
    # global orientation goes through fp32 sin/cos and an inverse Rodrigues whose conditioning degrades as 1/sin(angle):
    # 2e-6 for ordinary rows, 5e-5 for the stored orientation placed next to pi
    err = np.abs(pose.cpu().numpy() - g['fits_get_pose'])
    assert err.max() < 5e-5 and (err > 2e-6).sum() <= 3
    assert np.array_equal(betas.cpu().numpy(), g['fits_get_betas'])
    # body pose entries are only permuted / negated: bit exact
    assert np.array_equal(pose.cpu().numpy()[:, 3:], g['fits_get_pose'][:, 3:])
    fits_dict.fits_set(store, idx, rot, fl, torch.from_numpy(g['fits_update']), torch.from_numpy(g['fits_new_pose']).cuda(),
                       torch.from_numpy(g['fits_new_betas']).cuda())
    after = store.cpu().numpy()
    np.testing.assert_allclose(after, g['fits_store_after'], atol=5e-6)
    untouched = np.ones(after.shape[0], bool)
    untouched[g['fits_index'][g['fits_update'].astype(bool)]] = False
    assert np.array_equal(after[untouched], g['fits_store'][untouched])


def test_fits_dict_class_round_trip(tmp_path):
    gen = torch.Generator().manual_seed(5)
    arr = torch.cat([0.4 * torch.randn(500, 72, generator=gen), 0.5 * torch.randn(500, 10, generator=gen)], dim=1).numpy()
    fd = fits_dict.FitsDict(type('O', (), {'checkpoint_dir': str(tmp_path)})(), None, fits={'slp': arr, 'h36m': arr[::-1].copy()})
    names = ['slp'] * 100 + ['h36m'] * 60
    idx = torch.cat([torch.randperm(500, generator=gen)[:100], torch.randperm(500, generator=gen)[:60]])
    rot = 30 * torch.randn(160, generator=gen)
    fl = (torch.rand(160, generator=gen) < 0.5)
    pose, betas = fd[(names, idx, rot, fl)]
    ref_pose = torch.empty(160, 72)
    for lo, hi, a in ((0, 100, arr), (100, 160, arr[::-1].copy())):
        p, _ = adjacent.fits_get(torch.from_numpy(a), idx[lo:hi], rot[lo:hi], fl[lo:hi].to(torch.uint8))
        ref_pose[lo:hi] = p
    np.testing.assert_allclose(pose.cpu().numpy(), ref_pose.numpy(), atol=1e-5)
    # writing back what was read (get then set with update everywhere) reproduces the store up to the rotation round trip
    fd[(names, idx, rot, fl, torch.ones(160, dtype=torch.uint8))] = (pose, betas)
    np.testing.assert_allclose(fd.fits_dict['slp'].cpu().numpy(), arr, atol=2e-5)
    fd.save()
    assert np.load(str(tmp_path / 'slp_fits.npy')).shape == (500, 82)
    assert fd.flipped_parts.tolist() == constants.SMPL_POSE_FLIP_PERM


def test_fits_index_bounds_checked_on_the_device(tmp_path):
    """Out-of-range indices never touch memory outside the store: the get returns NaN rows, the set skips them, and the
    sticky device flag turns into an IndexError at the next check (FitsDict.check / save, or at once without a flag)."""
    gen = torch.Generator().manual_seed(6)
    arr = torch.randn(50, 82, generator=gen).numpy()
    store = torch.from_numpy(arr.copy()).cuda()
    idx = torch.tensor([3, 50, -1, 49])
    with pytest.raises(IndexError):
        fits_dict.fits_get(store, idx, torch.zeros(4), torch.zeros(4, dtype=torch.uint8))
    flag = torch.zeros(1, dtype=torch.int32, device='cuda')
    pose, betas = fits_dict.fits_get(store, idx, torch.zeros(4), torch.zeros(4, dtype=torch.uint8), status=flag)
    assert int(flag.item()) == 1
    assert torch.isnan(pose[1:3]).all() and torch.isnan(betas[1:3]).all()
    np.testing.assert_allclose(betas[[0, 3]].cpu().numpy(), arr[[3, 49], 72:], rtol=0, atol=0)
    flag.zero_()
    fits_dict.fits_set(store, idx, torch.zeros(4), torch.zeros(4, dtype=torch.uint8), torch.ones(4, dtype=torch.uint8),
                       torch.zeros(4, 72).cuda(), torch.ones(4, 10).cuda(), status=flag)
    assert int(flag.item()) == 1
    after = store.cpu().numpy()
    assert np.array_equal(after[[i for i in range(50) if i not in (3, 49)]], arr[[i for i in range(50) if i not in (3, 49)]])
    assert np.all(after[[3, 49], 72:] == 1)
    # the class: an index beyond ITS dataset (but inside the shared store) is caught too
    fd = fits_dict.FitsDict(type('O', (), {'checkpoint_dir': str(tmp_path)})(), None, fits={'a': arr, 'b': arr[:20].copy()})
    p, b = fd[(['a', 'b', 'b'], torch.tensor([49, 19, 20]), torch.zeros(3), torch.zeros(3, dtype=torch.uint8))]
    assert torch.isnan(p[2]).all() and not torch.isnan(p[:2]).any()
    with pytest.raises(IndexError):
        fd.check()
    fd.check()                                                             # the flag is cleared by the raise
    assert fd.fits_dict['b'].shape == (20, 82) and fd.fits_dict['b'].data_ptr() == fd._store[50:].data_ptr()


def test_keep_better():
    gen = torch.Generator().manual_seed(9)
    B = 777
    best = [torch.rand(B, generator=gen) * 100, torch.randn(B, 72, generator=gen), torch.randn(B, 10, generator=gen), torch.randn(B, 3, generator=gen)]
    new = [torch.rand(B, 49, generator=gen) * 100, torch.randn(B, 72, generator=gen), torch.randn(B, 10, generator=gen), torch.randn(B, 3, generator=gen)]
    ref = adjacent.keep_better(*best, *new)
    dev = [t.clone().cuda() for t in best]
    upd = sharded.keep_better_(*dev, *[t.cuda() for t in new])
    margin = (new[0].mean(-1) - best[0]).abs() > 1e-4            # away from ties the decision is identical
    assert torch.equal(upd.cpu()[margin], ref[4][margin])
    same = upd.cpu() == ref[4]
    for a, b in zip(dev[1:], ref[1:4]):
        assert torch.equal(a.cpu()[same], b[same])
    np.testing.assert_allclose(dev[0].cpu().numpy()[same], ref[0].numpy()[same], rtol=1e-6)
