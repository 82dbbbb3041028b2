"""The CPU restatement of the steps either side of SMPLify (oracle/adjacent.py) against the vectors the reference's own
utils/geometry.py and train/fits_dict.py produced (tests/golden/adjacent.npz, oracle/run_reference_adjacent.py)."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import adjacent, tgm_shim

cv2 = pytest.importorskip('cv2')


@pytest.fixture(scope='module')
def g():
    return golden('adjacent.npz')


def test_rot6d(g):
    out = adjacent.rot6d_to_rotmat(torch.from_numpy(g['rot6d_in']))
    assert np.array_equal(out.numpy(), g['rot6d_out'])
    eye = torch.matmul(out.transpose(1, 2), out)
    np.testing.assert_allclose(eye.numpy(), np.broadcast_to(np.eye(3, dtype=np.float32), eye.shape), atol=2e-6)


def test_rotmat_to_axis_angle(g):
    aa = adjacent.rotmat_to_axis_angle(torch.from_numpy(g['rotmat_in']))
    assert np.array_equal(aa.numpy(), g['axis_angle_out'])
    # round trip through torchgeometry's own inverse for the exactly orthonormal rows
    R = tgm_shim.angle_axis_to_rotation_matrix(aa[12:])[:, :3, :3]
    np.testing.assert_allclose(R.numpy(), g['rotmat_in'][12:], atol=3e-6)


def test_estimate_translation(g):
    t = adjacent.estimate_translation(torch.from_numpy(g['et_S']), torch.from_numpy(g['et_kp']))
    np.testing.assert_allclose(t.numpy(), g['et_out'], rtol=1e-6, atol=1e-6)


def test_fits_get_set(g):
    store = torch.from_numpy(g['fits_store'].copy())
    idx, rot, fl = torch.from_numpy(g['fits_index']), torch.from_numpy(g['fits_rot']), torch.from_numpy(g['fits_flipped'])
    pose, betas = adjacent.fits_get(store, idx, rot, fl)
    assert np.array_equal(pose.numpy(), g['fits_get_pose'])
    assert np.array_equal(betas.numpy(), g['fits_get_betas'])
    after = adjacent.fits_set(store.clone(), idx, rot, fl, torch.from_numpy(g['fits_update']), torch.from_numpy(g['fits_new_pose']),
                              torch.from_numpy(g['fits_new_betas']))
    assert np.array_equal(after.numpy(), g['fits_store_after'])
    # rows that were not selected for update are untouched, bit for bit
    untouched = np.ones(store.shape[0], bool)
    untouched[g['fits_index'][g['fits_update'].astype(bool)]] = False
    assert np.array_equal(after.numpy()[untouched], g['fits_store'][untouched])


def test_flip_is_an_involution_and_perm_is_bit_exact(g):
    t = golden('tables.npz')
    assert adjacent.POSE_FLIP_PERM == t['pose_flip_perm'].tolist()
    pose = torch.from_numpy(g['fits_new_pose'])
    ones = torch.ones(pose.shape[0], dtype=torch.uint8)
    assert torch.equal(adjacent.flip_pose(adjacent.flip_pose(pose, ones), ones), pose)
