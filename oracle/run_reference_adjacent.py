"""Generate tests/golden/adjacent.npz by executing the REFERENCE'S OWN utils/geometry.py and train/fits_dict.py.

TEST INFRASTRUCTURE ONLY; runs only where /root/reference exists.  torchgeometry is absent, so train/fits_dict.py is
imported with oracle/tgm_shim.py registered in its place ([recall] restatement); cv2 is the real OpenCV.

    python -m oracle.run_reference_adjacent
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

from . import run_reference, tgm_shim


def main():
    from inbed_pose_estimation_b200 import synthetic
    tmp = tempfile.mkdtemp(prefix='inbed_ref_adj_')
    synthetic.write_data_dir(tmp, seed=0)
    ns = run_reference.import_reference(tmp)
    tgm_shim.install()
    import importlib
    fd_mod = importlib.import_module('train.fits_dict')
    assert os.path.abspath(fd_mod.__file__).startswith(run_reference.REF)
    rs = np.random.RandomState(11)
    out = {}

    # utils/geometry.py:47-61
    x6 = torch.tensor(rs.randn(64, 6).astype(np.float32))
    out['rot6d_in'] = x6.numpy()
    out['rot6d_out'] = ns.geometry.rot6d_to_rotmat(x6).numpy()

    # trainer.py:702-706 on matrices produced by the reference's own batch_rodrigues (+ identity, + near-pi, + a few
    # slightly non-orthonormal ones as a CNN would emit)
    aa = rs.randn(120, 3).astype(np.float32) * np.array([0.2, 1.0, 2.5], dtype=np.float32)[rs.randint(0, 3, 120)][:, None]
    aa[:4] = 0.
    aa[4] = [np.pi - 1e-4, 0, 0]
    aa[5] = [0, 3.1, 0.2]
    R = ns.geometry.batch_rodrigues(torch.tensor(aa))
    R[6:12] += torch.tensor(1e-3 * rs.randn(6, 3, 3).astype(np.float32))
    hom = torch.cat([R, torch.tensor([0, 0, 1], dtype=torch.float32).view(1, 3, 1).expand(R.shape[0], -1, -1)], dim=-1)
    pred_pose = tgm_shim.rotation_matrix_to_angle_axis(hom).contiguous()
    out['aa_nan_count'] = np.asarray(int(torch.isnan(pred_pose).sum()))
    pred_pose[torch.isnan(pred_pose)] = 0.0
    out['rotmat_in'] = R.numpy()
    out['axis_angle_out'] = pred_pose.numpy()

    # utils/geometry.py:156-181
    S = torch.tensor(rs.randn(16, 49, 3).astype(np.float32) * 0.4)
    kp = torch.tensor(np.concatenate([rs.uniform(0, 224, (16, 49, 2)), rs.uniform(0, 1, (16, 49, 1))], axis=-1).astype(np.float32))
    kp[:, 39:, 2] = 0.
    out['et_S'] = S.numpy()
    out['et_kp'] = kp.numpy()
    out['et_out'] = ns.geometry.estimate_translation(S, kp, focal_length=5000., img_size=224.).numpy()

    # train/fits_dict.py: flip_pose / rotate_pose / __getitem__ / __setitem__ through the reference class itself
    fd = fd_mod.FitsDict.__new__(fd_mod.FitsDict)
    fd.flipped_parts = torch.tensor(ns.constants.SMPL_POSE_FLIP_PERM, dtype=torch.int64)
    store = torch.tensor(np.concatenate([0.4 * rs.randn(40, 72), 0.5 * rs.randn(40, 10)], axis=1).astype(np.float32))
    store[3, :3] = 0.                                      # identity global orientation
    store[5, :3] = torch.tensor([3.1, 0.05, -0.02])        # close to pi
    fd.fits_dict = {'slp': store.clone()}
    names = ['slp'] * 24
    idx = torch.tensor(rs.permutation(40)[:24])
    rot = torch.tensor((rs.randn(24) * 30).astype(np.float32))
    rot[:3] = 0.
    flipped = torch.tensor(rs.randint(0, 2, 24).astype(np.uint8))
    pose, betas = fd[(names, idx, rot, flipped)]
    out['fits_store'] = store.numpy()
    out['fits_index'] = idx.numpy()
    out['fits_rot'] = rot.numpy()
    out['fits_flipped'] = flipped.numpy()
    out['fits_get_pose'] = pose.numpy()
    out['fits_get_betas'] = betas.numpy()
    new_pose = torch.tensor((0.4 * rs.randn(24, 72)).astype(np.float32))
    new_betas = torch.tensor((0.5 * rs.randn(24, 10)).astype(np.float32))
    update = torch.tensor(rs.randint(0, 2, 24).astype(np.uint8))
    fd[(names, idx, rot, flipped, update)] = (new_pose, new_betas)
    out['fits_new_pose'] = new_pose.numpy()
    out['fits_new_betas'] = new_betas.numpy()
    out['fits_update'] = update.numpy()
    out['fits_store_after'] = fd.fits_dict['slp'].numpy()
    path = os.path.join(run_reference.GOLDEN, 'adjacent.npz')
    np.savez_compressed(path, **out)
    print('wrote', path, 'NaNs scrubbed in the axis-angle case:', int(out['aa_nan_count']))


if __name__ == '__main__':
    main()
