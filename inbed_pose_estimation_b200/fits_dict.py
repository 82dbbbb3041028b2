"""Best-fit dictionary with the reference's call surface (reference train/fits_dict.py:10-94), device resident.

The reference keeps one CPU tensor [N, 82] (72 pose + 10 betas) per dataset and, on every get / set, loops over the
batch in Python and calls cv2.Rodrigues per sample.  Here the store lives in HBM and each call is one kernel:
gather + rotate + flip (__getitem__), un-flip + un-rotate + masked scatter (__setitem__).  The on-disk format is
unchanged (<checkpoint_dir>/<dataset>_fits.npy, falling back to config.STATIC_FITS_DIR).
"""
import ctypes
import os

import numpy as np
import torch

from . import _native, config, constants

_PERM = np.ascontiguousarray(np.asarray(constants.SMPL_POSE_FLIP_PERM, dtype=np.int32))


def _perm_ptr():
    return ctypes.c_void_p(_PERM.ctypes.data)


def _u8(t, dev):
    return torch.as_tensor(t).to(dev).to(torch.uint8).contiguous()


def fits_get(store, index, rot, is_flipped):
    """store [N,82] (CUDA) -> (pose [B,72], betas [B,10]) for rows `index`, rotated by `rot` degrees then flipped."""
    dev = store.device
    if dev.type != 'cuda':
        raise RuntimeError('FitsDict store must live on a CUDA device (no CPU fallback)')
    idx = torch.as_tensor(index).to(dev).to(torch.int64).contiguous()
    B = idx.shape[0]
    r = torch.as_tensor(rot).to(dev).float().contiguous()
    f = _u8(is_flipped, dev)
    pose = torch.empty((B, 72), device=dev, dtype=torch.float32)
    betas = torch.empty((B, 10), device=dev, dtype=torch.float32)
    if B:
        if int(idx.min()) < 0 or int(idx.max()) >= store.shape[0]:
            raise IndexError('fits index out of range')
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_fits_get(
                B, _native.ptr(store), ctypes.c_void_p(idx.data_ptr()), _native.ptr(r), ctypes.c_void_p(f.data_ptr()), _perm_ptr(),
                _native.ptr(pose), _native.ptr(betas), torch.cuda.current_stream(dev).cuda_stream))
    return pose, betas


def fits_set(store, index, rot, is_flipped, update, pose, betas):
    """Undo flip and rotation of (pose, betas) and overwrite the rows of `store` where `update` is set."""
    dev = store.device
    idx = torch.as_tensor(index).to(dev).to(torch.int64).contiguous()
    B = idx.shape[0]
    if not B:
        return
    if int(idx.min()) < 0 or int(idx.max()) >= store.shape[0]:
        raise IndexError('fits index out of range')
    r = torch.as_tensor(rot).to(dev).float().contiguous()
    f, u = _u8(is_flipped, dev), _u8(update, dev)
    p = pose.detach().to(dev).float().contiguous()
    b = betas.detach().to(dev).float().contiguous()
    with torch.cuda.device(dev):
        _native.check(_native.lib().smplb200_fits_set(
            B, _native.ptr(store), ctypes.c_void_p(idx.data_ptr()), _native.ptr(r), ctypes.c_void_p(f.data_ptr()),
            ctypes.c_void_p(u.data_ptr()), _perm_ptr(), _native.ptr(p), _native.ptr(b), torch.cuda.current_stream(dev).cuda_stream))


class FitsDict(object):
    """ Dictionary keeping track of the best fit per image in the training set (device resident). """

    def __init__(self, options, train_dataset, device=torch.device('cuda'), fits=None):
        """`fits`: optional {dataset name: array [N, 82]} to use instead of the .npy files."""
        self.options = options
        self.train_dataset = train_dataset
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('inbed_pose_estimation_b200.FitsDict keeps its store on a CUDA device (no CPU fallback)')
        self.fits_dict = {}
        self.flipped_parts = torch.tensor(constants.SMPL_POSE_FLIP_PERM, dtype=torch.int64)
        names = list(fits.keys()) if fits is not None else list(train_dataset.dataset_dict.keys())
        for ds_name in names:
            if fits is not None:
                arr = np.asarray(fits[ds_name], dtype=np.float32)
            else:
                try:
                    arr = np.load(os.path.join(options.checkpoint_dir, ds_name + '_fits.npy'))
                except IOError:
                    # Dictionary does not exist, so populate with static fits
                    arr = np.load(os.path.join(config.STATIC_FITS_DIR, ds_name + '_fits.npy'))
            if arr.ndim != 2 or arr.shape[1] != 82:
                raise ValueError('%s fits must be [N, 82]' % ds_name)
            self.fits_dict[ds_name] = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(self.device)

    def save(self):
        """ Save dictionary state to disk """
        for ds_name, store in self.fits_dict.items():
            np.save(os.path.join(self.options.checkpoint_dir, ds_name + '_fits.npy'), store.cpu().numpy())

    def _groups(self, dataset_name):
        groups = {}
        for n, ds in enumerate(dataset_name):
            groups.setdefault(ds, []).append(n)
        return groups

    def __getitem__(self, x):
        """ Retrieve dictionary entries: (pose [B,72], betas [B,10]) on the store's device """
        dataset_name, ind, rot, is_flipped = x
        B = len(dataset_name)
        ind, rot = torch.as_tensor(ind), torch.as_tensor(rot)
        flipped = torch.as_tensor(is_flipped)
        pose = torch.empty((B, 72), device=self.device, dtype=torch.float32)
        betas = torch.empty((B, 10), device=self.device, dtype=torch.float32)
        for ds, rows in self._groups(dataset_name).items():
            rows_t = torch.as_tensor(rows)
            p, b = fits_get(self.fits_dict[ds], ind[rows_t], rot[rows_t], flipped[rows_t])
            pose[rows_t.to(self.device)] = p
            betas[rows_t.to(self.device)] = b
        return pose, betas

    def __setitem__(self, x, val):
        """ Update dictionary entries """
        dataset_name, ind, rot, is_flipped, update = x
        pose, betas = val
        ind, rot = torch.as_tensor(ind), torch.as_tensor(rot)
        flipped, update = torch.as_tensor(is_flipped), torch.as_tensor(update)
        for ds, rows in self._groups(dataset_name).items():
            rows_t = torch.as_tensor(rows)
            rd = rows_t.to(pose.device)
            fits_set(self.fits_dict[ds], ind[rows_t], rot[rows_t], flipped[rows_t], update[rows_t], pose[rd], betas[rd])

    def flip_pose(self, pose, is_flipped):
        """flip SMPL pose parameters (through the get kernel on a scratch store)"""
        B = pose.shape[0]
        store = torch.cat([pose.detach().to(self.device).float(), torch.zeros((B, 10), device=self.device)], dim=1).contiguous()
        p, _ = fits_get(store, torch.arange(B), torch.zeros(B), is_flipped)
        return p

    def rotate_pose(self, pose, rot):
        """Rotate SMPL pose parameters by rot degrees"""
        B = pose.shape[0]
        store = torch.cat([pose.detach().to(self.device).float(), torch.zeros((B, 10), device=self.device)], dim=1).contiguous()
        p, _ = fits_get(store, torch.arange(B), rot, torch.zeros(B, dtype=torch.uint8))
        return p
