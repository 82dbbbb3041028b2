"""Best-fit dictionary with the reference's call surface (reference train/fits_dict.py:10-94), device resident.

The reference keeps one CPU tensor [N, 82] (72 pose + 10 betas) per dataset and, on every get / set, loops over the
batch in Python and calls cv2.Rodrigues per sample.  Here the store lives in HBM and each call is one kernel:
gather + rotate + flip (__getitem__), un-flip + un-rotate + masked scatter (__setitem__); index bounds are validated
inside the kernel and reported through a device flag, so neither call synchronises with the host.  The on-disk format is
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


def _status_flag(dev):
    return torch.zeros(1, device=dev, dtype=torch.int32)


def check_status(status):
    """Raise IndexError if any fits_get / fits_set call that was given `status` saw an out-of-range index.  This is the
    one host synchronisation of the bounds check; call it whenever convenient (FitsDict does so in save())."""
    if status is not None and int(status.item()) != 0:
        status.zero_()
        raise IndexError('fits index out of range')


def fits_get(store, index, rot, is_flipped, status=None):
    """store [N,82] (CUDA) -> (pose [B,72], betas [B,10]) for rows `index`, rotated by `rot` degrees then flipped.
    Indices are checked on the device: rows outside [0, N) come back as NaN and raise `status` (an int32 device tensor,
    see check_status); without `status` the check is made here, at the cost of a host synchronisation."""
    dev = store.device
    if dev.type != 'cuda':
        raise RuntimeError('FitsDict store must live on a CUDA device (no CPU fallback)')
    idx = torch.as_tensor(index).to(dev, torch.int64, non_blocking=True).contiguous()
    B = idx.shape[0]
    r = torch.as_tensor(rot).to(dev, torch.float32, non_blocking=True).contiguous()
    f = _u8(is_flipped, dev)
    pose = torch.empty((B, 72), device=dev, dtype=torch.float32)
    betas = torch.empty((B, 10), device=dev, dtype=torch.float32)
    if B:
        flag = status if status is not None else _status_flag(dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_fits_get(
                B, _native.ptr(store), int(store.shape[0]), ctypes.c_void_p(idx.data_ptr()), _native.ptr(r),
                ctypes.c_void_p(f.data_ptr()), _perm_ptr(), _native.ptr(pose), _native.ptr(betas), ctypes.c_void_p(flag.data_ptr()),
                torch.cuda.current_stream(dev).cuda_stream))
        if status is None:
            check_status(flag)
    return pose, betas


def fits_set(store, index, rot, is_flipped, update, pose, betas, status=None):
    """Undo flip and rotation of (pose, betas) and overwrite the rows of `store` where `update` is set.  Out-of-range rows
    are skipped and reported through `status` like in fits_get."""
    dev = store.device
    idx = torch.as_tensor(index).to(dev, torch.int64, non_blocking=True).contiguous()
    B = idx.shape[0]
    if not B:
        return
    r = torch.as_tensor(rot).to(dev, torch.float32, non_blocking=True).contiguous()
    f, u = _u8(is_flipped, dev), _u8(update, dev)
    p = pose.detach().to(dev).float().contiguous()
    b = betas.detach().to(dev).float().contiguous()
    flag = status if status is not None else _status_flag(dev)
    with torch.cuda.device(dev):
        _native.check(_native.lib().smplb200_fits_set(
            B, _native.ptr(store), int(store.shape[0]), ctypes.c_void_p(idx.data_ptr()), _native.ptr(r), ctypes.c_void_p(f.data_ptr()),
            ctypes.c_void_p(u.data_ptr()), _perm_ptr(), _native.ptr(p), _native.ptr(b), ctypes.c_void_p(flag.data_ptr()),
            torch.cuda.current_stream(dev).cuda_stream))
    if status is None:
        check_status(flag)


class FitsDict(object):
    """ Dictionary keeping track of the best fit per image in the training set (device resident).

    All datasets share ONE device store (their rows back to back; `fits_dict[name]` is a view of its rows), so a mixed
    batch is one kernel launch: the per-dataset row offset is added to the indices on the host, where the dataset names
    live anyway.  Index bounds are checked inside the kernels; the sticky device flag is read by check() / save(). """

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
        arrays = []
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
            arrays.append(np.ascontiguousarray(arr, dtype=np.float32))
        self._offset, self._rows = {}, {}
        row = 0
        for ds_name, arr in zip(names, arrays):
            self._offset[ds_name], self._rows[ds_name] = row, arr.shape[0]
            row += arr.shape[0]
        self._store = torch.from_numpy(np.concatenate(arrays, axis=0) if arrays else np.zeros((0, 82), np.float32)).to(self.device)
        for ds_name in names:
            self.fits_dict[ds_name] = self._store[self._offset[ds_name]:self._offset[ds_name] + self._rows[ds_name]]
        self._status = _status_flag(self.device)

    def check(self):
        """Raise IndexError if a get / set since the last check used an index outside its dataset (one host sync)."""
        check_status(self._status)

    def save(self):
        """ Save dictionary state to disk """
        self.check()
        for ds_name, store in self.fits_dict.items():
            np.save(os.path.join(self.options.checkpoint_dir, ds_name + '_fits.npy'), store.cpu().numpy())

    def _global_index(self, dataset_name, ind):
        """Row of the shared store for every (dataset, index) pair; -1 where the index leaves its dataset, so that the
        kernel's bounds check reports it (host arithmetic on host data: the names are Python strings)."""
        ind = np.asarray(torch.as_tensor(ind).cpu(), dtype=np.int64).reshape(-1)
        off = np.fromiter((self._offset[d] for d in dataset_name), dtype=np.int64, count=len(dataset_name))
        rows = np.fromiter((self._rows[d] for d in dataset_name), dtype=np.int64, count=len(dataset_name))
        return torch.from_numpy(np.where((ind >= 0) & (ind < rows), ind + off, -1))

    def __getitem__(self, x):
        """ Retrieve dictionary entries: (pose [B,72], betas [B,10]) on the store's device """
        dataset_name, ind, rot, is_flipped = x
        return fits_get(self._store, self._global_index(dataset_name, ind), rot, is_flipped, status=self._status)

    def __setitem__(self, x, val):
        """ Update dictionary entries """
        dataset_name, ind, rot, is_flipped, update = x
        pose, betas = val
        fits_set(self._store, self._global_index(dataset_name, ind), rot, is_flipped, update, pose, betas, status=self._status)

    def flip_pose(self, pose, is_flipped):
        """flip SMPL pose parameters (through the get kernel on a scratch store)"""
        B = pose.shape[0]
        store = torch.cat([pose.detach().to(self.device).float(), torch.zeros((B, 10), device=self.device)], dim=1).contiguous()
        p, _ = fits_get(store, torch.arange(B), torch.zeros(B), is_flipped, status=self._status)
        return p

    def rotate_pose(self, pose, rot):
        """Rotate SMPL pose parameters by rot degrees"""
        B = pose.shape[0]
        store = torch.cat([pose.detach().to(self.device).float(), torch.zeros((B, 10), device=self.device)], dim=1).contiguous()
        p, _ = fits_get(store, torch.arange(B), rot, torch.zeros(B, dtype=torch.uint8), status=self._status)
        return p
