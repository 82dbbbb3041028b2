"""Builds and binds the TEST-ONLY host emulation of the tile code (tests/emu/emu_api.cu)."""
import ctypes
import os
import shutil
import subprocess

import numpy as np

from inbed_pose_estimation_b200 import _native, synthetic

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.environ.get('SMPLB200_EMU_LIB') or os.path.join(HERE, 'libsmplify_emu.so')   # override: numerics experiments
_vp, _ci, _cf, _cd = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_double


def available():
    return shutil.which('nvcc') is not None or os.path.exists('/usr/local/cuda/bin/nvcc')


def _build():
    if os.environ.get('SMPLB200_EMU_LIB'):
        return
    csrc = _native.CSRC
    deps = [os.path.join(csrc, f) for f in os.listdir(csrc)] + [os.path.join(HERE, 'emu_api.cu')]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    cmd = [nvcc, '-O2', '-std=c++17', '-DSMPLB200_EMU', '-gencode', 'arch=compute_100a,code=sm_100a', '-shared', '-Xcompiler', '-fPIC',
           '-o', LIB, os.path.join(HERE, 'emu_api.cu'), os.path.join(csrc, 'model_host.cpp')]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
    if res.returncode != 0:
        raise RuntimeError(res.stdout)


def model_arrays(seed=0, structure='dense'):
    m = synthetic.make_smpl_model(seed, structure)
    arrays = {k: np.asarray(m[k], dtype=np.float32) for k in ('v_template', 'shapedirs', 'posedirs', 'J_regressor', 'weights')}
    arrays['J_regressor_extra'] = synthetic.make_extra_regressor(seed + 1, structure)
    arrays['parents'] = np.asarray(m['kintree_table'][0]).astype(np.int64)
    return arrays


def prior_arrays(seed=0):
    from inbed_pose_estimation_b200.prior import gmm_constants
    return gmm_constants(synthetic.make_gmm(seed + 2))


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


class Emu(object):
    def __init__(self, seed=0, structure='dense', prior_seed=None):
        _build()
        self.lib = ctypes.CDLL(LIB)
        self.lib.emu_model_create.restype = _vp
        self.lib.emu_model_create.argtypes = [ctypes.POINTER(_native.ModelDesc)]
        self.lib.emu_last_error.restype = ctypes.c_char_p
        self.lib.emu_fit.argtypes = [_vp, _ci, _ci, _cd, _cf, _ci] + [_vp] * 13
        self.lib.emu_prior.argtypes = [_vp, _ci] + [_vp] * 7
        self.lib.emu_pose.argtypes = [_vp, _ci, _ci, _ci] + [_vp] * 8 + [_ci, _vp, _vp]
        self.lib.emu_model_array.restype = ctypes.POINTER(ctypes.c_float)
        self.lib.emu_model_array.argtypes = [_vp, ctypes.c_char_p, ctypes.POINTER(_ci)]
        desc, keep = _native.make_desc(model_arrays(seed, structure), prior_arrays(seed if prior_seed is None else prior_seed))
        self.model = self.lib.emu_model_create(ctypes.byref(desc))
        if not self.model:
            raise RuntimeError(self.lib.emu_last_error().decode())

    def array(self, name):
        n = _ci()
        p = self.lib.emu_model_array(self.model, name.encode(), ctypes.byref(n))
        return np.ctypeslib.as_array(p, shape=(n.value,)).copy()

    def fit(self, inp, num_iters=100, step_size=1e-2, focal=5000., loss_only=False, want_trace=True):
        B = inp['pose'].shape[0]
        c = lambda a: np.ascontiguousarray(a, dtype=np.float32).copy()
        pose, betas, cam, cen, kp = (c(inp[k]) for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints'))
        out = {'joints': np.zeros((B, 49, 3), np.float32), 'pose': np.zeros((B, 72), np.float32),
               'betas': np.zeros((B, 10), np.float32), 'cam_t': np.zeros((B, 3), np.float32),
               'reproj': np.zeros((B, 49), np.float32), 'A': np.zeros((B, 24, 12), np.float32),
               'x': np.zeros((B, 224), np.float32)}
        trace = np.zeros((2 * num_iters, B), np.float32) if (want_trace and not loss_only) else None
        self.lib.emu_fit(self.model, B, num_iters, step_size, focal, int(loss_only), _p(pose), _p(betas), _p(cam), _p(cen),
                         _p(kp), _p(out['joints']), _p(out['pose']), _p(out['betas']), _p(out['cam_t']), _p(out['reproj']),
                         _p(trace), _p(out['A']), _p(out['x']))
        out['trace'] = trace
        out['keypoints'] = kp
        return out

    def prior_terms(self, pose, betas):
        B = pose.shape[0]
        pose, betas = np.ascontiguousarray(pose, np.float32), np.ascontiguousarray(betas, np.float32)
        out = {'terms': np.zeros((B, 3), np.float32), 'components': np.zeros((B, 8), np.float32), 'argmin': np.zeros(B, np.int32),
               'grad_body_pose': np.zeros((B, 69), np.float32), 'grad_betas': np.zeros((B, 10), np.float32)}
        self.lib.emu_prior(self.model, B, _p(pose), _p(betas), _p(out['terms']), _p(out['components']), _p(out['argmin']),
                           _p(out['grad_body_pose']), _p(out['grad_betas']))
        return out

    def pose_forward(self, pose, betas, rotmat_mode=False):
        B = betas.shape[0]
        pose = np.ascontiguousarray(pose, np.float32)
        betas = np.ascontiguousarray(betas, np.float32)
        joints = np.zeros((B, 49, 3), np.float32)
        A = np.zeros((B, 24, 12), np.float32)
        x = np.zeros((B, 224), np.float32)
        self.lib.emu_pose(self.model, B, int(rotmat_mode), 0, _p(pose), _p(betas), _p(joints), _p(A), _p(x),
                          None, None, None, 0, None, None)
        return joints, A, x

    def pose_backward(self, pose, betas, d_joints=None, dA=None, dx=None, rotmat_mode=False):
        B = betas.shape[0]
        pose = np.ascontiguousarray(pose, np.float32)
        betas = np.ascontiguousarray(betas, np.float32)
        dj = None if d_joints is None else np.ascontiguousarray(d_joints, np.float32)
        dA = None if dA is None else np.ascontiguousarray(dA, np.float32)
        dx = None if dx is None else np.ascontiguousarray(dx, np.float32)
        d_pose = np.zeros((B, 216 if rotmat_mode else 72), np.float32)
        d_betas = np.zeros((B, 10), np.float32)
        self.lib.emu_pose(self.model, B, int(rotmat_mode), 1, _p(pose), _p(betas), None, None, None,
                          _p(dj), _p(dA), _p(dx), 1 if dA is not None else 0, _p(d_pose), _p(d_betas))
        return d_pose, d_betas
