"""ctypes binding of libsmplify_b200.so (include/smplify_b200.h).

The shared library is built in-tree by ``build()`` with nvcc for sm_100a and loaded with
ctypes; torch only supplies device memory, the current stream and autograd plumbing.
There is no fallback: if the library is missing and cannot be built, or no CUDA device
is present when a model is created, the calls raise.
"""
import ctypes
import os
import shutil
import subprocess
import threading

import numpy as np

from . import constants as C

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, 'csrc')
LIB_PATH = os.environ.get('SMPLB200_LIB') or os.path.join(_HERE, 'libsmplify_b200.so')   # override: profiling builds only
SOURCES = ['kernels.cu', 'lbs_tc.cu', 'adjacent.cu', 'train_losses.cu', 'api.cu', 'probe.cu', 'model_host.cpp']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']

VPOSED_PITCH = 20736     # SMPLB200_VPOSED_PITCH: floats per row of the saved v_posed buffer

_lock = threading.Lock()
_lib = None


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    return None


def _source_hash():
    """Digest of everything the library is built from (file names + contents): staleness must not depend on file times,
    which a copy of the tree (the GPU box's snapshot) does not preserve."""
    import hashlib
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    deps.append(os.path.join(os.path.dirname(_HERE), 'include', 'smplify_b200.h'))
    for d in deps:
        if os.path.isfile(d):
            h.update(os.path.basename(d).encode())
            with open(d, 'rb') as f:
                h.update(f.read())
    h.update(' '.join(NVCC_FLAGS + SOURCES).encode())
    return h.hexdigest()


def _stale(lib_path=None):
    lib_path = lib_path or LIB_PATH
    if not os.path.exists(lib_path):
        return True
    try:
        with open(lib_path + '.srchash') as f:
            return f.read().strip() != _source_hash()
    except OSError:
        return True


def build(force=False, verbose=False, extra_flags=(), out=None):
    """Compile the CUDA library for sm_100a into the package directory.  Safe to call from several processes at once
    (one rank per GPU): an exclusive file lock serialises the builders and the later ones find the library fresh."""
    import fcntl
    out = out or LIB_PATH
    if out == LIB_PATH and not force and not _stale():
        return LIB_PATH
    with open(out + '.lock', 'w') as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if out == LIB_PATH and not force and not _stale():
                return LIB_PATH                                  # another process built it while we waited
            nvcc = _nvcc()
            if nvcc is None:
                raise RuntimeError('nvcc not found: cannot build %s' % out)
            tmp = '%s.tmp.%d' % (out, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + list(extra_flags) + (['-Xptxas', '-v'] if verbose else []) + \
                ['-o', tmp] + [os.path.join(CSRC, s) for s in SOURCES]
            res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, universal_newlines=True)
            if res.returncode != 0:
                raise RuntimeError('nvcc failed:\n' + res.stdout)
            os.replace(tmp, out)
            if not extra_flags:
                with open(out + '.srchash', 'w') as f:
                    f.write(_source_hash())
            if verbose:
                print(res.stdout)
            return out
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


_f32p = ctypes.POINTER(ctypes.c_float)
_i32p = ctypes.POINTER(ctypes.c_int32)


class ModelDesc(ctypes.Structure):
    _fields_ = [
        ('v_template', _f32p), ('shapedirs', _f32p), ('posedirs', _f32p), ('J_regressor', _f32p),
        ('weights', _f32p), ('J_regressor_extra', _f32p), ('parents', _i32p),
        ('extra_vertex_ids', _i32p), ('joint_map', _i32p), ('ign_joints', _i32p),
        ('num_ign_joints', ctypes.c_int32), ('cam_op_joints', _i32p), ('cam_gt_joints', _i32p),
        ('angle_prior_ids', _i32p), ('angle_prior_signs', _f32p), ('gmm_means', _f32p),
        ('gmm_precisions', _f32p), ('gmm_nll_weights', _f32p),
    ]


def make_desc(arrays, prior=None):
    """(ModelDesc, keepalive list) from the numpy arrays of an SMPL model.

    arrays: dict with v_template [6890,3], shapedirs [6890,3,10], posedirs [6890,3,207],
            J_regressor [24,6890], weights [6890,24], J_regressor_extra [9,6890], parents [24].
    prior:  None or dict with means [8,69], precisions [8,69,69], nll_weights [8] (fp32)."""
    keep = []

    def f32(a, shape):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
        if tuple(a.shape) != tuple(shape):
            raise ValueError('expected shape %s, got %s' % (shape, a.shape))
        keep.append(a)
        return a.ctypes.data_as(_f32p)

    def i32(a, n):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.int32))
        if a.shape != (n,):
            raise ValueError('expected %d int entries, got %s' % (n, a.shape))
        keep.append(a)
        return a.ctypes.data_as(_i32p)

    V, J = C.NUM_VERTS, C.NUM_SMPL_JOINTS
    d = ModelDesc()
    d.v_template = f32(arrays['v_template'], (V, 3))
    d.shapedirs = f32(arrays['shapedirs'], (V, 3, C.NUM_BETAS))
    d.posedirs = f32(arrays['posedirs'], (V, 3, C.NUM_POSE_FEATURES))
    d.J_regressor = f32(arrays['J_regressor'], (J, V))
    d.weights = f32(arrays['weights'], (V, J))
    d.J_regressor_extra = f32(arrays['J_regressor_extra'], (9, V))
    parents = np.asarray(arrays['parents']).astype(np.int64).copy()
    parents[0] = -1
    d.parents = i32(parents, J)
    d.extra_vertex_ids = i32(C.SMPL_EXTRA_VERTEX_IDS, 21)
    d.joint_map = i32([C.JOINT_MAP[n] for n in C.JOINT_NAMES], C.NUM_JOINTS_OUT)
    d.ign_joints = i32(C.SMPLIFY_IGNORED_JOINTS, len(C.SMPLIFY_IGNORED_JOINTS))
    d.num_ign_joints = len(C.SMPLIFY_IGNORED_JOINTS)
    d.cam_op_joints = i32(C.CAMERA_OP_JOINTS, 4)
    d.cam_gt_joints = i32(C.CAMERA_GT_JOINTS, 4)
    d.angle_prior_ids = i32(C.ANGLE_PRIOR_IDS, 4)
    d.angle_prior_signs = f32(C.ANGLE_PRIOR_SIGNS, (4,))
    if prior is not None:
        d.gmm_means = f32(prior['means'], (8, 69))
        d.gmm_precisions = f32(prior['precisions'], (8, 69, 69))
        d.gmm_nll_weights = f32(np.asarray(prior['nll_weights']).reshape(-1), (8,))
    return d, keep


def _declare(lib):
    vp, sz, ci, cf, cd, i64 = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_int64
    lib.smplb200_version.restype = ci
    lib.smplb200_last_error.restype = ctypes.c_char_p
    lib.smplb200_launch_count.restype = ctypes.c_longlong
    lib.smplb200_launch_count.argtypes = [ci]
    lib.smplb200_model_create.restype = ci
    lib.smplb200_model_create.argtypes = [ctypes.POINTER(ModelDesc), ci, ctypes.POINTER(vp)]
    lib.smplb200_model_destroy.restype = None
    lib.smplb200_model_destroy.argtypes = [vp]
    for name in ('smplb200_fit_workspace_bytes', 'smplb200_smpl_workspace_bytes'):
        getattr(lib, name).restype = sz
        getattr(lib, name).argtypes = [ci]
    lib.smplb200_smplify_fit.restype = ci
    lib.smplb200_smplify_fit.argtypes = [vp, ci, ci, cd, cf] + [vp] * 13 + [vp, sz, vp]
    lib.smplb200_smplify_fitting_loss.restype = ci
    lib.smplb200_smplify_fitting_loss.argtypes = [vp, ci, cf] + [vp] * 6 + [vp, sz, vp]
    lib.smplb200_prior_terms.restype = ci
    lib.smplb200_prior_terms.argtypes = [vp, ci] + [vp] * 8
    lib.smplb200_smpl_forward.restype = ci
    lib.smplb200_smpl_forward.argtypes = [vp, ci, ci] + [vp] * 5 + [vp, sz, vp]
    lib.smplb200_smpl_backward.restype = ci
    lib.smplb200_smpl_backward.argtypes = [vp, ci, ci] + [vp] * 7 + [vp, sz, vp]
    lib.smplb200_batch_rodrigues.restype = ci
    lib.smplb200_batch_rodrigues.argtypes = [ci, vp, vp, vp]
    lib.smplb200_batch_rodrigues_backward.restype = ci
    lib.smplb200_batch_rodrigues_backward.argtypes = [ci, vp, vp, vp, vp]
    lib.smplb200_perspective_projection.restype = ci
    lib.smplb200_perspective_projection.argtypes = [ci, ci, vp, vp, vp, vp, ci, vp, ci, vp, vp]
    lib.smplb200_perspective_projection_backward.restype = ci
    lib.smplb200_perspective_projection_backward.argtypes = [ci, ci, vp, vp, vp, vp, ci, ci, vp, vp, vp, vp, vp]
    lib.smplb200_probe_fp32_peak.restype = ci
    lib.smplb200_probe_fp32_peak.argtypes = [ci, ctypes.POINTER(ctypes.c_double)]
    lib.smplb200_probe_tf32_peak.restype = ci
    lib.smplb200_probe_tf32_peak.argtypes = [ctypes.POINTER(ctypes.c_double)]
    lib.smplb200_rot6d_to_rotmat.restype = ci
    lib.smplb200_rot6d_to_rotmat.argtypes = [ci, vp, vp, vp]
    lib.smplb200_rotmat_to_axis_angle.restype = ci
    lib.smplb200_rotmat_to_axis_angle.argtypes = [ci, vp, vp, ci, vp]
    lib.smplb200_estimate_translation.restype = ci
    lib.smplb200_estimate_translation.argtypes = [ci, vp, vp, cf, cf, vp, vp]
    lib.smplb200_fits_get.restype = ci
    lib.smplb200_fits_get.argtypes = [ci, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.smplb200_fits_set.restype = ci
    lib.smplb200_fits_set.argtypes = [ci, vp, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.smplb200_keep_better.restype = ci
    lib.smplb200_keep_better.argtypes = [ci] + [vp] * 12
    lib.smplb200_finalize_fits.restype = ci
    lib.smplb200_finalize_fits.argtypes = [ci, cf] + [vp] * 14
    lib.smplb200_train_loss_workspace_bytes.restype = sz
    lib.smplb200_train_loss_workspace_bytes.argtypes = [ci]
    lib.smplb200_smpl_param_losses.restype = ci
    lib.smplb200_smpl_param_losses.argtypes = [ci] + [vp] * 10
    lib.smplb200_keypoint_loss.restype = ci
    lib.smplb200_keypoint_loss.argtypes = [ci, vp, vp, cf, cf, vp, vp, vp, vp]
    lib.smplb200_keypoint_3d_loss.restype = ci
    lib.smplb200_keypoint_3d_loss.argtypes = [ci] + [vp] * 7
    lib.smplb200_shape_loss.restype = ci
    lib.smplb200_shape_loss.argtypes = [ci] + [vp] * 7
    lib.smplb200_weak_perspective_projection.restype = ci
    lib.smplb200_weak_perspective_projection.argtypes = [ci, ci, vp, vp, cf, cf, vp, vp, vp]
    lib.smplb200_weak_perspective_projection_backward.restype = ci
    lib.smplb200_weak_perspective_projection_backward.argtypes = [ci, ci, vp, vp, cf, cf, vp, vp, vp, vp, vp]
    lib.smplb200_fit_split_plan.restype = ci
    lib.smplb200_fit_split_plan.argtypes = [ci, ci]
    lib.smplb200_fit_pair_plan.restype = ci
    lib.smplb200_fit_pair_plan.argtypes = [ci, ci] + [ctypes.POINTER(ci)] * 2
    lib.smplb200_fit_tile_plan.restype = None
    lib.smplb200_fit_tile_plan.argtypes = [ci, ci] + [ctypes.POINTER(ci)] * 3
    lib.smplb200_smplify_fit_host.restype = ci
    lib.smplb200_smplify_fit_host.argtypes = [vp, ci, ci, cd, cf] + [vp] * 11
    return lib


EXPORTED_SYMBOLS = (
    'smplb200_version', 'smplb200_last_error', 'smplb200_model_create', 'smplb200_model_destroy',
    'smplb200_fit_workspace_bytes', 'smplb200_smpl_workspace_bytes', 'smplb200_smplify_fit',
    'smplb200_smplify_fitting_loss', 'smplb200_smpl_forward', 'smplb200_smpl_backward',
    'smplb200_batch_rodrigues', 'smplb200_batch_rodrigues_backward', 'smplb200_perspective_projection',
    'smplb200_perspective_projection_backward', 'smplb200_smplify_fit_host', 'smplb200_launch_count', 'smplb200_probe_fp32_peak',
    'smplb200_rot6d_to_rotmat', 'smplb200_rotmat_to_axis_angle', 'smplb200_estimate_translation', 'smplb200_fits_get',
    'smplb200_fits_set', 'smplb200_keep_better', 'smplb200_finalize_fits', 'smplb200_train_loss_workspace_bytes',
    'smplb200_fit_tile_plan', 'smplb200_weak_perspective_projection', 'smplb200_weak_perspective_projection_backward',
    'smplb200_smpl_param_losses', 'smplb200_keypoint_loss', 'smplb200_keypoint_3d_loss', 'smplb200_shape_loss',
    'smplb200_prior_terms', 'smplb200_probe_tf32_peak', 'smplb200_fit_pair_plan', 'smplb200_fit_split_plan',
)


def lib():
    """The loaded library (built on first use when nvcc is available)."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.environ.get('SMPLB200_LIB') and _stale():      # an explicitly named library is loaded as it is
                build()
            _lib = _declare(ctypes.CDLL(LIB_PATH))
        return _lib


def fit_tile_plan(batch, sms=148):
    """(n16, small, n_small): tiles of 16 samples, then n_small tiles of `small` samples (smplb200_fit_tile_plan)."""
    a, b, c = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    lib().smplb200_fit_tile_plan(int(batch), int(sms), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
    return a.value, b.value, c.value


def fit_split_plan(batch, sms=0):
    """CTAs per 4-sample tile of the small-batch cluster kernel for this batch (8 / 4 / 2), 0 = the tile kernels run it.
    sms = 0: what the launch uses on the current CUDA device (cluster occupancy); sms > 0: by SM count alone (upper bound)."""
    return int(lib().smplb200_fit_split_plan(int(batch), int(sms)))


def fit_pair_plan(batch, sms=148):
    """(uses_pairs, n16, n12): whether smplb200_smplify_fit runs `batch` on the pair kernel, and its 2x16- / 2x12-sample pairs."""
    a, b = ctypes.c_int(0), ctypes.c_int(0)
    used = lib().smplb200_fit_pair_plan(int(batch), int(sms), ctypes.byref(a), ctypes.byref(b))
    return bool(used), a.value, b.value


def check(rc):
    if rc != 0:
        raise RuntimeError('libsmplify_b200: ' + lib().smplb200_last_error().decode('utf-8', 'replace'))


def ptr(t):
    """Device (or host) address of a contiguous fp32 torch tensor / numpy array, or NULL."""
    if t is None:
        return None
    if isinstance(t, np.ndarray):
        assert t.dtype == np.float32 and t.flags['C_CONTIGUOUS']
        return ctypes.c_void_p(t.ctypes.data)
    assert t.is_contiguous() and t.dtype.is_floating_point and t.element_size() == 4, 'need contiguous fp32'
    return ctypes.c_void_p(t.data_ptr())


class NativeModel(object):
    """Owner of one device-resident constant blob (smplb200_model)."""

    def __init__(self, arrays, prior=None, device_index=0):
        desc, keep = make_desc(arrays, prior)
        handle = ctypes.c_void_p()
        check(lib().smplb200_model_create(ctypes.byref(desc), int(device_index), ctypes.byref(handle)))
        del keep
        self.handle = handle
        self.device_index = int(device_index)
        self.has_prior = prior is not None

    def close(self):
        if getattr(self, 'handle', None):
            lib().smplb200_model_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
