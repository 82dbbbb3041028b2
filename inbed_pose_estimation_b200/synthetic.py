"""Seeded synthetic SMPL-shaped model, GMM prior and fit inputs.

The real SMPL pkl, gmm_08.pkl and the SLP data cannot be shipped, so every test
and benchmark runs on a random model of SMPL's exact shape: 6890 vertices,
24 joints, 10 betas, 207 pose features, dense regressors and skinning weights.
All arrays come from ``numpy.random.RandomState`` (legacy, bit-stable across
numpy versions) so the GPU box regenerates the same model from the seed.
"""
import os
import pickle

import numpy as np

from . import constants as C


def _sparse_rows(rs, rows, cols, max_nnz, row_sum_jitter=0.0):
    """[rows, cols] matrix with 1..max_nnz positive entries per row at random columns and exact zeros elsewhere; rows sum
    to 1, or to 1 + row_sum_jitter * N(0,1) (a regressor whose rows are not exactly affine)."""
    m = np.zeros((rows, cols))
    for r in range(rows):
        nnz = rs.randint(1, max_nnz + 1)
        idx = rs.choice(cols, size=nnz, replace=False)
        w = rs.rand(nnz) + 0.05
        m[r, idx] = w / w.sum()
        if row_sum_jitter:
            m[r] *= 1.0 + row_sum_jitter * rs.randn()
    return m


def make_smpl_model(seed=0, structure='dense'):
    """Dictionary with the keys of an SMPL model pkl (float64, like the file).

    structure 'dense' (SURVEY.md 8d): every regressor / skinning entry is non-zero.
    structure 'sparse': the sparsity pattern of the real SMPL file - a J_regressor with at most 10 non-zero vertices per
    joint, at most 4 non-zero skinning weights per vertex, exact zeros everywhere else."""
    rs = np.random.RandomState(seed)
    V, J = C.NUM_VERTS, C.NUM_SMPL_JOINTS
    v_template = 0.3 * rs.randn(V, 3)
    shapedirs = 0.01 * rs.randn(V, 3, C.NUM_BETAS)
    posedirs = 0.001 * rs.randn(V, 3, C.NUM_POSE_FEATURES)
    if structure == 'dense':
        j_reg = rs.rand(J, V) ** 20
        j_reg /= j_reg.sum(axis=1, keepdims=True)
        weights = rs.rand(V, J) ** 8
        weights /= weights.sum(axis=1, keepdims=True)
    elif structure == 'sparse':
        j_reg = _sparse_rows(rs, J, V, 10)
        weights = _sparse_rows(rs, V, J, 4)
    else:
        raise ValueError('unknown structure %r' % (structure,))
    kintree = np.zeros((2, J), dtype=np.int64)
    kintree[0] = np.array(C.SMPL_PARENTS, dtype=np.int64)
    kintree[0, 0] = 2 ** 32 - 1          # the real file stores uint32(-1) for the root
    kintree[1] = np.arange(J)
    faces = rs.randint(0, V, size=(13776, 3)).astype(np.uint32)
    return {
        'v_template': v_template, 'shapedirs': shapedirs, 'posedirs': posedirs,
        'J_regressor': j_reg, 'weights': weights, 'kintree_table': kintree, 'f': faces,
    }


def make_extra_regressor(seed=1, structure='dense'):
    rs = np.random.RandomState(seed)
    if structure == 'sparse':           # <= 10 vertices per extra joint, row sums within a few per cent of 1 (not exactly affine)
        return _sparse_rows(rs, 9, C.NUM_VERTS, 10, row_sum_jitter=0.02).astype(np.float32)
    j = rs.rand(9, C.NUM_VERTS) ** 20
    j /= j.sum(axis=1, keepdims=True)
    return j.astype(np.float32)


def make_gmm(seed=2, num_gaussians=8):
    """{'means','covars','weights'} like gmm_08.pkl; determinants kept close so the
    merged negative-log-likelihood weights stay representable in fp32."""
    rs = np.random.RandomState(seed)
    D = 3 * (C.NUM_SMPL_JOINTS - 1)
    means = 0.2 * rs.randn(num_gaussians, D)
    covars = np.empty((num_gaussians, D, D))
    for k in range(num_gaussians):
        a = 0.05 * rs.randn(D, D)
        covars[k] = a @ a.T + 0.05 * np.eye(D)
    w = rs.rand(num_gaussians)
    w /= w.sum()
    return {'means': means, 'covars': covars, 'weights': w}


def write_data_dir(root, seed=0, structure='dense'):
    """Lay the synthetic files out the way the reference's config.py expects
    (data/smpl/SMPL_NEUTRAL.pkl, data/J_regressor_extra.npy, data/gmm_08.pkl)."""
    os.makedirs(os.path.join(root, 'data', 'smpl'), exist_ok=True)
    with open(os.path.join(root, 'data', 'smpl', 'SMPL_NEUTRAL.pkl'), 'wb') as f:
        pickle.dump(make_smpl_model(seed, structure), f, protocol=2)
    np.save(os.path.join(root, 'data', 'J_regressor_extra.npy'), make_extra_regressor(seed + 1, structure))
    with open(os.path.join(root, 'data', 'gmm_08.pkl'), 'wb') as f:
        pickle.dump(make_gmm(seed + 2), f, protocol=2)
    return root


def make_fit_inputs(batch, seed=0, variant='default'):
    """Per-sample SMPLify inputs as float32 numpy arrays.

    variant 'default': all 49 confidences are 1.
    variant 'trainer': confidences of the ignored joints pre-zeroed (the trainer calls
        get_fitting_loss first, reference train/trainer.py:246 -> smplify.py:156).
    variant 'slp': only the 14 annotated ground-truth slots 25..38 carry confidence.
    """
    rs = np.random.RandomState(1000 + seed)
    pose = (0.2 * rs.randn(batch, 72)).astype(np.float32)
    betas = (0.5 * rs.randn(batch, 10)).astype(np.float32)
    cam_t = (np.array([0., 0., 20.]) + 0.1 * rs.randn(batch, 3)).astype(np.float32)
    center = np.full((batch, 2), 0.5 * C.IMG_RES, dtype=np.float32)
    kp = np.empty((batch, C.NUM_JOINTS_OUT, 3), dtype=np.float32)
    kp[:, :, :2] = rs.uniform(0., C.IMG_RES, size=(batch, C.NUM_JOINTS_OUT, 2))
    kp[:, :, 2] = 1.
    if variant == 'trainer':
        kp[:, C.SMPLIFY_IGNORED_JOINTS, 2] = 0.
    elif variant == 'slp':
        kp[:, :25, 2] = 0.
        kp[:, 39:, 2] = 0.
        kp[:, C.SMPLIFY_IGNORED_JOINTS, 2] = 0.
    elif variant != 'default':
        raise ValueError('unknown variant %r' % (variant,))
    return {'pose': pose, 'betas': betas, 'cam_t': cam_t, 'center': center, 'keypoints': kp}


def model_arrays(seed=0, structure='dense'):
    """The synthetic model as the float32 arrays SMPL(model_arrays=...) takes."""
    m = make_smpl_model(seed, structure)
    arrays = {k: np.asarray(m[k], dtype=np.float32) for k in ('v_template', 'shapedirs', 'posedirs', 'J_regressor', 'weights')}
    parents = np.asarray(m['kintree_table'][0]).astype(np.int64)
    parents[0] = -1
    arrays['parents'] = parents
    arrays['faces'] = np.asarray(m['f']).astype(np.int64)
    return arrays


def build_smplify(device='cuda', num_iters=100, seed=0, step_size=1e-2, focal_length=5000, structure='dense'):
    """SMPLify on the synthetic model, no files needed."""
    from .prior import MaxMixturePrior
    from .smpl import SMPL
    from .smplify import SMPLify
    smpl = SMPL(model_arrays=model_arrays(seed, structure), j_regressor_extra=make_extra_regressor(seed + 1, structure))
    prior = MaxMixturePrior.from_gmm(make_gmm(seed + 2))
    return SMPLify(step_size=step_size, num_iters=num_iters, focal_length=focal_length, device=device,
                   smpl=smpl, pose_prior=prior)
