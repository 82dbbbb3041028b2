"""SMPL body model with the reference's call surface (reference models/smpl.py:11-33, on top of
smplx.SMPL), computed by the CUDA library.

    smpl = SMPL(config.SMPL_MODEL_DIR, batch_size=32, create_transl=False).to('cuda')
    out = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)          # axis-angle
    out = smpl(global_orient=R[:, :1], body_pose=R[:, 1:], betas=betas, pose2rot=False)  # rotation matrices
    out.vertices [B,6890,3], out.joints [B,49,3]   - differentiable w.r.t. all three inputs
"""
import os
import pickle

import numpy as np
import torch

from . import _native, config, constants


class ModelOutput(object):
    """Field-compatible with smplx.body_models.ModelOutput as the reference uses it
    (models/smpl.py:27-32: vertices, global_orient, body_pose, joints, betas, full_pose)."""
    __slots__ = ('vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose',
                 'expression', 'left_hand_pose', 'right_hand_pose', 'jaw_pose')

    def __init__(self, vertices=None, joints=None, full_pose=None, betas=None, global_orient=None,
                 body_pose=None, expression=None, left_hand_pose=None, right_hand_pose=None, jaw_pose=None):
        self.vertices, self.joints, self.full_pose, self.betas = vertices, joints, full_pose, betas
        self.global_orient, self.body_pose = global_orient, body_pose
        self.expression, self.left_hand_pose, self.right_hand_pose, self.jaw_pose = \
            expression, left_hand_pose, right_hand_pose, jaw_pose


def _dense(a):
    """numpy array from the things SMPL pkl files hold (ndarray, chumpy array, scipy sparse)."""
    if hasattr(a, 'toarray'):
        a = a.toarray()
    elif hasattr(a, 'r'):
        a = a.r
    return np.asarray(a)


def load_model_arrays(model_path, gender='neutral'):
    if os.path.isdir(model_path):
        model_path = os.path.join(model_path, 'SMPL_{}.pkl'.format(gender.upper()))
    assert os.path.exists(model_path), 'Path {} does not exist!'.format(model_path)
    with open(model_path, 'rb') as f:
        data = pickle.load(f, encoding='latin1')
    arrays = {k: _dense(data[k]).astype(np.float32) for k in ('v_template', 'posedirs', 'J_regressor', 'weights')}
    arrays['shapedirs'] = _dense(data['shapedirs'])[:, :, :constants.NUM_BETAS].astype(np.float32)
    parents = _dense(data['kintree_table'])[0].astype(np.int64)
    parents[0] = -1
    arrays['parents'] = parents
    arrays['faces'] = _dense(data['f']).astype(np.int64)
    return arrays


class _SMPLFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, pose, betas, rotmat_mode, need_vertices):
        B = betas.shape[0]
        dev = betas.device
        if dev.type != 'cuda':
            module.native(dev)                       # raises: CUDA only, no CPU fallback
        lib = _native.lib()
        pose_c = pose.detach().contiguous().float()
        betas_c = betas.detach().contiguous().float()
        verts = torch.empty((B, constants.NUM_VERTS, 3), device=dev, dtype=torch.float32) if need_vertices else None
        joints = torch.empty((B, constants.NUM_JOINTS_OUT, 3), device=dev, dtype=torch.float32)
        needs_grad = bool(ctx.needs_input_grad[1] or ctx.needs_input_grad[2])   # forward runs under no_grad
        # v_posed of the forward pass, kept for the backward call: [B][20736] (padded, 16-byte aligned rows - TMA target)
        vposed = (torch.empty((B, _native.VPOSED_PITCH), device=dev, dtype=torch.float32)
                  if (needs_grad and need_vertices) else None)
        with torch.cuda.device(dev):
            native = module.native(dev)              # first use creates the constant blob: on `dev`, not the current device
            ws = module.workspace(dev, B)
            if B > 0:
                _native.check(lib.smplb200_smpl_forward(
                    native.handle, B, int(rotmat_mode), _native.ptr(pose_c), _native.ptr(betas_c), _native.ptr(verts),
                    _native.ptr(joints), _native.ptr(vposed), ws.data_ptr(), ws.numel(),
                    torch.cuda.current_stream(dev).cuda_stream))
        ctx.module, ctx.rotmat_mode, ctx.pose_shape = module, rotmat_mode, pose.shape
        ctx.save_for_backward(pose_c, betas_c, vposed if vposed is not None else torch.empty(0, device=dev))
        ctx.has_vposed = vposed is not None
        ctx.wanted_vertices = bool(need_vertices)
        ctx.set_materialize_grads(False)          # a joints-only loss must not run the vertex backward on a zero gradient
        if verts is None:
            verts = torch.empty(0, device=dev)
            ctx.mark_non_differentiable(verts)
        return verts, joints

    @staticmethod
    def backward(ctx, g_verts, g_joints):
        pose_c, betas_c, vposed = ctx.saved_tensors
        module = ctx.module
        B = betas_c.shape[0]
        dev = betas_c.device
        lib = _native.lib()
        gv = g_verts.contiguous().float() if (g_verts is not None and ctx.has_vposed) else None
        if g_verts is not None and g_verts.numel() > 0 and not ctx.has_vposed and ctx.wanted_vertices:
            raise RuntimeError('SMPL backward: vertex gradient arrived but v_posed was not saved')
        gj = g_joints.contiguous().float() if g_joints is not None else None
        if gv is None and gj is None:              # neither output was used by the loss
            return None, torch.zeros(ctx.pose_shape, device=dev), torch.zeros_like(betas_c), None, None
        d_pose = torch.empty_like(pose_c)
        d_betas = torch.empty_like(betas_c)
        if B == 0:
            return None, d_pose.view(ctx.pose_shape), d_betas, None, None
        with torch.cuda.device(dev):
            ws = module.workspace(dev, B)
            _native.check(lib.smplb200_smpl_backward(
                module.native(dev).handle, B, int(ctx.rotmat_mode), _native.ptr(pose_c), _native.ptr(betas_c),
                _native.ptr(vposed) if ctx.has_vposed else None, _native.ptr(gv), _native.ptr(gj),
                _native.ptr(d_pose), _native.ptr(d_betas), ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream(dev).cuda_stream))
        return None, d_pose.view(ctx.pose_shape), d_betas, None, None


class SMPL(torch.nn.Module):
    """Drop-in for the reference's ``models.smpl.SMPL`` (constructor: models/smpl.py:14-19 and
    smplify/smplify.py:36-38; forward: models/smpl.py:21-33)."""

    def __init__(self, model_path=config.SMPL_MODEL_DIR, batch_size=1, create_transl=False, gender='neutral',
                 model_arrays=None, j_regressor_extra=None, **kwargs):
        super(SMPL, self).__init__()
        arrays = dict(model_arrays) if model_arrays is not None else load_model_arrays(model_path, gender)
        if j_regressor_extra is None:
            j_regressor_extra = np.load(config.JOINT_REGRESSOR_TRAIN_EXTRA)
        arrays['J_regressor_extra'] = np.asarray(j_regressor_extra, dtype=np.float32)
        self.batch_size = batch_size
        self.faces = arrays.pop('faces', None)
        self._arrays = arrays
        self.register_buffer('J_regressor_extra', torch.tensor(arrays['J_regressor_extra'], dtype=torch.float32))
        self.register_buffer('parents', torch.tensor(arrays['parents'], dtype=torch.long))
        self.joint_map = torch.tensor([constants.JOINT_MAP[i] for i in constants.JOINT_NAMES], dtype=torch.long)
        self._prior = None
        self._native = {}
        self._ws = {}

    # -- native state -----------------------------------------------------------------------------
    def attach_prior(self, prior_constants):
        """Give the constant blob the GMM prior SMPLify needs (drops blobs built without it)."""
        self._prior = prior_constants
        for n in self._native.values():
            n.close()
        self._native = {}

    def native(self, device):
        device = torch.device(device)
        if device.type != 'cuda':
            raise RuntimeError('inbed_pose_estimation_b200.SMPL runs on CUDA (sm_100a) only; got tensors on %s - '
                               'there is no CPU fallback' % device)
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if idx not in self._native:
            self._native[idx] = _native.NativeModel(self._arrays, self._prior, idx)
        return self._native[idx]

    def workspace(self, device, batch):
        idx = torch.device(device).index
        idx = torch.cuda.current_device() if idx is None else idx
        need = _native.lib().smplb200_smpl_workspace_bytes(int(batch))
        ws = self._ws.get(idx)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=torch.device('cuda', idx))
            self._ws[idx] = ws
        return ws

    # -- forward ----------------------------------------------------------------------------------
    def forward(self, global_orient=None, body_pose=None, betas=None, pose2rot=True, return_full_pose=False,
                return_verts=True, **kwargs):
        if global_orient is None or body_pose is None or betas is None:
            raise ValueError('global_orient, body_pose and betas are required')
        B = betas.shape[0]
        if pose2rot:
            full_pose = torch.cat([global_orient.reshape(B, 3), body_pose.reshape(B, 69)], dim=1)
        else:
            full_pose = torch.cat([global_orient.reshape(B, 1, 3, 3), body_pose.reshape(B, 23, 3, 3)], dim=1)
        verts, joints = _SMPLFunction.apply(self, full_pose, betas, not pose2rot, bool(return_verts))
        return ModelOutput(vertices=verts if return_verts else None, joints=joints, betas=betas,
                           global_orient=global_orient, body_pose=body_pose,
                           full_pose=full_pose if return_full_pose else None)
