"""Geometry ops of the hot path with the reference's signatures (utils/geometry.py), CUDA-backed.

batch_rodrigues        - utils/geometry.py:9-23 (+ quat_to_rotmat :25-45)
perspective_projection - utils/geometry.py:79-114 (both the 2-D and the out_3d=True form)
Both are differentiable (hand-written backward kernels) and CUDA only.
"""
import torch

from . import _native


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _require_cuda(t, name):
    if t.device.type != 'cuda':
        raise RuntimeError('%s runs on CUDA (sm_100a) only; got a tensor on %s - there is no CPU fallback' % (name, t.device))


class _Rodrigues(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta):
        th = theta.detach().contiguous().float()
        out = torch.empty((th.shape[0], 3, 3), device=th.device, dtype=torch.float32)
        with torch.cuda.device(th.device):
            _native.check(_native.lib().smplb200_batch_rodrigues(th.shape[0], _native.ptr(th), _native.ptr(out), _stream(th.device)))
        ctx.save_for_backward(th)
        return out

    @staticmethod
    def backward(ctx, g):
        (th,) = ctx.saved_tensors
        g = g.contiguous().float()
        d = torch.empty_like(th)
        with torch.cuda.device(th.device):
            _native.check(_native.lib().smplb200_batch_rodrigues_backward(th.shape[0], _native.ptr(th), _native.ptr(g), _native.ptr(d),
                                                                          _stream(th.device)))
        return d


def batch_rodrigues(theta):
    """Axis-angle [N,3] -> rotation matrices [N,3,3] through the normalised half-angle quaternion."""
    _require_cuda(theta, 'batch_rodrigues')
    if theta.dim() != 2 or theta.shape[1] != 3:
        raise ValueError('theta must be [N, 3]')
    return _Rodrigues.apply(theta)


class _Projection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, rotation, translation, focal, center, out_3d):
        B, N = points.shape[0], points.shape[1]
        dev = points.device
        p = points.detach().contiguous().float()
        r = rotation.detach().expand(B, 3, 3).contiguous().float()
        t = translation.detach().contiguous().float()
        c = center.detach().contiguous().float()
        per_batch = int(focal.numel() > 1)
        out = torch.empty((B, N, 3 if out_3d else 2), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_perspective_projection(
                B, N, _native.ptr(p), _native.ptr(r), _native.ptr(t), _native.ptr(focal), per_batch, _native.ptr(c), int(out_3d),
                _native.ptr(out), _stream(dev)))
        ctx.save_for_backward(p, r, t, focal)
        ctx.per_batch, ctx.out_3d, ctx.rot_shape = per_batch, bool(out_3d), tuple(rotation.shape)
        return out

    @staticmethod
    def backward(ctx, g):
        p, r, t, focal = ctx.saved_tensors
        B, N = p.shape[0], p.shape[1]
        g = g.contiguous().float()
        gp, gr, gt = torch.empty_like(p), torch.empty_like(r), torch.empty_like(t)
        with torch.cuda.device(p.device):
            _native.check(_native.lib().smplb200_perspective_projection_backward(
                B, N, _native.ptr(p), _native.ptr(r), _native.ptr(t), _native.ptr(focal), ctx.per_batch, int(ctx.out_3d),
                _native.ptr(g), _native.ptr(gp), _native.ptr(gr), _native.ptr(gt), _stream(p.device)))
        if ctx.rot_shape != tuple(gr.shape):               # a broadcast rotation ([3,3] or [1,3,3]) gets the summed gradient
            gr = gr.sum_to_size(ctx.rot_shape)
        return gp, gr, gt, None, None, None


def perspective_projection(points, rotation, translation, focal_length, camera_center, out_3d=False):
    """points [B,N,3], rotation [B,3,3] (or broadcastable to it), translation [B,3], focal_length scalar or [B],
    camera_center [B,2] -> [B,N,2]; with out_3d=True (reference utils/geometry.py:108-114; train/trainer.py:621-626,
    models/hmr.py:1720) -> [B,N,3] whose third channel is the camera-space depth.  Differentiable w.r.t. points, rotation
    and translation; the reference's callers pass constant intrinsics, so focal_length / camera_center that require grad
    are rejected instead of silently getting none."""
    _require_cuda(points, 'perspective_projection')
    if points.dim() != 3 or points.shape[2] != 3:
        raise ValueError('points must be [B, N, 3]')
    B = points.shape[0]
    if tuple(translation.shape) != (B, 3) or tuple(camera_center.shape) != (B, 2):
        raise ValueError('translation must be [B, 3] and camera_center [B, 2]')
    if rotation.dim() not in (2, 3) or tuple(rotation.shape[-2:]) != (3, 3) or (rotation.dim() == 3 and rotation.shape[0] not in (1, B)):
        raise ValueError('rotation must be [B, 3, 3], [1, 3, 3] or [3, 3]')
    if torch.is_tensor(focal_length):
        if focal_length.requires_grad:
            raise NotImplementedError('perspective_projection: no gradient w.r.t. focal_length (constant intrinsics only)')
        focal = focal_length.detach().to(points.device, torch.float32).reshape(-1).contiguous()
        if focal.numel() not in (1, B):
            raise ValueError('focal_length must be a scalar or have one entry per batch element')
    else:
        focal = torch.full((1,), float(focal_length), device=points.device, dtype=torch.float32)
    if camera_center.requires_grad:
        raise NotImplementedError('perspective_projection: no gradient w.r.t. camera_center (constant intrinsics only)')
    return _Projection.apply(points, rotation, translation, focal, camera_center, bool(out_3d))


# ---- the steps either side of SMPLify -----------------------------------------------------------------
class _WeakPerspective(torch.autograd.Function):
    @staticmethod
    def forward(ctx, joints, pred_camera, focal, img_res):
        dev = joints.device
        B, N = joints.shape[0], joints.shape[1]
        j = joints.detach().contiguous().float()
        c = pred_camera.detach().contiguous().float()
        cam_t = torch.empty((B, 3), device=dev, dtype=torch.float32)
        kp = torch.empty((B, N, 2), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_weak_perspective_projection(
                B, N, _native.ptr(j), _native.ptr(c), float(focal), float(img_res), _native.ptr(cam_t), _native.ptr(kp), _stream(dev)))
        ctx.save_for_backward(j, c)
        ctx.consts = (float(focal), float(img_res))
        ctx.set_materialize_grads(False)
        return kp, cam_t

    @staticmethod
    def backward(ctx, g_kp, g_cam_t):
        j, c = ctx.saved_tensors
        B, N = j.shape[0], j.shape[1]
        g_kp = torch.zeros((B, N, 2), device=j.device) if g_kp is None else g_kp.contiguous().float()
        g_t = None if g_cam_t is None else g_cam_t.contiguous().float()
        gj, gc = torch.empty_like(j), torch.empty_like(c)
        with torch.cuda.device(j.device):
            _native.check(_native.lib().smplb200_weak_perspective_projection_backward(
                B, N, _native.ptr(j), _native.ptr(c), ctx.consts[0], ctx.consts[1], _native.ptr(g_kp), _native.ptr(g_t),
                _native.ptr(gj), _native.ptr(gc), _stream(j.device)))
        return gj, gc, None, None


def weak_perspective_projection(joints, pred_camera, focal_length=5000., img_res=224):
    """trainer.py:187-199 in one kernel: pred_camera [B,3] = (s, tx, ty) -> cam_t = (tx, ty, 2 f / (img_res s + 1e-9)); joints
    [B,N,3] projected with it (identity rotation, zero centre) and normalised by img_res / 2.  Returns (keypoints_2d [B,N,2],
    cam_t [B,3]); differentiable w.r.t. joints and pred_camera (cam_t.detach() is SMPLify's init_cam_t, trainer.py:713)."""
    _require_cuda(joints, 'weak_perspective_projection')
    if joints.dim() != 3 or joints.shape[2] != 3 or tuple(pred_camera.shape) != (joints.shape[0], 3):
        raise ValueError('expected joints [B, N, 3] and pred_camera [B, 3]')
    return _WeakPerspective.apply(joints, pred_camera, focal_length, img_res)


def rot6d_to_rotmat(x):
    """6-D rotation representation -> rotation matrices (reference utils/geometry.py:47-61).
    x: (B, 6) or anything viewable as (-1, 3, 2); returns (N, 3, 3).  Forward only: in the reference this op sits
    inside the CNN regressors (out of scope); the SMPLify path consumes its detached output."""
    _require_cuda(x, 'rot6d_to_rotmat')
    x6 = x.detach().reshape(-1, 6).contiguous().float()
    out = torch.empty((x6.shape[0], 3, 3), device=x6.device, dtype=torch.float32)
    with torch.cuda.device(x6.device):
        _native.check(_native.lib().smplb200_rot6d_to_rotmat(x6.shape[0], _native.ptr(x6), _native.ptr(out), _stream(x6.device)))
    return out


def rotmat_to_rot6d(matrix):
    """Drop the last row (reference utils/geometry.py:64-77)."""
    return matrix[..., :2, :].clone().reshape(*matrix.size()[:-2], 6)


def rotation_matrix_to_angle_axis(rotation_matrix, scrub_nan=False):
    """torchgeometry.rotation_matrix_to_angle_axis as the trainer uses it (train/trainer.py:702-706):
    (N, 3, 3) or (N, 3, 4) rotation matrices -> (N, 3) axis-angle; scrub_nan=True also applies the trainer's
    `pred_pose[torch.isnan(pred_pose)] = 0.0` patch in the same kernel."""
    _require_cuda(rotation_matrix, 'rotation_matrix_to_angle_axis')
    if rotation_matrix.dim() != 3 or rotation_matrix.shape[1] != 3 or rotation_matrix.shape[2] not in (3, 4):
        raise ValueError('rotation_matrix must be (N, 3, 3) or (N, 3, 4)')
    R = rotation_matrix.detach()[:, :, :3].contiguous().float()
    out = torch.empty((R.shape[0], 3), device=R.device, dtype=torch.float32)
    with torch.cuda.device(R.device):
        _native.check(_native.lib().smplb200_rotmat_to_axis_angle(R.shape[0], _native.ptr(R), _native.ptr(out), int(bool(scrub_nan)),
                                                                  _stream(R.device)))
    return out


def estimate_translation(S, joints_2d, focal_length=5000., img_size=224.):
    """Camera translation that best reprojects the 24 ground-truth joints (reference utils/geometry.py:156-181).
    S: (B, 49, 3) 3-D joints, joints_2d: (B, 49, 3) (x, y, confidence) -> (B, 3).  The reference copies to the host
    and solves one 3x3 system per sample with numpy; here one thread per sample does the same float64 arithmetic."""
    _require_cuda(S, 'estimate_translation')
    B = S.shape[0]
    if tuple(S.shape) != (B, 49, 3) or tuple(joints_2d.shape) != (B, 49, 3):
        raise ValueError('S and joints_2d must be (B, 49, 3)')
    Sc, kp = S.detach().contiguous().float(), joints_2d.detach().to(S.device).contiguous().float()
    out = torch.empty((B, 3), device=S.device, dtype=torch.float32)
    with torch.cuda.device(S.device):
        _native.check(_native.lib().smplb200_estimate_translation(B, _native.ptr(Sc), _native.ptr(kp), float(focal_length), float(img_size),
                                                                  _native.ptr(out), _stream(S.device)))
    return out
