"""What the reference's train step does with the SMPLify result, CUDA-backed.

Same names and argument meaning as the `Trainer` methods they replace (train/trainer.py):

    keypoint_loss(pred_keypoints_2d, gt_keypoints_2d, openpose_weight, gt_weight)     trainer.py:88-98
    keypoint_3d_loss(pred_keypoints_3d, gt_keypoints_3d, has_pose_3d)                 trainer.py:100-117
    shape_loss(pred_vertices, gt_vertices, has_smpl)                                  trainer.py:158-164
    smpl_losses(pred_rotmat, pred_betas, gt_pose, gt_betas, has_smpl)                 trainer.py:165-178
    finalize_fits_(...)                                                               trainer.py:735-748

Each loss is one pass over its inputs that also produces d(loss)/d(prediction); the number of selected rows is counted on
the device, so - unlike the reference's boolean-mask indexing - nothing synchronises with the host.  Consequence: the
losses are always 0-dim tensors; with an empty mask the reference returns a zero tensor of shape [1], here it is a 0-dim
zero.  Gradients flow to the predictions only (the reference's targets are detached data).  CUDA only, no fallback.
"""
import ctypes

import torch

from . import _native


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


def _require_cuda(t, name):
    if t.device.type != 'cuda':
        raise RuntimeError('%s runs on CUDA (sm_100a) only; got a tensor on %s - there is no CPU fallback' % (name, t.device))


def _f32(t, dev):
    return t.detach().to(dev).float().contiguous()


def _mask(t, dev, n):
    m = (t.detach().to(dev) != 0).to(torch.uint8).contiguous()
    if m.shape != (n,):
        raise ValueError('mask must have shape [%d], got %s' % (n, tuple(m.shape)))
    return m


def _u8p(t):
    return ctypes.c_void_p(t.data_ptr())


def _workspace(batch, dev):
    n = int(_native.lib().smplb200_train_loss_workspace_bytes(int(batch)))
    return torch.empty(n // 8, dtype=torch.float64, device=dev)


class _LossWithGrad(torch.autograd.Function):
    """forward(run, n_losses, *predictions): `run(grads_wanted)` launches the fused loss + gradient kernels and returns
    (losses [n_losses], tuple of gradient tensors - one per prediction); backward scales the saved gradients."""

    @staticmethod
    def forward(ctx, run, n_losses, *preds):
        losses, grads = run()
        ctx.n_preds = len(preds)
        ctx.set_materialize_grads(False)            # an unused loss leaves its prediction without a gradient, as in eager torch
        ctx.save_for_backward(*grads)
        return tuple(losses[i] for i in range(n_losses))

    @staticmethod
    def backward(ctx, *gouts):
        grads = ctx.saved_tensors
        # loss i owns gradient i (smpl_losses: pose loss -> rotmat, betas loss -> betas; the others have one of each)
        out = [None, None] + [None] * ctx.n_preds
        for i, g in enumerate(grads):
            if i < len(gouts) and gouts[i] is not None:
                out[2 + i] = g * gouts[i]
        return tuple(out)


def keypoint_loss(pred_keypoints_2d, gt_keypoints_2d, openpose_weight, gt_weight):
    """(conf * MSE(pred, gt[..., :2])).mean(), conf = gt[..., 2] scaled per slot group (trainer.py:88-98)."""
    _require_cuda(pred_keypoints_2d, 'keypoint_loss')
    dev = pred_keypoints_2d.device
    B = pred_keypoints_2d.shape[0]
    if tuple(pred_keypoints_2d.shape[1:]) != (49, 2) or tuple(gt_keypoints_2d.shape) != (B, 49, 3):
        raise ValueError('expected pred [B,49,2] and gt [B,49,3]')
    p, g = _f32(pred_keypoints_2d, dev), _f32(gt_keypoints_2d, dev)

    def run():
        loss = torch.empty(1, device=dev)
        grad = torch.empty_like(p)
        ws = _workspace(B, dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_keypoint_loss(B, _native.ptr(p), _native.ptr(g), float(openpose_weight),
                                                               float(gt_weight), _native.ptr(loss), _native.ptr(grad),
                                                               ctypes.c_void_p(ws.data_ptr()), _stream(dev)))
        return loss, (grad,)
    return _LossWithGrad.apply(run, 1, pred_keypoints_2d)[0]


def keypoint_3d_loss(pred_keypoints_3d, gt_keypoints_3d, has_pose_3d):
    """Pelvis-centred, confidence-weighted 3D keypoint loss on the rows with 3D labels (trainer.py:100-117).
    pred_keypoints_3d is the full [B,49,3] joint set (the reference slices [:, 25:] itself)."""
    _require_cuda(pred_keypoints_3d, 'keypoint_3d_loss')
    dev = pred_keypoints_3d.device
    B = pred_keypoints_3d.shape[0]
    if tuple(pred_keypoints_3d.shape[1:]) != (49, 3) or tuple(gt_keypoints_3d.shape) != (B, 24, 4):
        raise ValueError('expected pred [B,49,3] and gt [B,24,4]')
    p, g, m = _f32(pred_keypoints_3d, dev), _f32(gt_keypoints_3d, dev), _mask(has_pose_3d, dev, B)

    def run():
        loss = torch.empty(1, device=dev)
        grad = torch.empty_like(p)
        ws = _workspace(B, dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_keypoint_3d_loss(B, _native.ptr(p), _native.ptr(g), _u8p(m), _native.ptr(loss),
                                                                  _native.ptr(grad), ctypes.c_void_p(ws.data_ptr()), _stream(dev)))
        return loss, (grad,)
    return _LossWithGrad.apply(run, 1, pred_keypoints_3d)[0]


def shape_loss(pred_vertices, gt_vertices, has_smpl):
    """nn.L1Loss between predicted and target vertices on the rows with has_smpl (trainer.py:158-164)."""
    _require_cuda(pred_vertices, 'shape_loss')
    dev = pred_vertices.device
    B = pred_vertices.shape[0]
    if tuple(pred_vertices.shape[1:]) != (6890, 3) or tuple(gt_vertices.shape) != (B, 6890, 3):
        raise ValueError('expected vertices [B,6890,3]')
    p, g, m = _f32(pred_vertices, dev), _f32(gt_vertices, dev), _mask(has_smpl, dev, B)
    want_grad = pred_vertices.requires_grad and torch.is_grad_enabled()

    def run():
        loss = torch.empty(1, device=dev)
        grad = torch.empty_like(p) if want_grad else None          # 83 KB / sample: only written when someone will read it
        ws = _workspace(B, dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_shape_loss(B, _native.ptr(p), _native.ptr(g), _u8p(m), _native.ptr(loss),
                                                            _native.ptr(grad), ctypes.c_void_p(ws.data_ptr()), _stream(dev)))
        return loss, ((grad,) if want_grad else ())
    return _LossWithGrad.apply(run, 1, pred_vertices)[0]


def smpl_losses(pred_rotmat, pred_betas, gt_pose, gt_betas, has_smpl):
    """(MSE(pred_rotmat, batch_rodrigues(gt_pose)), MSE(pred_betas, gt_betas)) on the rows with has_smpl (trainer.py:165-178)."""
    _require_cuda(pred_rotmat, 'smpl_losses')
    dev = pred_rotmat.device
    B = pred_rotmat.shape[0]
    if tuple(pred_rotmat.shape[1:]) != (24, 3, 3) or tuple(pred_betas.shape) != (B, 10) or gt_pose.numel() != B * 72:
        raise ValueError('expected pred_rotmat [B,24,3,3], pred_betas [B,10], gt_pose [B,72]')
    r, b = _f32(pred_rotmat, dev), _f32(pred_betas, dev)
    gp, gb, m = _f32(gt_pose, dev).view(B, 72), _f32(gt_betas, dev), _mask(has_smpl, dev, B)

    def run():
        losses = torch.empty(2, device=dev)
        gr, gbt = torch.empty_like(r), torch.empty_like(b)
        ws = _workspace(B, dev)
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_smpl_param_losses(B, _native.ptr(r), _native.ptr(b), _native.ptr(gp), _native.ptr(gb),
                                                                   _u8p(m), _native.ptr(losses), _native.ptr(gr), _native.ptr(gbt),
                                                                   ctypes.c_void_p(ws.data_ptr()), _stream(dev)))
        return losses, (gr, gbt)
    return _LossWithGrad.apply(run, 2, pred_rotmat, pred_betas)


def finalize_fits_(opt_pose, opt_betas, opt_cam_t, opt_joints, opt_vertices, opt_joint_loss, has_smpl,
                   gt_pose, gt_betas, gt_cam_t, gt_model_joints, gt_vertices, smplify_threshold=100.):
    """In place, one kernel (trainer.py:735-748): zero the betas rows with any |beta| > 3, overwrite the rows with
    has_smpl by the ground truth, and return valid_fit = (opt_joint_loss < smplify_threshold) | has_smpl (bool [B]).
    opt_vertices / gt_vertices may both be None."""
    _require_cuda(opt_pose, 'finalize_fits_')
    dev = opt_pose.device
    B = opt_pose.shape[0]
    opts = [opt_pose, opt_betas, opt_cam_t, opt_joints] + ([opt_vertices] if opt_vertices is not None else [])
    for t in opts:
        if not (t.is_contiguous() and t.dtype == torch.float32 and t.device == dev):
            raise ValueError('opt_* tensors must be contiguous fp32 CUDA tensors (they are updated in place)')
    if (opt_vertices is None) != (gt_vertices is None):
        raise ValueError('opt_vertices and gt_vertices must be given together')
    m = _mask(has_smpl, dev, B)
    valid = torch.zeros(B, dtype=torch.uint8, device=dev)
    gv = _f32(gt_vertices, dev) if gt_vertices is not None else None
    if B:
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_finalize_fits(
                B, float(smplify_threshold), _u8p(m), _native.ptr(_f32(gt_pose, dev)), _native.ptr(_f32(gt_betas, dev)),
                _native.ptr(_f32(gt_cam_t, dev)), _native.ptr(_f32(gt_model_joints, dev)), _native.ptr(gv),
                _native.ptr(_f32(opt_joint_loss, dev)), _native.ptr(opt_pose), _native.ptr(opt_betas), _native.ptr(opt_cam_t),
                _native.ptr(opt_joints), _native.ptr(opt_vertices), _u8p(valid), _stream(dev)))
    return valid.bool()
