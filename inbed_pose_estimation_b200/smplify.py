"""SMPLify with the reference's call surface (reference smplify/smplify.py:13-172); the whole
two-stage fit runs as one CUDA kernel launch (+ one vertex kernel) in libsmplify_b200.so.

    smplify = SMPLify(step_size=1e-2, batch_size=32, num_iters=100, focal_length=5000)
    vertices, joints, pose, betas, cam_t, reprojection_loss = smplify(
        init_pose, init_betas, init_cam_t, camera_center, keypoints_2d)

Side effects are the reference's: the confidences keypoints_2d[:, [1,9,12,27,28], 2] are zeroed
in place (smplify.py:105 and :156); all outputs are new, detached tensors.
"""
import torch

from . import _native, config, constants
from .prior import MaxMixturePrior
from .smpl import SMPL


class SMPLify(object):
    """Implementation of single-stage SMPLify (B200)."""

    def __init__(self, step_size=1e-2, batch_size=66, num_iters=100, focal_length=5000,
                 device=torch.device('cuda'), smpl=None, pose_prior=None):
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise RuntimeError('inbed_pose_estimation_b200.SMPLify runs on CUDA (sm_100a) only; there is no CPU fallback')
        self.focal_length = focal_length
        self.step_size = step_size
        ign_joints = ['OP Neck', 'OP RHip', 'OP LHip', 'Right Hip', 'Left Hip']
        self.ign_joints = [constants.JOINT_IDS[i] for i in ign_joints]
        self.num_iters = num_iters
        # GMM pose prior (smplify.py:32-34) and body model (:36-38)
        self.pose_prior = pose_prior if pose_prior is not None else \
            MaxMixturePrior(prior_folder=config.GMM_PRIOR_DIR, num_gaussians=8, dtype=torch.float32)
        self.pose_prior = self.pose_prior.to(self.device)
        self.smpl = smpl if smpl is not None else SMPL(config.SMPL_MODEL_DIR, batch_size=batch_size, create_transl=False)
        self.smpl = self.smpl.to(self.device)
        self.smpl.attach_prior(self.pose_prior.native_constants())
        self._ws = None
        self.last_loss_trace = None

    # ---------------------------------------------------------------------------------------------
    def _workspace(self, batch):
        need = _native.lib().smplb200_fit_workspace_bytes(int(batch))
        if self._ws is None or self._ws.numel() < need or self._ws.device != self._dev:
            self._ws = torch.empty(need, dtype=torch.uint8, device=self._dev)
        return self._ws

    def _prep(self, t, shape, name):
        if t.device.type != 'cuda':
            raise RuntimeError('%s must be a CUDA tensor (no CPU fallback)' % name)
        if tuple(t.shape) != tuple(shape):
            raise ValueError('%s must have shape %s, got %s' % (name, shape, tuple(t.shape)))
        return t.detach().contiguous().float()

    def _keypoints(self, keypoints_2d, B):
        """Contiguous fp32 view of the caller's tensor, or a copy whose confidences are written back."""
        if tuple(keypoints_2d.shape) != (B, constants.NUM_JOINTS_OUT, 3):
            raise ValueError('keypoints_2d must be [B, 49, 3]')
        if keypoints_2d.device.type != 'cuda':
            raise RuntimeError('keypoints_2d must be a CUDA tensor (no CPU fallback)')
        if keypoints_2d.is_contiguous() and keypoints_2d.dtype == torch.float32:
            return keypoints_2d.detach(), None
        return keypoints_2d.detach().contiguous().float(), keypoints_2d

    def __call__(self, init_pose, init_betas, init_cam_t, camera_center, keypoints_2d, return_loss_trace=False, packed_out=None):
        """Perform body fitting.  Returns (vertices [B,6890,3], joints [B,49,3], pose [B,72], betas [B,10],
        camera_translation [B,3], reprojection_loss [B,49]).  packed_out: optional contiguous fp32 CUDA tensor [B,134] that
        additionally receives pose | betas | camera | reprojection of every sample as one row, written by the fit kernel itself
        (the row a sharded refit all-gathers, sharded.PACKED)."""
        B = init_pose.shape[0]
        self._dev = init_pose.device
        pose = self._prep(init_pose, (B, 72), 'init_pose')
        betas = self._prep(init_betas, (B, 10), 'init_betas')
        cam = self._prep(init_cam_t, (B, 3), 'init_cam_t')
        cen = self._prep(camera_center, (B, 2), 'camera_center')
        kp, writeback = self._keypoints(keypoints_2d, B)
        dev = self._dev
        new = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        vertices, joints = new(B, constants.NUM_VERTS, 3), new(B, constants.NUM_JOINTS_OUT, 3)
        o_pose, o_betas, o_cam, reproj = new(B, 72), new(B, 10), new(B, 3), new(B, constants.NUM_JOINTS_OUT)
        trace = new(2 * self.num_iters, B) if return_loss_trace else None
        self.last_loss_trace = trace
        if packed_out is not None and not (packed_out.is_cuda and packed_out.is_contiguous() and packed_out.dtype == torch.float32
                                           and tuple(packed_out.shape) == (B, 134)):
            raise ValueError('packed_out must be a contiguous fp32 CUDA tensor [B, 134]')
        if B == 0:
            return vertices, joints, o_pose, o_betas, o_cam, reproj
        ws = self._workspace(B)
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_smplify_fit(
                self.smpl.native(dev).handle, B, int(self.num_iters), float(self.step_size), float(self.focal_length),
                _native.ptr(pose), _native.ptr(betas), _native.ptr(cam), _native.ptr(cen), _native.ptr(kp),
                _native.ptr(vertices), _native.ptr(joints), _native.ptr(o_pose), _native.ptr(o_betas), _native.ptr(o_cam),
                _native.ptr(reproj), _native.ptr(trace), _native.ptr(packed_out), ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream(dev).cuda_stream))
        if writeback is not None:
            writeback[:, self.ign_joints, 2] = 0.
        self.last_loss_trace = trace
        return vertices, joints, o_pose, o_betas, o_cam, reproj

    def get_fitting_loss(self, pose, betas, cam_t, camera_center, keypoints_2d):
        """Given body and camera parameters, compute the per-joint reprojection loss [B,49]."""
        B = pose.shape[0]
        self._dev = pose.device
        pose_c = self._prep(pose, (B, 72), 'pose')
        betas_c = self._prep(betas, (B, 10), 'betas')
        cam = self._prep(cam_t, (B, 3), 'cam_t')
        cen = self._prep(camera_center, (B, 2), 'camera_center')
        kp, writeback = self._keypoints(keypoints_2d, B)
        reproj = torch.empty((B, constants.NUM_JOINTS_OUT), device=self._dev, dtype=torch.float32)
        if B == 0:
            return reproj
        ws = self._workspace(B)
        with torch.cuda.device(self._dev):
            _native.check(_native.lib().smplb200_smplify_fitting_loss(
                self.smpl.native(self._dev).handle, B, float(self.focal_length), _native.ptr(pose_c), _native.ptr(betas_c),
                _native.ptr(cam), _native.ptr(cen), _native.ptr(kp), _native.ptr(reproj), ws.data_ptr(), ws.numel(),
                torch.cuda.current_stream(self._dev).cuda_stream))
        if writeback is not None:
            writeback[:, self.ign_joints, 2] = 0.
        return reproj

    def prior_terms(self, pose, betas):
        """The three prior terms of body_fitting_loss (reference smplify/losses.py:46-52) evaluated on their own by the same
        device code the fit runs every iteration.  pose [B,72], betas [B,10] -> dict with 'terms' [B,3] (weighted pose prior,
        angle prior, shape prior), 'components' [B,8], 'argmin' [B] (int32), 'grad_body_pose' [B,69], 'grad_betas' [B,10].
        Not part of the reference's surface: a diagnostic / test hook."""
        import ctypes
        B = pose.shape[0]
        self._dev = pose.device
        pose_c = self._prep(pose, (B, 72), 'pose')
        betas_c = self._prep(betas, (B, 10), 'betas')
        dev = self._dev
        out = {'terms': torch.empty((B, 3), device=dev), 'components': torch.empty((B, 8), device=dev),
               'argmin': torch.empty((B,), device=dev, dtype=torch.int32), 'grad_body_pose': torch.empty((B, 69), device=dev),
               'grad_betas': torch.empty((B, 10), device=dev)}
        if B:
            with torch.cuda.device(dev):
                _native.check(_native.lib().smplb200_prior_terms(
                    self.smpl.native(dev).handle, B, _native.ptr(pose_c), _native.ptr(betas_c), _native.ptr(out['terms']),
                    _native.ptr(out['components']), ctypes.c_void_p(out['argmin'].data_ptr()), _native.ptr(out['grad_body_pose']),
                    _native.ptr(out['grad_betas']), torch.cuda.current_stream(dev).cuda_stream))
        return out
