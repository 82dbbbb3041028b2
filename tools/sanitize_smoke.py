"""Small end-to-end invocations of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
Few iterations and small batches: the tools slow kernels down by 10-100x."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from inbed_pose_estimation_b200 import geometry, synthetic, train_losses as TL  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 2
fitter = synthetic.build_smplify('cuda', num_iters=iters, seed=0)
for B in (5, 8 * 3 + 1, 12 * 148 - 3 if len(sys.argv) > 2 else 40, 16 * 148 + 30 if len(sys.argv) > 2 else 70):
    inp = synthetic.make_fit_inputs(B, seed=B)
    args = [torch.from_numpy(inp[k]).cuda() for k in ('pose', 'betas', 'cam_t', 'center', 'keypoints')]
    out = fitter(*args)
    p = args[0].clone().requires_grad_(True)
    b = args[1].clone().requires_grad_(True)
    o = fitter.smpl(global_orient=p[:, :3], body_pose=p[:, 3:], betas=b)
    (o.vertices.square().sum() + o.joints.square().sum()).backward()
    R = geometry.batch_rodrigues(args[0].reshape(-1, 3)).view(B, 24, 3, 3).clone().requires_grad_(True)
    o = fitter.smpl(global_orient=R[:, :1], body_pose=R[:, 1:], betas=b.detach(), pose2rot=False)
    o.vertices.sum().backward()
    kp2 = geometry.perspective_projection(o.joints.detach(), torch.eye(3, device='cuda').expand(B, 3, 3), args[2], 5000., args[3])
    mask = (torch.arange(B, device='cuda') % 3 != 0).to(torch.uint8)
    TL.keypoint_loss(kp2 / 112., args[4], 0., 1.)
    TL.shape_loss(o.vertices.detach(), out[0], mask)
    TL.smpl_losses(R.detach(), b.detach(), out[2], out[3], mask)
    TL.keypoint_3d_loss(o.joints.detach(), torch.rand(B, 24, 4, device='cuda'), mask)
    geometry.estimate_translation(o.joints.detach(), args[4])
    geometry.rotation_matrix_to_angle_axis(R.detach().view(-1, 3, 3), scrub_nan=True)
    torch.cuda.synchronize()
    print('B = %d ok' % B)
