"""TEST INFRASTRUCTURE ONLY - CPU restatement of the steps either side of SMPLify.

Follows, in order: utils/geometry.py:47-61 (rot6d_to_rotmat), train/trainer.py:702-706 (rotmat -> axis-angle through
torchgeometry + NaN scrub), utils/geometry.py:118-181 (estimate_translation), train/fits_dict.py:34-94 (FitsDict get /
set with flip_pose :62-70 and rotate_pose :72-94, which calls cv2.Rodrigues per sample), train/trainer.py:716-727
(keep-if-better).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline may import this module.
Pinned against tests/golden/adjacent.npz, which oracle/run_reference_adjacent.py produced by running the reference's
own utils/geometry.py and train/fits_dict.py (torchgeometry through oracle/tgm_shim.py [recall], OpenCV real).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import tgm_shim

POSE_FLIP_PERM = [3 * j + c for j in [0, 2, 1, 3, 5, 4, 6, 8, 7, 9, 11, 10, 12, 14, 13, 15, 17, 16, 19, 18, 21, 20, 23, 22]
                  for c in range(3)]


def rot6d_to_rotmat(x):
    x = x.view(-1, 3, 2)
    a1, a2 = x[:, :, 0], x[:, :, 1]
    b1 = F.normalize(a1)
    b2 = F.normalize(a2 - torch.einsum('bi,bi->b', b1, a2).unsqueeze(-1) * b1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-1)


def rotmat_to_axis_angle(rotmat, scrub_nan=True):
    """trainer.py:702-706 on [N,3,3] matrices."""
    n = rotmat.shape[0]
    hom = torch.cat([rotmat.view(-1, 3, 3), torch.tensor([0, 0, 1], dtype=torch.float32).view(1, 3, 1).expand(n, -1, -1)], dim=-1)
    aa = tgm_shim.rotation_matrix_to_angle_axis(hom).contiguous()
    if scrub_nan:
        aa[torch.isnan(aa)] = 0.0
    return aa


def estimate_translation(S, joints_2d, focal_length=5000., img_size=224.):
    """float64 weighted least squares per sample over joints 25..48."""
    S = S[:, 25:, :].double().numpy()
    conf = joints_2d[:, 25:, 2].double().numpy()
    uv = joints_2d[:, 25:, :2].double().numpy()
    out = np.zeros((S.shape[0], 3), dtype=np.float32)
    c = img_size / 2.
    for i in range(S.shape[0]):
        w = np.repeat(np.sqrt(conf[i]), 2)
        Q = np.zeros((2 * S.shape[1], 3))
        Q[0::2, 0] = focal_length
        Q[1::2, 1] = focal_length
        Q[:, 2] = c - uv[i].reshape(-1)
        rhs = (uv[i].reshape(-1) - c) * np.repeat(S[i, :, 2], 2) - focal_length * S[i, :, :2].reshape(-1)
        Q, rhs = Q * w[:, None], rhs * w
        out[i] = np.linalg.solve(Q.T @ Q, Q.T @ rhs)
    return torch.from_numpy(out)


def flip_pose(pose, is_flipped):
    is_flipped = is_flipped.bool()
    out = pose.clone()
    out[is_flipped, :] = pose[is_flipped][:, POSE_FLIP_PERM]
    out[is_flipped, 1::3] *= -1
    out[is_flipped, 2::3] *= -1
    return out


def rotate_pose(pose, rot):
    import cv2
    pose = pose.clone()
    rot = rot.float()
    cos, sin = torch.cos(-np.pi * rot / 180.), torch.sin(-np.pi * rot / 180.)
    zeros = torch.zeros_like(cos)
    r3 = torch.zeros(cos.shape[0], 1, 3)
    r3[:, 0, -1] = 1
    R = torch.cat([torch.stack([cos, -sin, zeros], dim=-1).unsqueeze(1), torch.stack([sin, cos, zeros], dim=-1).unsqueeze(1), r3], dim=1)
    g = tgm_shim.angle_axis_to_rotation_matrix(pose[:, :3])[:, :3, :3]
    g = torch.matmul(R, g).numpy()
    aa = np.zeros((pose.shape[0], 3))
    for i in range(pose.shape[0]):
        v, _ = cv2.Rodrigues(g[i])
        aa[i] = v.squeeze()
    pose[:, :3] = torch.from_numpy(aa).to(pose.dtype)
    return pose


def fits_get(store, index, rot, is_flipped):
    params = store[index]
    return flip_pose(rotate_pose(params[:, :72].clone(), rot), is_flipped), params[:, 72:].clone()


def fits_set(store, index, rot, is_flipped, update, pose, betas):
    pose = rotate_pose(flip_pose(pose, is_flipped), -rot)
    params = torch.cat((pose, betas), dim=-1)
    for n, i in enumerate(index.tolist()):
        if bool(update[n]):
            store[i] = params[n]
    return store


def keep_better(best_loss, best_pose, best_betas, best_cam, new_reproj, new_pose, new_betas, new_cam):
    new_loss = new_reproj.mean(dim=-1)
    update = new_loss < best_loss
    best_loss, best_pose, best_betas, best_cam = best_loss.clone(), best_pose.clone(), best_betas.clone(), best_cam.clone()
    best_loss[update] = new_loss[update]
    best_pose[update, :] = new_pose[update, :]
    best_betas[update, :] = new_betas[update, :]
    best_cam[update, :] = new_cam[update, :]
    return best_loss, best_pose, best_betas, best_cam, update


# ---------------------------------------------------------------------------------------------------------------------
# What the train step does with the SMPLify result: restatement of train/trainer.py:88-117, 158-178
# and 735-748, pinned by tests/golden/train_losses.npz (oracle/run_reference_trainer.py runs the reference's own source).
# ---------------------------------------------------------------------------------------------------------------------
def golden_vertex_pair(seed, n):
    """The two [n,6890,3] vertex tensors of the train_losses golden cases (too large to commit; legacy RandomState
    streams are frozen across numpy versions).  The first 100 vertices of row 0 coincide: exact zeros at the L1 kink."""
    rs = np.random.RandomState(int(seed))
    a = (rs.randn(n, 6890, 3) * 0.3).astype(np.float32)
    b = (rs.randn(n, 6890, 3) * 0.3).astype(np.float32)
    b[0, :100] = a[0, :100]
    return a, b


def keypoint_loss(pred_keypoints_2d, gt_keypoints_2d, openpose_weight, gt_weight):
    """trainer.py:88-98."""
    conf = gt_keypoints_2d[:, :, -1].unsqueeze(-1).clone()
    conf[:, :25] *= openpose_weight
    conf[:, 25:] *= gt_weight
    return (conf * (pred_keypoints_2d - gt_keypoints_2d[:, :, :-1]) ** 2).mean()


def keypoint_3d_loss(pred_keypoints_3d, gt_keypoints_3d, has_pose_3d):
    """trainer.py:100-117 (0-dim zero instead of the reference's shape-[1] zero when no row has 3D labels)."""
    sel = has_pose_3d != 0
    pred = pred_keypoints_3d[:, 25:, :][sel]
    conf = gt_keypoints_3d[:, :, -1].unsqueeze(-1)[sel]
    gt = gt_keypoints_3d[:, :, :-1][sel]
    if len(gt) == 0:
        return pred_keypoints_3d.new_zeros(())
    gt = gt - ((gt[:, 2, :] + gt[:, 3, :]) / 2)[:, None, :]
    pred = pred - ((pred[:, 2, :] + pred[:, 3, :]) / 2)[:, None, :]
    return (conf * (pred - gt) ** 2).mean()


def shape_loss(pred_vertices, gt_vertices, has_smpl):
    """trainer.py:158-164."""
    sel = has_smpl != 0
    if int(sel.sum()) == 0:
        return pred_vertices.new_zeros(())
    return (pred_vertices[sel] - gt_vertices[sel]).abs().mean()


def quat_batch_rodrigues(theta):
    """utils/geometry.py:9-45 (batch_rodrigues through quat_to_rotmat)."""
    l1norm = torch.norm(theta + 1e-8, p=2, dim=1)
    angle = l1norm.unsqueeze(-1)
    normalized = theta / angle
    angle = angle * 0.5
    quat = torch.cat([torch.cos(angle), torch.sin(angle) * normalized], dim=1)
    quat = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    w2, x2, y2, z2 = w * w, x * x, y * y, z * z
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2], dim=1).view(-1, 3, 3)


def smpl_losses(pred_rotmat, pred_betas, gt_pose, gt_betas, has_smpl):
    """trainer.py:165-178."""
    sel = has_smpl != 0
    if int(sel.sum()) == 0:
        z = pred_rotmat.new_zeros(())
        return z, z.clone()
    gt_rotmat = quat_batch_rodrigues(gt_pose.reshape(-1, 3)).view(-1, 24, 3, 3)
    return ((pred_rotmat[sel] - gt_rotmat[sel]) ** 2).mean(), ((pred_betas[sel] - gt_betas[sel]) ** 2).mean()


def finalize_fits(opt_pose, opt_betas, opt_cam_t, opt_joints, opt_vertices, opt_joint_loss, has_smpl,
                  gt_pose, gt_betas, gt_cam_t, gt_model_joints, gt_vertices, smplify_threshold=100.):
    """trainer.py:735-748 on copies; returns (pose, betas, cam_t, joints, vertices, valid_fit)."""
    pose, betas, cam, joints = opt_pose.clone(), opt_betas.clone(), opt_cam_t.clone(), opt_joints.clone()
    verts = opt_vertices.clone() if opt_vertices is not None else None
    has = has_smpl != 0
    betas[(betas.abs() > 3).any(dim=-1)] = 0.
    if verts is not None:
        verts[has] = gt_vertices[has]
    cam[has] = gt_cam_t[has]
    joints[has] = gt_model_joints[has]
    pose[has] = gt_pose[has]
    betas[has] = gt_betas[has]
    valid_fit = (opt_joint_loss < smplify_threshold) | has
    return pose, betas, cam, joints, verts, valid_fit


def weak_perspective_projection(joints, pred_camera, focal_length=5000., img_res=224):
    """trainer.py:187-199: (keypoints_2d normalised to [-1, 1], cam_t)."""
    cam_t = torch.stack([pred_camera[:, 1], pred_camera[:, 2], 2 * focal_length / (img_res * pred_camera[:, 0] + 1e-9)], dim=-1)
    p = joints + cam_t.unsqueeze(1)
    proj = p / p[:, :, -1].unsqueeze(-1)
    kp = focal_length * proj[:, :, :2]
    return kp / (img_res / 2.), cam_t
