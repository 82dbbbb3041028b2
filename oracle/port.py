"""CPU oracle: a torch restatement of the reference's SMPLify + SMPL hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product (inbed_pose_estimation_b200/)
imports this file; it is the checker used by tests/, __graft_entry__.smoke() and
the cpu_baseline / --impl reference legs of bench.py.

What it restates (all citations are into /root/reference):
  * smplify/smplify.py:40-136   two-stage fit, torch.optim.Adam, in-place conf zeroing
  * smplify/smplify.py:138-172  get_fitting_loss
  * smplify/losses.py:11-90     gmof, angle_prior, body_fitting_loss, camera_fitting_loss
  * smplify/prior.py:142-160, 181-196   MaxMixturePrior constants + merged_log_likelihood
  * models/smpl.py:14-33        54-joint assembly and joint_map re-indexing
  * utils/geometry.py:9-45, 79-107   quaternion Rodrigues, perspective_projection
  * the third-party `smplx` package (requirements.txt:10, unpinned, 0.1.x API; its
    source is NOT on this box): lbs / batch_rodrigues / batch_rigid_transform /
    vertices2joints / VertexJointSelector, restated from the published algorithm
    (SURVEY.md §8a rows a6-a12).

Pinning: the reference ships no tests or golden vectors, so PARITY IS UNPINNED by
the reference's own repo.  This port is instead pinned against the reference's own
files executed verbatim in the build container (oracle/run_reference.py, which
plugs oracle/smplx_shim.py in place of the absent smplx); the resulting vectors
are committed under tests/golden/ and tests/test_oracle.py replays them.  The
smplx arithmetic itself is restated from recall in both and is the one part no
reference artefact on this box can confirm.

The ops are deliberately the same eager torch ops, in the same order, as the
reference issues them, so that timing this port on host cores is a fair stand-in
for timing the reference's CPU path (bench.py cpu_baseline, kind "port").
"""
import math

import numpy as np
import torch

from inbed_pose_estimation_b200 import constants as C


# ----------------------------------------------------------------------------------------
# smplx restatement
# ----------------------------------------------------------------------------------------
def exp_map_rodrigues(rot_vecs):
    """smplx.lbs.batch_rodrigues: angle = |r + 1e-8|, R = I + sin K + (1-cos) K^2."""
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + 1e-8, dim=1, keepdim=True)
    axis = rot_vecs / angle
    c = torch.cos(angle).unsqueeze(1)
    s = torch.sin(angle).unsqueeze(1)
    rx, ry, rz = torch.split(axis, 1, dim=1)
    zero = torch.zeros((n, 1), dtype=rot_vecs.dtype, device=rot_vecs.device)
    K = torch.cat([zero, -rz, ry, rz, zero, -rx, -ry, rx, zero], dim=1).view(n, 3, 3)
    eye = torch.eye(3, dtype=rot_vecs.dtype, device=rot_vecs.device).unsqueeze(0)
    return eye + s * K + (1 - c) * torch.bmm(K, K)


def regress_joints(regressor, verts):
    """smplx.lbs.vertices2joints."""
    return torch.einsum('bik,ji->bjk', [verts, regressor])


def rigid_chain(rot_mats, joints, parents):
    """smplx.lbs.batch_rigid_transform: world transforms along the kinematic tree.

    `parents` is an int64 tensor; the python loop indexes a list with its elements,
    exactly like upstream (that is where the reference's host syncs come from)."""
    B, NJ = joints.shape[:2]
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] -= joints[:, parents[1:]]
    top = torch.cat([rot_mats.reshape(-1, 3, 3), rel.reshape(-1, 3, 1)], dim=2)
    bottom = torch.zeros((B * NJ, 1, 4), dtype=joints.dtype, device=joints.device)
    bottom[:, 0, 3] = 1
    local = torch.cat([top, bottom], dim=1).view(B, NJ, 4, 4)
    chain = [local[:, 0]]
    for i in range(1, parents.shape[0]):
        chain.append(torch.matmul(chain[parents[i]], local[:, i]))
    world = torch.stack(chain, dim=1)
    posed_joints = world[:, :, :3, 3]
    joints_h = torch.nn.functional.pad(joints, [0, 0, 0, 1])
    rel_world = world - torch.nn.functional.pad(torch.matmul(world, joints_h),
                                                [3, 0, 0, 0, 0, 0, 0, 0])
    return posed_joints, rel_world


def linear_blend_skinning(betas, pose, v_template, shapedirs, posedirs, J_regressor, parents,
                          lbs_weights, pose2rot=True):
    """smplx.lbs.lbs -> (vertices [B,6890,3], posed joints [B,24,3])."""
    B = max(betas.shape[0], pose.shape[0])
    dt, dev = betas.dtype, betas.device
    v_shaped = v_template + torch.einsum('bl,mkl->bmk', [betas, shapedirs])
    J = regress_joints(J_regressor, v_shaped)
    eye = torch.eye(3, dtype=dt, device=dev)
    if pose2rot:
        rot_mats = exp_map_rodrigues(pose.reshape(-1, 3)).view(B, -1, 3, 3)
        feat = (rot_mats[:, 1:, :, :] - eye).view(B, -1)
    else:
        rot_mats = pose.view(B, -1, 3, 3)
        feat = (rot_mats[:, 1:] - eye).reshape(B, -1)
    v_posed = torch.matmul(feat, posedirs).view(B, -1, 3) + v_shaped
    posed_joints, A = rigid_chain(rot_mats, J, parents)
    W = lbs_weights.unsqueeze(0).expand(B, -1, -1)
    T = torch.matmul(W, A.view(B, J_regressor.shape[0], 16)).view(B, -1, 4, 4)
    ones = torch.ones((B, v_posed.shape[1], 1), dtype=dt, device=dev)
    v_h = torch.matmul(T, torch.cat([v_posed, ones], dim=2).unsqueeze(-1))
    return v_h[:, :, :3, 0], posed_joints


class BodyModelOutput(object):
    """Field-compatible stand-in for smplx.body_models.ModelOutput."""
    __slots__ = ('vertices', 'joints', 'full_pose', 'betas', 'global_orient', 'body_pose')

    def __init__(self, vertices=None, joints=None, full_pose=None, betas=None,
                 global_orient=None, body_pose=None, **_unused):
        self.vertices, self.joints, self.full_pose = vertices, joints, full_pose
        self.betas, self.global_orient, self.body_pose = betas, global_orient, body_pose


class OracleSMPL(torch.nn.Module):
    """smplx.SMPL.forward + the reference subclass (models/smpl.py:14-33)."""

    def __init__(self, model, j_regressor_extra, dtype=torch.float32):
        super().__init__()
        f = lambda a: torch.tensor(np.asarray(a, dtype=np.float64), dtype=dtype)
        self.register_buffer('v_template', f(model['v_template']))
        self.register_buffer('shapedirs', f(model['shapedirs']))
        nv = model['posedirs'].shape[0]
        self.register_buffer('posedirs', f(np.reshape(model['posedirs'], [nv * 3, -1]).T))
        self.register_buffer('J_regressor', f(model['J_regressor']))
        self.register_buffer('lbs_weights', f(model['weights']))
        parents = torch.tensor(np.asarray(model['kintree_table'][0]).astype(np.int64))
        parents[0] = -1
        self.register_buffer('parents', parents)
        self.register_buffer('extra_vertex_ids', torch.tensor(C.SMPL_EXTRA_VERTEX_IDS, dtype=torch.long))
        self.register_buffer('J_regressor_extra', torch.tensor(np.asarray(j_regressor_extra), dtype=dtype))
        self.joint_map = torch.tensor([C.JOINT_MAP[n] for n in C.JOINT_NAMES], dtype=torch.long)
        self.faces = np.asarray(model['f'])

    def forward(self, global_orient, body_pose, betas, pose2rot=True, return_full_pose=False, **_kw):
        full_pose = torch.cat([global_orient, body_pose], dim=1)
        verts, chain_joints = linear_blend_skinning(
            betas, full_pose, self.v_template, self.shapedirs, self.posedirs, self.J_regressor,
            self.parents, self.lbs_weights, pose2rot=pose2rot)
        picked = torch.index_select(verts, 1, self.extra_vertex_ids)      # VertexJointSelector
        joints45 = torch.cat([chain_joints, picked], dim=1)
        extra = regress_joints(self.J_regressor_extra, verts)            # models/smpl.py:24
        joints = torch.cat([joints45, extra], dim=1)[:, self.joint_map, :]  # :25-26
        return BodyModelOutput(vertices=verts, joints=joints, betas=betas,
                               global_orient=global_orient, body_pose=body_pose,
                               full_pose=full_pose if return_full_pose else None)


# ----------------------------------------------------------------------------------------
# utils/geometry.py
# ----------------------------------------------------------------------------------------
def quaternion_rodrigues(theta):
    """utils/geometry.py:9-45 (half-angle quaternion, renormalised, 9 quadratic forms)."""
    norm = torch.norm(theta + 1e-8, p=2, dim=1).unsqueeze(-1)
    unit = theta / norm
    half = norm * 0.5
    quat = torch.cat([torch.cos(half), torch.sin(half) * unit], dim=1)
    quat = quat / quat.norm(p=2, dim=1, keepdim=True)
    w, x, y, z = quat[:, 0], quat[:, 1], quat[:, 2], quat[:, 3]
    w2, x2, y2, z2 = w.pow(2), x.pow(2), y.pow(2), z.pow(2)
    wx, wy, wz, xy, xz, yz = w * x, w * y, w * z, x * y, x * z, y * z
    return torch.stack([w2 + x2 - y2 - z2, 2 * xy - 2 * wz, 2 * wy + 2 * xz,
                        2 * wz + 2 * xy, w2 - x2 + y2 - z2, 2 * yz - 2 * wx,
                        2 * xz - 2 * wy, 2 * wx + 2 * yz, w2 - x2 - y2 + z2],
                       dim=1).view(theta.shape[0], 3, 3)


def project_points(points, rotation, translation, focal_length, camera_center, out_3d=False):
    """utils/geometry.py:79-114 (out_3d: the third channel is row 2 of K applied to the un-normalised point = its depth)."""
    B = points.shape[0]
    K = torch.zeros([B, 3, 3], device=points.device, dtype=points.dtype)
    K[:, 0, 0] = focal_length
    K[:, 1, 1] = focal_length
    K[:, 2, 2] = 1.
    K[:, :-1, -1] = camera_center
    points = torch.einsum('bij,bkj->bki', rotation, points)
    points = points + translation.unsqueeze(1)
    proj = points / points[:, :, -1].unsqueeze(-1)
    proj = torch.einsum('bij,bkj->bki', K, proj)
    if not out_3d:
        return proj[:, :, :-1]
    proj[:, :, -1] = torch.einsum('bij,bkj->bki', K, points)[:, :, -1]
    return proj


# ----------------------------------------------------------------------------------------
# smplify/prior.py, smplify/losses.py
# ----------------------------------------------------------------------------------------
class OracleMaxMixturePrior(torch.nn.Module):
    """prior.py:142-160 (constants) and :181-196 (merged negative log-likelihood)."""

    def __init__(self, gmm, dtype=torch.float32):
        super().__init__()
        np_dt = np.float32 if dtype == torch.float32 else np.float64
        means = gmm['means'].astype(np_dt)
        covs = gmm['covars'].astype(np_dt)
        self.register_buffer('means', torch.tensor(means, dtype=dtype))
        prec = np.stack([np.linalg.inv(c) for c in covs]).astype(np_dt)
        self.register_buffer('precisions', torch.tensor(prec, dtype=dtype))
        sqrdets = np.array([np.sqrt(np.linalg.det(c)) for c in gmm['covars']])
        const = (2 * np.pi) ** (69 / 2.)
        nll = np.asarray(gmm['weights'] / (const * (sqrdets / sqrdets.min())))
        self.register_buffer('nll_weights', torch.tensor(nll, dtype=dtype).unsqueeze(0))

    def forward(self, pose, betas=None):
        diff = pose.unsqueeze(1) - self.means
        pd = torch.einsum('mij,bmj->bmi', [self.precisions, diff])
        quad = (pd * diff).sum(dim=-1)
        ll = 0.5 * quad - torch.log(self.nll_weights)
        return torch.min(ll, dim=1)[0]


def gmof(x, sigma):
    """losses.py:11-17."""
    x2 = x ** 2
    s2 = sigma ** 2
    return (s2 * x2) / (s2 + x2)


def angle_prior(pose):
    """losses.py:19-24."""
    sign = torch.tensor(C.ANGLE_PRIOR_SIGNS, device=pose.device, dtype=pose.dtype)
    return torch.exp(pose[:, C.ANGLE_PRIOR_IDS] * sign) ** 2


def body_loss_per_sample(body_pose, betas, model_joints, camera_t, camera_center, joints_2d,
                         joints_conf, pose_prior, focal_length=5000, sigma=100,
                         pose_prior_weight=4.78, shape_prior_weight=5, angle_prior_weight=15.2):
    """losses.py:26-54 -> (total_loss [B], reprojection_loss [B,49])."""
    B = body_pose.shape[0]
    rot = torch.eye(3, device=body_pose.device, dtype=body_pose.dtype).unsqueeze(0).expand(B, -1, -1)
    proj = project_points(model_joints, rot, camera_t, focal_length, camera_center)
    reproj = (joints_conf ** 2) * gmof(proj - joints_2d, sigma).sum(dim=-1)
    prior = (pose_prior_weight ** 2) * pose_prior(body_pose, betas)
    angle = (angle_prior_weight ** 2) * angle_prior(body_pose).sum(dim=-1)
    shape = (shape_prior_weight ** 2) * (betas ** 2).sum(dim=-1)
    return reproj.sum(dim=-1) + prior + angle + shape, reproj


def camera_loss_per_sample(model_joints, camera_t, camera_t_est, camera_center, joints_2d,
                           joints_conf, focal_length=5000, depth_loss_weight=100):
    """losses.py:60-89 -> total_loss [B]."""
    B = model_joints.shape[0]
    rot = torch.eye(3, device=model_joints.device, dtype=model_joints.dtype).unsqueeze(0).expand(B, -1, -1)
    proj = project_points(model_joints, rot, camera_t, focal_length, camera_center)
    op, gt = C.CAMERA_OP_JOINTS, C.CAMERA_GT_JOINTS
    err_op = (joints_2d[:, op] - proj[:, op]) ** 2
    err_gt = (joints_2d[:, gt] - proj[:, gt]) ** 2
    valid = (joints_conf[:, op].min(dim=-1)[0][:, None, None] > 0).to(proj.dtype)
    reproj = (valid * err_op + (1 - valid) * err_gt).sum(dim=(1, 2))
    depth = (depth_loss_weight ** 2) * (camera_t[:, 2] - camera_t_est[:, 2]) ** 2
    return reproj + depth


# ----------------------------------------------------------------------------------------
# smplify/smplify.py
# ----------------------------------------------------------------------------------------
class OracleSMPLify(object):
    """smplify.py:13-172.  `trace`, when a list, receives the per-sample loss [B] of
    every iteration (stage 1 then stage 2) - the reference keeps only the scalar sum."""

    def __init__(self, smpl, pose_prior, step_size=1e-2, num_iters=100, focal_length=5000):
        self.smpl, self.pose_prior = smpl, pose_prior
        self.step_size, self.num_iters, self.focal_length = step_size, num_iters, focal_length
        self.ign_joints = list(C.SMPLIFY_IGNORED_JOINTS)

    def __call__(self, init_pose, init_betas, init_cam_t, camera_center, keypoints_2d, trace=None):
        cam_t = init_cam_t.clone()
        joints_2d = keypoints_2d[:, :, :2]
        joints_conf = keypoints_2d[:, :, -1]
        body_pose = init_pose[:, 3:].detach().clone()
        global_orient = init_pose[:, :3].detach().clone()
        betas = init_betas.detach().clone()

        global_orient.requires_grad = True
        cam_t.requires_grad = True
        opt = torch.optim.Adam([global_orient, cam_t], lr=self.step_size, betas=(0.9, 0.999))
        for _ in range(self.num_iters):
            out = self.smpl(global_orient=global_orient, body_pose=body_pose, betas=betas)
            per = camera_loss_per_sample(out.joints, cam_t, init_cam_t, camera_center, joints_2d,
                                         joints_conf, focal_length=self.focal_length)
            if trace is not None:
                trace.append(per.detach().clone())
            loss = per.sum()
            opt.zero_grad()
            loss.backward()
            opt.step()

        cam_t.requires_grad = False
        body_pose.requires_grad = True
        betas.requires_grad = True
        joints_conf[:, self.ign_joints] = 0.        # in place on the caller's tensor (:105)
        opt = torch.optim.Adam([body_pose, betas, global_orient], lr=self.step_size, betas=(0.9, 0.999))
        for _ in range(self.num_iters):
            out = self.smpl(global_orient=global_orient, body_pose=body_pose, betas=betas)
            per, _ = body_loss_per_sample(body_pose, betas, out.joints, cam_t, camera_center,
                                          joints_2d, joints_conf, self.pose_prior,
                                          focal_length=self.focal_length)
            if trace is not None:
                trace.append(per.detach().clone())
            loss = per.sum()
            opt.zero_grad()
            loss.backward()
            opt.step()

        with torch.no_grad():
            out = self.smpl(global_orient=global_orient, body_pose=body_pose, betas=betas,
                            return_full_pose=True)
            _, reproj = body_loss_per_sample(body_pose, betas, out.joints, cam_t, camera_center,
                                             joints_2d, joints_conf, self.pose_prior,
                                             focal_length=self.focal_length)
        pose = torch.cat([global_orient, body_pose], dim=-1).detach()
        return out.vertices.detach(), out.joints.detach(), pose, betas.detach(), cam_t, reproj

    def get_fitting_loss(self, pose, betas, cam_t, camera_center, keypoints_2d):
        joints_2d = keypoints_2d[:, :, :2]
        joints_conf = keypoints_2d[:, :, -1]
        joints_conf[:, self.ign_joints] = 0.        # in place (:156)
        with torch.no_grad():
            out = self.smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas,
                            return_full_pose=True)
            _, reproj = body_loss_per_sample(pose[:, 3:], betas, out.joints, cam_t, camera_center,
                                             joints_2d, joints_conf, self.pose_prior,
                                             focal_length=self.focal_length)
        return reproj


def build_oracle(seed=0, dtype=torch.float32, num_iters=100, structure='dense'):
    """Oracle SMPLify on the seeded synthetic model (same arrays the product loads)."""
    from inbed_pose_estimation_b200 import synthetic
    smpl = OracleSMPL(synthetic.make_smpl_model(seed, structure), synthetic.make_extra_regressor(seed + 1, structure), dtype=dtype)
    prior = OracleMaxMixturePrior(synthetic.make_gmm(seed + 2), dtype=dtype)
    return OracleSMPLify(smpl, prior, num_iters=num_iters)
