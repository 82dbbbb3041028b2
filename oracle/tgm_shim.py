"""TEST INFRASTRUCTURE ONLY - restatement of the two torchgeometry functions the reference calls.

torchgeometry (reference requirements.txt:13, unpinned; 0.1.x API: `rotation_matrix_to_angle_axis` at
train/trainer.py:4,704 and eval.py:306, `angle_axis_to_rotation_matrix` at train/fits_dict.py:5,84) is a
third-party package that is neither in /root/reference nor installed here.  Its published 0.1.2 algorithms are
restated below from memory [recall]; parity of everything that goes through them is therefore "unpinned" by any
artefact on this box (same status as the smplx shim).  `install()` registers the module as `torchgeometry` so that
the reference's own train/fits_dict.py can be imported to generate golden vectors.
"""
import sys
import types

import torch


def rotation_matrix_to_quaternion(rotation_matrix, eps=1e-6):
    """(N, 3, 4) or (N, 3, 3) -> (N, 4) quaternion (w, x, y, z); four-branch form on the transposed matrix."""
    rmat_t = torch.transpose(rotation_matrix[:, :, :3], 1, 2)
    mask_d2 = rmat_t[:, 2, 2] < eps
    mask_d0_d1 = rmat_t[:, 0, 0] > rmat_t[:, 1, 1]
    mask_d0_nd1 = rmat_t[:, 0, 0] < -rmat_t[:, 1, 1]

    t0 = 1 + rmat_t[:, 0, 0] - rmat_t[:, 1, 1] - rmat_t[:, 2, 2]
    q0 = torch.stack([rmat_t[:, 1, 2] - rmat_t[:, 2, 1], t0, rmat_t[:, 0, 1] + rmat_t[:, 1, 0],
                      rmat_t[:, 2, 0] + rmat_t[:, 0, 2]], -1)
    t1 = 1 - rmat_t[:, 0, 0] + rmat_t[:, 1, 1] - rmat_t[:, 2, 2]
    q1 = torch.stack([rmat_t[:, 2, 0] - rmat_t[:, 0, 2], rmat_t[:, 0, 1] + rmat_t[:, 1, 0], t1,
                      rmat_t[:, 1, 2] + rmat_t[:, 2, 1]], -1)
    t2 = 1 - rmat_t[:, 0, 0] - rmat_t[:, 1, 1] + rmat_t[:, 2, 2]
    q2 = torch.stack([rmat_t[:, 0, 1] - rmat_t[:, 1, 0], rmat_t[:, 2, 0] + rmat_t[:, 0, 2],
                      rmat_t[:, 1, 2] + rmat_t[:, 2, 1], t2], -1)
    t3 = 1 + rmat_t[:, 0, 0] + rmat_t[:, 1, 1] + rmat_t[:, 2, 2]
    q3 = torch.stack([t3, rmat_t[:, 1, 2] - rmat_t[:, 2, 1], rmat_t[:, 2, 0] - rmat_t[:, 0, 2],
                      rmat_t[:, 0, 1] - rmat_t[:, 1, 0]], -1)

    c0 = (mask_d2 & mask_d0_d1).view(-1, 1).type_as(q0)
    c1 = (mask_d2 & ~mask_d0_d1).view(-1, 1).type_as(q0)
    c2 = (~mask_d2 & mask_d0_nd1).view(-1, 1).type_as(q0)
    c3 = (~mask_d2 & ~mask_d0_nd1).view(-1, 1).type_as(q0)
    q = q0 * c0 + q1 * c1 + q2 * c2 + q3 * c3
    q = q / torch.sqrt(t0.view(-1, 1) * c0 + t1.view(-1, 1) * c1 + t2.view(-1, 1) * c2 + t3.view(-1, 1) * c3)
    q = q * 0.5
    return q


def quaternion_to_angle_axis(quaternion):
    q1, q2, q3 = quaternion[..., 1], quaternion[..., 2], quaternion[..., 3]
    sin_squared_theta = q1 * q1 + q2 * q2 + q3 * q3
    sin_theta = torch.sqrt(sin_squared_theta)
    cos_theta = quaternion[..., 0]
    two_theta = 2.0 * torch.where(cos_theta < 0.0, torch.atan2(-sin_theta, -cos_theta), torch.atan2(sin_theta, cos_theta))
    k_pos = two_theta / sin_theta
    k_neg = 2.0 * torch.ones_like(sin_theta)
    k = torch.where(sin_squared_theta > 0.0, k_pos, k_neg)
    angle_axis = torch.zeros_like(quaternion)[..., :3]
    angle_axis[..., 0] += q1 * k
    angle_axis[..., 1] += q2 * k
    angle_axis[..., 2] += q3 * k
    return angle_axis


def rotation_matrix_to_angle_axis(rotation_matrix):
    return quaternion_to_angle_axis(rotation_matrix_to_quaternion(rotation_matrix))


def angle_axis_to_rotation_matrix(angle_axis):
    """(N, 3) -> (N, 4, 4) homogeneous rotation; Rodrigues with axis = aa / (theta + 1e-6) where theta^2 > 1e-6, else the
    first-order Taylor form."""
    def _normal(aa, theta2, eps=1e-6):
        theta = torch.sqrt(theta2)
        wxyz = aa / (theta + eps)
        wx, wy, wz = torch.chunk(wxyz, 3, dim=1)
        c, s = torch.cos(theta), torch.sin(theta)
        k = 1.0 - c
        r00 = c + wx * wx * k
        r10 = wz * s + wx * wy * k
        r20 = -wy * s + wx * wz * k
        r01 = wx * wy * k - wz * s
        r11 = c + wy * wy * k
        r21 = wx * s + wy * wz * k
        r02 = wy * s + wx * wz * k
        r12 = -wx * s + wy * wz * k
        r22 = c + wz * wz * k
        return torch.cat([r00, r01, r02, r10, r11, r12, r20, r21, r22], dim=1).view(-1, 3, 3)

    def _taylor(aa):
        rx, ry, rz = torch.chunk(aa, 3, dim=1)
        one = torch.ones_like(rx)
        return torch.cat([one, -rz, ry, rz, one, -rx, -ry, rx, one], dim=1).view(-1, 3, 3)

    _aa = torch.unsqueeze(angle_axis, dim=1)
    theta2 = torch.squeeze(torch.matmul(_aa, _aa.transpose(1, 2)), dim=1)
    mask = (theta2 > 1e-6).view(-1, 1, 1)
    mask_pos, mask_neg = mask.type_as(theta2), (~mask).type_as(theta2)
    out = torch.eye(4).to(angle_axis.device).type_as(angle_axis).view(1, 4, 4).repeat(angle_axis.shape[0], 1, 1)
    out[..., :3, :3] = mask_pos * _normal(angle_axis, theta2) + mask_neg * _taylor(angle_axis)
    return out


def install():
    if 'torchgeometry' in sys.modules and not getattr(sys.modules['torchgeometry'], '_inbed_shim', False):
        raise RuntimeError('a real torchgeometry is installed; the shim must not shadow it')
    m = types.ModuleType('torchgeometry')
    m._inbed_shim = True
    m.rotation_matrix_to_angle_axis = rotation_matrix_to_angle_axis
    m.angle_axis_to_rotation_matrix = angle_axis_to_rotation_matrix
    m.rotation_matrix_to_quaternion = rotation_matrix_to_quaternion
    m.quaternion_to_angle_axis = quaternion_to_angle_axis
    sys.modules['torchgeometry'] = m
    return m
