"""A stand-in `smplx` package so the reference's own files import in this container.

TEST INFRASTRUCTURE ONLY (used by oracle/run_reference.py).  The reference depends on
the un-vendored pip package smplx (requirements.txt:10; `ModelOutput` pins it to the
0.1.x API, upstream SPIN uses smplx==0.1.13) whose source is absent from this box.
The shim exposes exactly the three names models/smpl.py:3-6 imports - smplx.SMPL,
smplx.lbs.vertices2joints, smplx.body_models.ModelOutput - backed by the restated
arithmetic in oracle/port.py.
"""
import os
import pickle
import sys
import types

import numpy as np
import torch

from . import port
from inbed_pose_estimation_b200 import constants as C


class ShimSMPL(torch.nn.Module):
    """Constructor and forward signature of smplx.SMPL as the reference uses them
    (models/smpl.py:15,23; smplify.py:36-38: SMPL(model_dir, batch_size=, create_transl=False))."""

    def __init__(self, model_path, batch_size=1, create_transl=False, gender='neutral',
                 dtype=torch.float32, **_kw):
        super().__init__()
        if os.path.isdir(model_path):
            model_path = os.path.join(model_path, 'SMPL_%s.pkl' % gender.upper())
        with open(model_path, 'rb') as f:
            data = pickle.load(f, encoding='latin1')
        f64 = lambda a: np.asarray(a, dtype=np.float64)
        self.batch_size, self.dtype = batch_size, dtype
        self.faces = np.asarray(data['f'])
        self.register_buffer('v_template', torch.tensor(f64(data['v_template']), dtype=dtype))
        self.register_buffer('shapedirs', torch.tensor(f64(data['shapedirs']), dtype=dtype))
        nv = data['posedirs'].shape[0]
        self.register_buffer('posedirs',
                             torch.tensor(np.reshape(f64(data['posedirs']), [nv * 3, -1]).T, dtype=dtype))
        self.register_buffer('J_regressor', torch.tensor(f64(data['J_regressor']), dtype=dtype))
        self.register_buffer('lbs_weights', torch.tensor(f64(data['weights']), dtype=dtype))
        parents = torch.tensor(np.asarray(data['kintree_table'][0]).astype(np.int64)).long()
        parents[0] = -1
        self.register_buffer('parents', parents)
        self.register_buffer('extra_joints_idxs', torch.tensor(C.SMPL_EXTRA_VERTEX_IDS, dtype=torch.long))

    def forward(self, betas=None, body_pose=None, global_orient=None, transl=None,
                return_verts=True, return_full_pose=False, pose2rot=True, **_kw):
        full_pose = torch.cat([global_orient, body_pose], dim=1)
        verts, joints = port.linear_blend_skinning(
            betas, full_pose, self.v_template, self.shapedirs, self.posedirs, self.J_regressor,
            self.parents, self.lbs_weights, pose2rot=pose2rot)
        picked = torch.index_select(verts, 1, self.extra_joints_idxs)
        joints = torch.cat([joints, picked], dim=1)
        return ModelOutput(vertices=verts if return_verts else None, joints=joints, betas=betas,
                           global_orient=global_orient, body_pose=body_pose,
                           full_pose=full_pose if return_full_pose else None)


class ModelOutput(port.BodyModelOutput):
    pass


def install():
    """Register the fake package in sys.modules (idempotent)."""
    if 'smplx' in sys.modules and getattr(sys.modules['smplx'], '_is_oracle_shim', False):
        return sys.modules['smplx']
    pkg = types.ModuleType('smplx')
    pkg._is_oracle_shim = True
    pkg.SMPL = ShimSMPL
    lbs = types.ModuleType('smplx.lbs')
    lbs.vertices2joints = port.regress_joints
    lbs.lbs = port.linear_blend_skinning
    lbs.batch_rodrigues = port.exp_map_rodrigues
    lbs.batch_rigid_transform = port.rigid_chain
    bm = types.ModuleType('smplx.body_models')
    bm.ModelOutput = ModelOutput
    bm.SMPL = ShimSMPL
    pkg.lbs, pkg.body_models = lbs, bm
    sys.modules['smplx'] = pkg
    sys.modules['smplx.lbs'] = lbs
    sys.modules['smplx.body_models'] = bm
    return pkg
