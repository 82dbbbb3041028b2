import os, sys, torch
sys.path.insert(0, '/root/repo')
from inbed_pose_estimation_b200 import _native, synthetic
from inbed_pose_estimation_b200.smpl import SMPL
B = 16384
dev = torch.device('cuda', 0)
lib = _native.lib()
smpl = SMPL(model_arrays=synthetic.model_arrays(0), j_regressor_extra=synthetic.make_extra_regressor(1)).to(dev)
h = smpl.native(dev).handle
st = torch.cuda.current_stream(dev).cuda_stream
pose = (0.2 * torch.randn(B, 72)).to(dev); betas = (0.5 * torch.randn(B, 10)).to(dev)
verts = torch.empty(B, 6890, 3, device=dev); vposed = torch.empty(B, 20736, device=dev); joints = torch.empty(B, 49, 3, device=dev)
dverts = torch.randn(B, 6890, 3, device=dev); djoints = torch.randn(B, 49, 3, device=dev)
dpose, dbetas = torch.empty(B, 72, device=dev), torch.empty(B, 10, device=dev)
ws = torch.empty(lib.smplb200_smpl_workspace_bytes(B), dtype=torch.uint8, device=dev)
P = _native.ptr
_native.check(lib.smplb200_smpl_forward(h, B, 0, P(pose), P(betas), P(verts), P(joints), P(vposed), ws.data_ptr(), ws.numel(), st))
for i in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _native.check(lib.smplb200_smpl_backward(h, B, 0, P(pose), P(betas), P(vposed), P(dverts), P(djoints), P(dpose), P(dbetas), ws.data_ptr(), ws.numel(), st))
    e1.record(); torch.cuda.synchronize()
print('dbg', os.environ.get('SMPLB200_DA_DBG'), 'bwd ms', e0.elapsed_time(e1))
