"""Generate tests/golden/*.npz by executing the REFERENCE'S OWN FILES verbatim.

TEST INFRASTRUCTURE ONLY; runs only where /root/reference exists (the build
container).  The GPU box never runs this - it replays the committed vectors.

Recipe: put /root/reference on sys.path; pre-register empty
namespace packages for utils / models / smplify / train / datasets so their heavy
__init__.py files (matplotlib, torchgeometry, ...) are never executed; install the
smplx shim; chdir into a temp dir holding the synthetic data files at the paths
config.py names; then import smplify.smplify, smplify.losses, smplify.prior,
models.smpl, utils.geometry, constants - all of them the reference's files.

    python -m oracle.run_reference            # writes tests/golden/
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

REF = os.environ.get('INBED_REFERENCE_ROOT', '/root/reference')
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def import_reference(data_root):
    """Returns a namespace with the reference modules; cwd is changed to data_root."""
    from . import smplx_shim
    if not os.path.isdir(REF):
        raise RuntimeError('reference tree %s not present' % REF)
    smplx_shim.install()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for pkg in ('utils', 'models', 'smplify', 'train', 'datasets'):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REF, pkg)]
            sys.modules[pkg] = m
    os.chdir(data_root)
    import importlib
    ns = types.SimpleNamespace()
    ns.constants = importlib.import_module('constants')
    ns.config = importlib.import_module('config')
    ns.geometry = importlib.import_module('utils.geometry')
    ns.smpl = importlib.import_module('models.smpl')
    ns.losses = importlib.import_module('smplify.losses')
    ns.prior = importlib.import_module('smplify.prior')
    ns.smplify = importlib.import_module('smplify.smplify')
    for mod in (ns.constants, ns.geometry, ns.smpl, ns.losses, ns.prior, ns.smplify):
        assert os.path.abspath(mod.__file__).startswith(REF), mod.__file__
    return ns


def tapped_fit(ns, fitter, inputs):
    """Run the reference SMPLify.__call__ while recording the scalar loss of each iteration."""
    trace = []
    mod = ns.smplify
    cam_fn, body_fn = mod.camera_fitting_loss, mod.body_fitting_loss

    def cam_tap(*a, **k):
        v = cam_fn(*a, **k)
        trace.append(float(v.detach()))
        return v

    def body_tap(*a, **k):
        v = body_fn(*a, **k)
        if k.get('output', 'sum') == 'sum':
            trace.append(float(v.detach()))
        return v

    mod.camera_fitting_loss, mod.body_fitting_loss = cam_tap, body_tap
    try:
        out = fitter(*[torch.from_numpy(inputs[k].copy()) for k in
                       ('pose', 'betas', 'cam_t', 'center', 'keypoints')])
    finally:
        mod.camera_fitting_loss, mod.body_fitting_loss = cam_fn, body_fn
    return out, np.asarray(trace, dtype=np.float64)


def main():
    from inbed_pose_estimation_b200 import synthetic
    os.makedirs(GOLDEN, exist_ok=True)
    tmp = tempfile.mkdtemp(prefix='inbed_ref_')
    synthetic.write_data_dir(tmp, seed=0)
    ns = import_reference(tmp)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cpu = torch.device('cpu')

    # ---- integer tables ---------------------------------------------------------------
    smpl = ns.smpl.SMPL(ns.config.SMPL_MODEL_DIR, batch_size=4, create_transl=False)
    ign = [ns.constants.JOINT_IDS[i] for i in ['OP Neck', 'OP RHip', 'OP LHip', 'Right Hip', 'Left Hip']]
    np.savez(os.path.join(GOLDEN, 'tables.npz'),
             joint_map=smpl.joint_map.numpy(), ign_joints=np.asarray(ign),
             pose_flip_perm=np.asarray(ns.constants.SMPL_POSE_FLIP_PERM),
             j49_flip_perm=np.asarray(ns.constants.J49_FLIP_PERM),
             joint_names=np.asarray(ns.constants.JOINT_NAMES))

    # ---- known-answer tests of SURVEY.md §4 -------------------------------------------
    kat = {}
    kat['gmof_100_100'] = ns.losses.gmof(torch.tensor(100.), 100).numpy()
    kat['angle_prior_zero'] = (ns.losses.angle_prior(torch.zeros(1, 69)).sum() * 15.2 ** 2).numpy()
    kat['proj_simple'] = ns.geometry.perspective_projection(
        torch.tensor([[[0., 0., 0.], [1., 2., 0.]]]), torch.eye(3)[None], torch.tensor([[0., 0., 10.]]),
        5000., torch.tensor([[112., 112.]])).numpy()
    kat['rodrigues_zero'] = ns.geometry.batch_rodrigues(torch.zeros(2, 3)).numpy()
    np.savez(os.path.join(GOLDEN, 'kat.npz'), **kat)

    # ---- geometry on random inputs (forward + autograd gradients) -------------------------
    rs = np.random.RandomState(7)
    theta = torch.tensor(rs.randn(64, 3).astype(np.float32), requires_grad=True)
    Rq = ns.geometry.batch_rodrigues(theta)
    gR = torch.tensor(rs.randn(64, 3, 3).astype(np.float32))
    (Rq * gR).sum().backward()
    pts = torch.tensor(rs.randn(8, 49, 3).astype(np.float32), requires_grad=True)
    rot = ns.geometry.batch_rodrigues(torch.tensor(0.3 * rs.randn(8, 3).astype(np.float32))).detach()
    rot.requires_grad_(True)
    tr = torch.tensor((np.array([0, 0, 20.]) + 0.1 * rs.randn(8, 3)).astype(np.float32), requires_grad=True)
    cen = torch.tensor((112 + rs.randn(8, 2)).astype(np.float32))
    pr = ns.geometry.perspective_projection(pts, rot, tr, 5000., cen)
    gP = torch.tensor(rs.randn(8, 49, 2).astype(np.float32))
    (pr * gP).sum().backward()
    np.savez(os.path.join(GOLDEN, 'geometry.npz'),
             theta=theta.detach().numpy(), rotmat=Rq.detach().numpy(), grad_rotmat=gR.numpy(),
             grad_theta=theta.grad.numpy(),
             points=pts.detach().numpy(), rotation=rot.detach().numpy(), translation=tr.detach().numpy(),
             center=cen.numpy(), projected=pr.detach().numpy(), grad_projected=gP.numpy(),
             grad_points=pts.grad.numpy(), grad_rotation=rot.grad.numpy(), grad_translation=tr.grad.numpy())

    # ---- SMPL forward, both pose modes, with gradients --------------------------------------
    inp = synthetic.make_fit_inputs(4, seed=0)
    pose = torch.tensor(inp['pose'], requires_grad=True)
    betas = torch.tensor(inp['betas'], requires_grad=True)
    out = smpl(global_orient=pose[:, :3], body_pose=pose[:, 3:], betas=betas)
    gv = torch.tensor(rs.randn(4, 6890, 3).astype(np.float32))
    gj = torch.tensor(rs.randn(4, 49, 3).astype(np.float32))
    ((out.vertices * gv).sum() + (out.joints * gj).sum()).backward()
    rotm = ns.smpl.smplx.lbs.batch_rodrigues(pose.detach().reshape(-1, 3)).view(4, 24, 3, 3).clone()
    rotm.requires_grad_(True)
    betas2 = betas.detach().clone().requires_grad_(True)
    out2 = smpl(global_orient=rotm[:, :1], body_pose=rotm[:, 1:], betas=betas2, pose2rot=False)
    ((out2.vertices * gv).sum() + (out2.joints * gj).sum()).backward()
    sub = slice(None, None, 8)
    np.savez_compressed(
        os.path.join(GOLDEN, 'smpl_forward.npz'),
        pose=inp['pose'], betas=inp['betas'], vertex_stride=8,
        vertices_sub=out.vertices.detach().numpy()[:, sub], joints=out.joints.detach().numpy(),
        vertices_checksum=out.vertices.detach().double().sum(dim=1).numpy(),
        grad_seed=7, grad_pose=pose.grad.numpy(), grad_betas=betas.grad.numpy(),
        rotmats=rotm.detach().numpy(), vertices_rotmat_sub=out2.vertices.detach().numpy()[:, sub],
        joints_rotmat=out2.joints.detach().numpy(),
        grad_rotmats=rotm.grad.numpy(), grad_betas_rotmat=betas2.grad.numpy())

    # ---- prior ---------------------------------------------------------------------------
    prior = ns.prior.MaxMixturePrior(prior_folder='data', num_gaussians=8, dtype=torch.float32)
    bp = torch.tensor((0.25 * rs.randn(16, 69)).astype(np.float32), requires_grad=True)
    nll = prior(bp, None)
    nll.sum().backward()
    np.savez(os.path.join(GOLDEN, 'prior.npz'), body_pose=bp.detach().numpy(), nll=nll.detach().numpy(),
             grad=bp.grad.numpy(), nll_weights=prior.nll_weights.numpy(),
             precisions_checksum=prior.precisions.double().sum(dim=(1, 2)).numpy())

    # ---- the fit itself: three confidence variants, B=4, 100+100 iterations ---------------
    fitter = ns.smplify.SMPLify(step_size=1e-2, batch_size=4, num_iters=100, focal_length=5000,
                                device=cpu)
    for variant in ('default', 'trainer', 'slp'):
        inp = synthetic.make_fit_inputs(4, seed=3, variant=variant)
        kp_t = torch.from_numpy(inp['keypoints'].copy())
        (verts, joints, pose_o, betas_o, cam_o, reproj), trace = tapped_fit(ns, fitter, inp)
        kp_for_loss = torch.from_numpy(inp['keypoints'].copy())
        floss = fitter.get_fitting_loss(torch.from_numpy(inp['pose']), torch.from_numpy(inp['betas']),
                                        torch.from_numpy(inp['cam_t']), torch.from_numpy(inp['center']),
                                        kp_for_loss)
        np.savez_compressed(
            os.path.join(GOLDEN, 'smplify_%s.npz' % variant),
            num_iters=100, seed=3, variant=variant,
            pose=inp['pose'], betas=inp['betas'], cam_t=inp['cam_t'], center=inp['center'],
            keypoints=inp['keypoints'],
            out_vertices_sub=verts.numpy()[:, sub], out_joints=joints.numpy(), out_pose=pose_o.numpy(),
            out_betas=betas_o.numpy(), out_cam_t=cam_o.detach().numpy(), out_reproj=reproj.numpy(),
            loss_trace=trace, init_fitting_loss=floss.numpy(),
            keypoints_after_loss=kp_for_loss.numpy())
        print('%-8s final mean reprojection %.4f  first/last loss %.4f / %.4f' %
              (variant, float(reproj.mean()), trace[0], trace[-1]))
    print('golden vectors written to', GOLDEN)


if __name__ == '__main__':
    main()
