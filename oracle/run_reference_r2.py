"""Round-2 golden vectors, produced like oracle/run_reference.py by executing the REFERENCE'S OWN FILES verbatim.

TEST INFRASTRUCTURE ONLY; runs only where /root/reference exists (the build container).

    python -m oracle.run_reference_r2            # writes tests/golden/{geometry_out3d,prior_near_tie,smpl_forward_sparse,smplify_sparse_*}.npz

  geometry_out3d.npz       utils/geometry.py:79-114 perspective_projection(out_3d=True), forward and autograd gradients, at the
                           trainer's shape (6890 points per sample, train/trainer.py:621-626) and at 49 joints
  prior_near_tie.npz       smplify/prior.py:181-196 on body poses placed where two mixture components are within a few 1e-6
                           (relative) of each other, plus losses.py:19-24 angle prior and the shape prior, with gradients
  smpl_forward_sparse.npz  models/smpl.py:21-33 on a SPARSE-structured synthetic model (<= 10 vertices per regressor row,
  smplify_sparse_*.npz     <= 4 skinning weights per vertex, exact zeros, extra-regressor rows that do not sum to 1) and the
                           full smplify/smplify.py:40-136 fit on it
"""
import os
import tempfile

import numpy as np
import torch

from .run_reference import GOLDEN, import_reference, tapped_fit


def near_tie_poses(prior, rs, count, rel_gap):
    """Body poses [count, 69] (fp32) at which the two best mixture components of the reference prior differ by about
    rel_gap (relative): a bisection in float64 along the segment between two random poses with different winners."""
    means, prec = prior.means.double(), prior.precisions.double()
    lognll = torch.log(prior.nll_weights.double()).reshape(-1)

    def ll(p):
        d = p[None, :] - means
        return 0.5 * torch.einsum('mi,mij,mj->m', d, prec, d) - lognll

    out = []
    while len(out) < count:
        a = torch.tensor(0.25 * rs.randn(69)) + means[rs.randint(8)]
        b = torch.tensor(0.25 * rs.randn(69)) + means[rs.randint(8)]
        ia, ib = int(ll(a).argmin()), int(ll(b).argmin())
        if ia == ib:
            continue
        lo, hi = 0.0, 1.0
        for _ in range(60):                                   # the winner at lo is ia, at hi it is something else
            mid = 0.5 * (lo + hi)
            if int(ll((1 - mid) * a + mid * b).argmin()) == ia:
                lo = mid
            else:
                hi = mid
        p_tie = (1 - hi) * a + hi * b
        v = ll(p_tie)
        two = torch.topk(-v, 2).indices
        # step away from the exact tie until the gap is about rel_gap of the value
        direction = (b - a) * (1.0 if rs.rand() < 0.5 else -1.0)
        step = 1e-9
        p = p_tie
        for _ in range(200):
            p = p_tie + step * direction
            v = ll(p)
            srt = torch.sort(v).values
            if float(srt[1] - srt[0]) >= rel_gap * abs(float(srt[0])):
                break
            step *= 1.3
        del two
        out.append(p.float())
    return torch.stack(out)


def main():
    from inbed_pose_estimation_b200 import synthetic
    os.makedirs(GOLDEN, exist_ok=True)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    cpu = torch.device('cpu')
    rs = np.random.RandomState(11)

    # ---- everything that needs the DENSE model / no model: geometry out_3d, near-tie prior ----------------
    tmp = tempfile.mkdtemp(prefix='inbed_ref_r2_')
    synthetic.write_data_dir(tmp, seed=0)
    ns = import_reference(tmp)

    g = {}
    for tag, npts in (('verts', 6890), ('joints', 49)):
        B = 3 if npts > 100 else 8
        pts = torch.tensor((0.5 * rs.randn(B, npts, 3)).astype(np.float32), requires_grad=True)
        rot = ns.geometry.batch_rodrigues(torch.tensor(0.3 * rs.randn(B, 3).astype(np.float32))).detach()
        rot.requires_grad_(True)
        tr = torch.tensor((np.array([0, 0, 20.]) + 0.1 * rs.randn(B, 3)).astype(np.float32), requires_grad=True)
        cen = torch.tensor((112 + rs.randn(B, 2)).astype(np.float32))
        pr = ns.geometry.perspective_projection(pts, rot, tr, 5000., cen, out_3d=True)
        gP = torch.tensor(rs.randn(B, npts, 3).astype(np.float32))
        (pr * gP).sum().backward()
        g.update({tag + '_points': pts.detach().numpy(), tag + '_rotation': rot.detach().numpy(),
                  tag + '_translation': tr.detach().numpy(), tag + '_center': cen.numpy(),
                  tag + '_projected': pr.detach().numpy(), tag + '_grad_projected': gP.numpy(),
                  tag + '_grad_points': pts.grad.numpy(), tag + '_grad_rotation': rot.grad.numpy(),
                  tag + '_grad_translation': tr.grad.numpy()})
    np.savez_compressed(os.path.join(GOLDEN, 'geometry_out3d.npz'), **g)

    prior = ns.prior.MaxMixturePrior(prior_folder='data', num_gaussians=8, dtype=torch.float32)
    gaps = [3e-6] * 24 + [1e-5] * 8
    poses = torch.cat([near_tie_poses(prior, rs, 1, gp) for gp in gaps])
    bp = poses.clone().requires_grad_(True)
    betas = torch.tensor((0.5 * rs.randn(len(gaps), 10)).astype(np.float32), requires_grad=True)
    nll = prior(bp, betas)                                                   # prior.py:227-229 -> merged_log_likelihood
    angle = ns.losses.angle_prior(bp).sum(dim=-1)                            # losses.py:19-24
    shape = (betas ** 2).sum(dim=-1)
    total = (4.78 ** 2) * nll + (15.2 ** 2) * angle + (5 ** 2) * shape         # the three prior terms of losses.py:46-52
    total.sum().backward()
    with torch.no_grad():                                                    # per-component values by the file's own buffers / ops
        diff = bp.unsqueeze(1) - prior.means
        quad = (torch.einsum('mij,bmj->bmi', [prior.precisions, diff]) * diff).sum(dim=-1)
        comp = 0.5 * quad - torch.log(prior.nll_weights)
    srt = torch.sort(comp, dim=1).values
    np.savez(os.path.join(GOLDEN, 'prior_near_tie.npz'), body_pose=poses.numpy(), betas=betas.detach().numpy(),
             nll=nll.detach().numpy(), angle=angle.detach().numpy(), shape=shape.detach().numpy(),
             components=comp.numpy(), argmin=comp.argmin(dim=1).numpy(), grad_body_pose=bp.grad.numpy(),
             grad_betas=betas.grad.numpy(), rel_gap=((srt[:, 1] - srt[:, 0]) / srt[:, 0].abs()).numpy())
    print('near-tie prior: relative gaps between the two best components %.2e .. %.2e' %
          (float(((srt[:, 1] - srt[:, 0]) / srt[:, 0].abs()).min()), float(((srt[:, 1] - srt[:, 0]) / srt[:, 0].abs()).max())))

    # ---- the SPARSE-structured model: the reference modules read their data files at construction, so a second data
    # directory is enough (the modules themselves are already imported) --------------------------------------------
    tmp2 = tempfile.mkdtemp(prefix='inbed_ref_r2s_')
    synthetic.write_data_dir(tmp2, seed=5, structure='sparse')
    os.chdir(tmp2)
    smpl = ns.smpl.SMPL(ns.config.SMPL_MODEL_DIR, batch_size=4, create_transl=False)
    inp = synthetic.make_fit_inputs(4, seed=8)
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
        os.path.join(GOLDEN, 'smpl_forward_sparse.npz'), model_seed=5,
        pose=inp['pose'], betas=inp['betas'], vertex_stride=8,
        vertices_sub=out.vertices.detach().numpy()[:, sub], joints=out.joints.detach().numpy(),
        vertices_checksum=out.vertices.detach().double().sum(dim=1).numpy(),
        grad_vertices=gv.numpy(), grad_joints=gj.numpy(), grad_pose=pose.grad.numpy(), grad_betas=betas.grad.numpy(),
        rotmats=rotm.detach().numpy(), vertices_rotmat_sub=out2.vertices.detach().numpy()[:, sub],
        joints_rotmat=out2.joints.detach().numpy(),
        grad_rotmats=rotm.grad.numpy(), grad_betas_rotmat=betas2.grad.numpy())

    fitter = ns.smplify.SMPLify(step_size=1e-2, batch_size=4, num_iters=100, focal_length=5000, device=cpu)
    for variant in ('default', 'slp'):
        inp = synthetic.make_fit_inputs(4, seed=13, variant=variant)
        (verts, joints, pose_o, betas_o, cam_o, reproj), trace = tapped_fit(ns, fitter, inp)
        np.savez_compressed(
            os.path.join(GOLDEN, 'smplify_sparse_%s.npz' % variant), model_seed=5,
            num_iters=100, seed=13, variant=variant,
            pose=inp['pose'], betas=inp['betas'], cam_t=inp['cam_t'], center=inp['center'], keypoints=inp['keypoints'],
            out_vertices_sub=verts.numpy()[:, sub], out_joints=joints.numpy(), out_pose=pose_o.numpy(),
            out_betas=betas_o.numpy(), out_cam_t=cam_o.detach().numpy(), out_reproj=reproj.numpy(), loss_trace=trace)
        print('sparse %-8s final mean reprojection %.4f  first/last loss %.4f / %.4f' %
              (variant, float(reproj.mean()), trace[0], trace[-1]))
    print('round-2 golden vectors written to', GOLDEN)


if __name__ == '__main__':
    main()
