"""Sharded bulk refit (BASELINE config 4): every sample's fit is independent, so the batch is split
contiguously over the ranks of a torch.distributed group, fitted locally with no inner-loop traffic,
and the packed results are all-gathered once (pose 72 + betas 10 + camera 3 + reprojection 49 floats
per sample).  The keep-if-better rule is the reference's (train/trainer.py:716-727)."""
import torch
import torch.distributed as dist

PACKED = 72 + 10 + 3 + 49


def shard_bounds(n, world, rank):
    """Contiguous rows [lo, hi) of rank `rank` when n rows are split over `world` ranks."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def pack_results(pose, betas, cam_t, reproj):
    return torch.cat([pose, betas, cam_t, reproj], dim=1)


def unpack_results(packed):
    return packed[:, :72], packed[:, 72:82], packed[:, 82:85], packed[:, 85:]


def gather_rows(local, n_total, group=None):
    """All-gather a [n_local, C] tensor whose row counts follow shard_bounds -> [n_total, C] on every rank.
    When the rows divide evenly over the ranks the local tensor is gathered as it is (no padded copy)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank(group)
    sizes = [shard_bounds(n_total, world, r) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    if all(hi - lo == max_rows for lo, hi in sizes) and local.is_contiguous():
        out = local.new_empty((n_total, local.shape[1]))
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    padded = local.new_zeros((max_rows, local.shape[1]))
    padded[:local.shape[0]] = local
    out = local.new_empty((world * max_rows, local.shape[1]))
    dist.all_gather_into_tensor(out, padded, group=group)
    rows = [out[r * max_rows:r * max_rows + (hi - lo)] for r, (lo, hi) in enumerate(sizes)]
    assert rows[rank].shape[0] == local.shape[0]
    return torch.cat(rows, dim=0)


def keep_if_better(old_fits, old_loss, new_fits, new_loss):
    """update = new_loss < old_loss (trainer.py:719); masked overwrite (:722-727)."""
    update = new_loss < old_loss
    fits = torch.where(update[:, None], new_fits, old_fits)
    loss = torch.where(update, new_loss, old_loss)
    return fits, loss, update


def keep_better_(best_loss, best_pose, best_betas, best_cam_t, new_reproj, new_pose, new_betas, new_cam_t,
                 best_joints=None, new_joints=None):
    """In-place CUDA form of the reference's bookkeeping after SMPLify (train/trainer.py:716-727), one kernel:
    update = new_reproj.mean(-1) < best_loss; the best_* rows where it holds are overwritten.  Returns update (bool)."""
    import ctypes
    from . import _native
    dev = best_loss.device
    if dev.type != 'cuda':
        raise RuntimeError('keep_better_ runs on CUDA tensors only (no CPU fallback)')
    B = best_loss.shape[0]
    for t in (best_loss, best_pose, best_betas, best_cam_t) + ((best_joints,) if best_joints is not None else ()):
        if not (t.is_contiguous() and t.dtype == torch.float32 and t.device == dev):
            raise ValueError('best_* tensors must be contiguous fp32 CUDA tensors (they are updated in place)')
    c = lambda t: t.detach().to(dev).float().contiguous()
    update = torch.zeros(B, dtype=torch.uint8, device=dev)
    nj = c(new_joints) if (best_joints is not None and new_joints is not None) else None
    if B:
        with torch.cuda.device(dev):
            _native.check(_native.lib().smplb200_keep_better(
                B, _native.ptr(c(new_reproj)), _native.ptr(c(new_pose)), _native.ptr(c(new_betas)), _native.ptr(c(new_cam_t)),
                _native.ptr(nj), _native.ptr(best_loss), _native.ptr(best_pose), _native.ptr(best_betas), _native.ptr(best_cam_t),
                _native.ptr(best_joints) if nj is not None else None, ctypes.c_void_p(update.data_ptr()),
                torch.cuda.current_stream(dev).cuda_stream))
    return update.bool()


class ShardedRefit(object):
    """refit(fits [N,82], cam_t [N,3], center [N,2], keypoints [N,49,3], old_loss [N]) on every rank.

    `fit_fn(pose, betas, cam_t, center, keypoints)` must return the SMPLify 6-tuple for its rows; it
    defaults to the CUDA SMPLify passed in.  Returns (fits [N,82], loss [N], updated [N] bool, cam_t [N,3]),
    identical on all ranks."""

    def __init__(self, smplify=None, fit_fn=None, group=None, device=None):
        self.fit_fn = fit_fn if fit_fn is not None else smplify
        self._packs = fit_fn is None and smplify is not None        # the CUDA SMPLify writes the packed rows itself
        self.group = group
        self.device = device if device is not None else (smplify.device if smplify is not None else None)

    def __call__(self, fits, cam_t, center, keypoints, old_loss):
        n = fits.shape[0]
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        lo, hi = shard_bounds(n, world, rank)
        dev = self.device if self.device is not None else fits.device
        sl = lambda t: t[lo:hi].to(dev, non_blocking=True).contiguous()
        if self._packs:
            local = torch.empty((hi - lo, PACKED), device=dev, dtype=torch.float32)
            self.fit_fn(sl(fits[:, :72]), sl(fits[:, 72:]), sl(cam_t), sl(center), sl(keypoints).clone(), packed_out=local)
        else:
            _, _, pose, betas, cam, reproj = self.fit_fn(sl(fits[:, :72]), sl(fits[:, 72:]), sl(cam_t), sl(center), sl(keypoints).clone())
            local = pack_results(pose, betas, cam.detach(), reproj)
        packed = gather_rows(local, n, self.group)
        pose, betas, cam, reproj = unpack_results(packed)
        new_loss = reproj.mean(dim=-1)                       # trainer.py:716
        new_fits = torch.cat([pose, betas], dim=1)
        fits_d, loss_d, update = keep_if_better(fits.to(packed.device), old_loss.to(packed.device), new_fits, new_loss)
        return fits_d, loss_d, update, cam
