"""Ray-sharded data parallelism (one process per GPU, torch.distributed: NCCL on GPUs, gloo in the CPU tests).

Rays are independent given replicated (grids, decoders, pose); what is NOT per-ray in the reference and therefore
needs an exchange when a batch is sharded:
  * torch.max(gt_depth) / torch.max(gt_depth*1.2)        (src/utils/Renderer.py:109,144)  -> all-reduce MAX of 2 floats
  * tmp.median() of the tracking residuals               (src/Tracker.py:113)             -> all-gather of the residuals
  * the scalar loss and every gradient that is summed over rays (pose, voxel, decoder grads) -> all-reduce SUM
The exchange helpers below work on whatever device the tensors live on (so the gloo tests exercise the same code).
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n, rank, world_size):
    """Contiguous [lo, hi) slice of an n-ray batch for `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def exchange_depth_max(depth_max):
    """In place: element-wise MAX over ranks of [max(gt), max(gt*1.2)]."""
    if world()[1] > 1:
        dist.all_reduce(depth_max, op=dist.ReduceOp.MAX)
    return depth_max


def gather_residuals(res_local, counts=None):
    """All shards' tracking residuals, concatenated in rank order (equal shard sizes unless `counts` is given)."""
    rank, ws = world()
    if ws == 1:
        return res_local
    if counts is None:
        out = torch.empty(ws * res_local.numel(), dtype=res_local.dtype, device=res_local.device)
        dist.all_gather_into_tensor(out, res_local.contiguous())
        return out
    parts = [torch.empty(c, dtype=res_local.dtype, device=res_local.device) for c in counts]
    dist.all_gather(parts, res_local.contiguous())
    return torch.cat(parts)


def reduce_sum(packed):
    """In place SUM all-reduce of a packed gradient buffer ([loss | pose grads | decoder grads | voxel grads])."""
    if world()[1] > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    return packed


class ShardedTrackingIteration:
    """One tracking iteration on this rank's shard of a global ray batch (split-phase: the exchanges sit between the
    kernels).  With world_size == 1 it degenerates to the same kernels without collectives."""

    def __init__(self, ctx):
        self.ctx = ctx                      # steps.IterationContext(kind='track')
        dev = ctx.dev
        self.res = torch.empty(ctx.n, dtype=torch.float64, device=dev)
        self.packed = torch.zeros(13, dtype=torch.float64, device=dev)      # [loss | d_c2w(12)]

    def run(self, c, decoders, rays_o, rays_d, dirs, gt_depth, gt_color, w_color=0.5, handle_dynamic=True, use_color=True):
        import ctypes as C
        from . import _lib
        from .renderer import _VP, _inputs, _linspaces, _stream
        L = _lib.lib()
        x = self.ctx
        n = x.n
        call, grids, _ = x.r._call(c, decoders, x.stage, gt_depth, x.dev)
        t_u, t_s = _linspaces(x.r.N_samples, x.r.N_surface, x.dev)
        _lib.check(L.nsb_batch_max_depth(_VP(gt_depth.data_ptr()), n, _VP(x.depth_max.data_ptr()), _stream()), "nsb_batch_max_depth")
        exchange_depth_max(x.depth_max)
        inp = _inputs(call, rays_o, rays_d, x.depth_max, t_u, t_s, [g.detach() for g in grids])
        fo = _lib.ForwardOutputs(x.depth.data_ptr(), x.var.data_ptr(), x.rgb.data_ptr(), x.z_vals.data_ptr(), x.raw.data_ptr(), None, x.masks.data_ptr())
        _lib.check(L.nsb_render_forward(C.byref(inp), C.byref(fo), _stream()), "nsb_render_forward")
        pool, n_pool = None, 0
        if handle_dynamic and world()[1] > 1:
            _lib.check(L.nsb_tracking_residuals(_VP(x.depth.data_ptr()), _VP(x.var.data_ptr()), _VP(gt_depth.data_ptr()), n,
                                                _VP(self.res.data_ptr()), _stream()), "nsb_tracking_residuals")
            allres = gather_residuals(self.res)
            pool, n_pool = _VP(allres.data_ptr()), allres.numel()
        _lib.check(L.nsb_tracking_seeds(_VP(x.depth.data_ptr()), _VP(x.var.data_ptr()), _VP(x.rgb.data_ptr()), _VP(gt_depth.data_ptr()),
                                        _VP(gt_color.data_ptr()), n, w_color, int(handle_dynamic), int(use_color), pool, n_pool,
                                        _VP(x.g_depth.data_ptr()), _VP(x.g_rgb.data_ptr()), _VP(x.loss.data_ptr()),
                                        _VP(x.ws.data_ptr()), L.nsb_tracking_seeds_workspace(n), _stream()), "nsb_tracking_seeds")
        bw = x._grads(c)
        bw.z_vals, bw.raw, bw.g_depth, bw.g_rgb, bw.masks = x.z_vals.data_ptr(), x.raw.data_ptr(), x.g_depth.data_ptr(), x.g_rgb.data_ptr(), x.masks.data_ptr()
        _lib.check(L.nsb_render_backward(C.byref(inp), C.byref(bw), _stream()), "nsb_render_backward")
        _lib.check(L.nsb_pose_grad(_VP(dirs.data_ptr()), _VP(x.d_rays_o.data_ptr()), _VP(x.d_rays_d.data_ptr()), n,
                                   _VP(self.packed.data_ptr() + 8), _stream()), "nsb_pose_grad")
        self.packed[:1].copy_(x.loss)
        reduce_sum(self.packed)
        return self.packed                  # [global loss | global d_c2w]
