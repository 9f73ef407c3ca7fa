"""CPU, world_size = 2, gloo: the ray-sharding logic of nice_slam_b200/dist.py (the N > 1 path of bench.py).

Each rank renders its shard with the oracle (CPU), uses the product's exchange helpers for the three batch-global
quantities (depth maxima MAX, residual all-gather for the median, SUM of loss + pose gradient) and the result must
equal the single-process iteration over the whole batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _tracking_shard(tp, grids, dec, bound, ro, rd, gd, gc, dirs, depth_max_ray, pool_fn):
    """Tracking iteration on a shard: depth maxima injected through an extra ray, median over `pool_fn(residuals)`."""
    ro2 = torch.cat([ro, depth_max_ray[0][None]]).requires_grad_(True)
    rd2 = torch.cat([rd, depth_max_ray[1][None]]).requires_grad_(True)
    gd2 = torch.cat([gd, depth_max_ray[2][None]])
    depth, var, color = tp.render_batch_ray(grids, dec, rd2, ro2, "color", gd2, bound)
    depth, var, color = depth[:-1], var[:-1], color[:-1]
    unc = var.detach()
    res = torch.abs(gd - depth) / torch.sqrt(unc + 1e-10)
    pool = pool_fn(res.detach())
    mask = (res < 10 * pool.median()) & (gd > 0)
    loss = res[mask].sum() + 0.5 * torch.abs(gc - color)[mask].sum()
    loss.backward()
    d_o, d_d = ro2.grad[:-1].double(), rd2.grad[:-1].double()
    d_c2w = torch.cat([d_d.t() @ dirs.double(), d_o.sum(0, keepdim=True).t()], 1)
    return torch.cat([loss.detach().reshape(1).double(), d_c2w.reshape(-1)])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scene_util as su
    from nice_slam_b200 import dist as nd
    from oracle import torch_port as tp
    sc = su.load_scenes()["room0"]
    grids, dec, bound = su.make_grids(sc, "soft"), su.load_decoders("soft"), su.scene_bound(sc)
    n = 48
    ro, rd, gd, gc = su.make_rays(sc, n, seed=21)
    keep = su.prefilter_host(ro, rd, gd, bound)
    ro, rd, gd, gc = ro[keep], rd[keep], gd[keep], gc[keep]
    n = ro.shape[0] - ro.shape[0] % world                   # equal shards (all_gather_into_tensor)
    ro, rd, gd, gc = ro[:n], rd[:n], gd[:n], gc[:n]
    dirs = torch.randn(n, 3, generator=torch.Generator().manual_seed(3))
    lo, hi = nd.shard_bounds(n, rank, world)
    assert (lo, hi) == (rank * n // world, (rank + 1) * n // world)
    # (1) depth maxima
    dm = torch.stack([gd[lo:hi].max(), (gd[lo:hi] * 1.2).max()])
    nd.exchange_depth_max(dm)
    assert float(dm[0]) == float(gd.max()) and float(dm[1]) == float((gd * 1.2).max())
    imax = int(torch.argmax(gd))
    extra = (ro[imax], rd[imax], gd[imax])
    # (2)+(3) sharded iteration with the product's exchange helpers
    packed = _tracking_shard(tp, grids, dec, bound, ro[lo:hi], rd[lo:hi], gd[lo:hi], gc[lo:hi], dirs[lo:hi], extra, nd.gather_residuals)
    nd.reduce_sum(packed)
    if rank == 0:
        full = _tracking_shard(tp, grids, dec, bound, ro, rd, gd, gc, dirs, extra, lambda r: r)
        q.put((packed, full))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_tracking_iteration_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    packed, full = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert torch.allclose(packed, full, rtol=1e-9, atol=1e-12), (packed, full)


def test_shard_bounds_cover_the_batch():
    from nice_slam_b200.dist import shard_bounds
    for n in (0, 1, 7, 200, 996, 5000):
        for w in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, w) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


# ------------------------------------------------------------------------------------------------ mapping
def _mapping_shard_packed(tp, steps, lib, grids, dec, bound, ro, rd, gd, gc, dirs, frame_of_ray, n_frames, masks, extra):
    """Mapping iteration (stage color) on a shard with the oracle; gradients packed the way steps.IterationContext lays them out:
    [loss | d c2w per keyframe | colour-decoder grads (canonical flat order) | compact masked voxel grads]."""
    keys = ("grid_fine", "grid_color", "grid_middle")
    g = {k: v.detach().clone().requires_grad_(k in keys) for k, v in grids.items()}
    dw = {n: {k: v.detach().clone().requires_grad_(n == "color") for k, v in W.items()} for n, W in dec.items()}
    ro2 = torch.cat([ro, extra[0][None]]).requires_grad_(True)
    rd2 = torch.cat([rd, extra[1][None]]).requires_grad_(True)
    depth, var, color = tp.render_batch_ray(g, dw, rd2, ro2, "color", torch.cat([gd, extra[2][None]]), bound)
    loss = tp.mapping_loss(depth[:-1], color[:-1], gd, gc, "color")
    loss.backward()
    counts = [(k, int(masks[k].sum())) for k in keys]
    sect, total = steps.packed_layout(n_frames, ("color",), counts)
    packed = torch.zeros(total)
    packed[0] = loss.detach().float()
    d_o, d_d = ro2.grad[:-1].double(), rd2.grad[:-1].double()
    fr = packed[sect["frames"][0]: sect["frames"][0] + 12 * n_frames].view(n_frames, 3, 4)
    for f in range(n_frames):
        sel = frame_of_ray == f
        fr[f] = torch.cat([d_d[sel].t() @ dirs[sel].double(), d_o[sel].sum(0, keepdim=True).t()], 1).float()
    off, nfl = sect["dec_color"]
    for name, o, cnt in lib.flat_layout(3):
        packed[off + o: off + o + cnt] = dw["color"][name].grad.reshape(-1)
    for k in keys:
        o, cnt = sect[k]
        dense = g[k].grad[0]                                   # [32,D,H,W]
        packed[o: o + cnt] = dense[:, masks[k]].t().reshape(-1)      # [n_selected,32], voxels in (d,h,w) order
    return packed


def _map_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import scene_util as su
    from nice_slam_b200 import _lib, steps
    from nice_slam_b200 import dist as nd
    from oracle import torch_port as tp
    sc = su.load_scenes()["room0"]
    grids, dec, bound = su.make_grids(sc, "soft"), su.load_decoders("soft"), su.scene_bound(sc)
    n, n_frames = 36, 3
    ro, rd, gd, gc = su.make_rays(sc, n, seed=33)
    gc = gc.float()
    dirs = torch.randn(n, 3, generator=torch.Generator().manual_seed(4))
    frame_of_ray = torch.arange(n) // (n // n_frames)
    gm = torch.Generator().manual_seed(5)
    masks = {k: torch.rand(grids[k].shape[2:], generator=gm) < 0.6 for k in ("grid_fine", "grid_color", "grid_middle")}
    lo, hi = nd.shard_bounds(n, rank, world)
    imax = int(torch.argmax(gd))
    extra = (ro[imax], rd[imax], gd[imax])
    sl = slice(lo, hi)
    packed = _mapping_shard_packed(tp, steps, _lib, grids, dec, bound, ro[sl], rd[sl], gd[sl], gc[sl], dirs[sl], frame_of_ray[sl], n_frames, masks, extra)
    nd.reduce_sum(packed)                                        # the ONE collective of a mapping iteration
    if rank == 0:
        full = _mapping_shard_packed(tp, steps, _lib, grids, dec, bound, ro, rd, gd, gc, dirs, frame_of_ray, n_frames, masks, extra)
        sect, _ = steps.packed_layout(n_frames, ("color",), [(k, int(masks[k].sum())) for k in ("grid_fine", "grid_color", "grid_middle")])
        errs = {"all": (float((packed - full).abs().max() / full.abs().max()), float((packed - full).norm() / full.norm()))}
        for name, (o, cnt) in sect.items():                      # per section, so a mis-laid-out section cannot hide behind a large one
            a, b = packed[o: o + cnt], full[o: o + cnt]
            errs[name] = (float((a - b).abs().max() / (b.abs().max() + 1e-30)), float(b.abs().max()))
        errs["loss"] = (abs(float(packed[0] - full[0])) / abs(float(full[0])), float(full[0]))
        q.put(errs)                                              # (the packed blocks are tens of MB: compare here, ship the verdict)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mapping_iteration_packed_allreduce_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_map_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    errs = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for name, (err, scale) in errs.items():
        assert err < 1e-5, (name, err, scale)
        assert scale > 0 or name == "all", name                  # every section carries signal
