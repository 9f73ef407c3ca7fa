"""GPU, world_size = 2 (skipped on single-GPU boxes): the ray-sharded tracking iteration on real devices -- in-kernel exchanges over
NVLink peer memory and the NCCL collectives -- against the single-GPU iteration over the whole batch, and the sharded masked mapping
iteration (one all-reduce of the packed block) against the single-GPU one."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import scene_util as su
    from gpu_util import make_renderer
    from nice_slam_b200.dist import ShardedMappingIteration, ShardedTrackingIteration, shard_bounds
    from nice_slam_b200.masked import MaskedVoxels
    from nice_slam_b200.steps import IterationContext
    sc = su.load_scenes()["room0"]
    renderer, c, dec = make_renderer(sc, su.make_grids(sc, "soft"), su.load_decoders("soft"), dev)
    n_shard = 150
    n = n_shard * world
    ro, rd, gd, gc = su.make_rays(sc, n, seed=55)
    dirs = torch.randn(n, 3, generator=torch.Generator().manual_seed(6))
    lo, hi = shard_bounds(n, rank, world)
    out = {}
    # ---- tracking: full batch on this GPU (reference), then the shard with both exchange back-ends
    full = IterationContext(renderer, n, "color", dev, kind="track")
    full.run(c, dec, ro.to(dev), rd.to(dev), gd.to(dev), gc.double().to(dev), dirs=dirs.to(dev))
    want = torch.cat([full.loss.reshape(1), full.d_c2w.reshape(-1)]).clone()
    ctx = IterationContext(renderer, n_shard, "color", dev, kind="track")
    ctx.load_device_inputs(ro[lo:hi].to(dev), rd[lo:hi].to(dev), gd[lo:hi].to(dev), gc[lo:hi].double().to(dev))
    for name in ("auto", "nccl", "auto_global_max"):
        sh = ShardedTrackingIteration(ctx, exchange=name.split("_")[0])
        sh.prepare(c, dec, dirs[lo:hi].to(dev), global_gt_depth=gd.to(dev) if name.endswith("global_max") else None)
        got = sh.enqueue().clone()
        out["track_" + name] = (float((got - want).abs().max() / want.abs().max()), sh.peers is not None)
        g = sh.build_graph()
        if g is not None:
            sh.packed.zero_(); g.replay(); torch.cuda.synchronize()
            out["track_" + name + "_graph"] = (float((sh.packed - want).abs().max() / want.abs().max()), True)
        # the shard's ray gradients are the full batch's rows [lo, hi) (the seeds use the batch-global median, so this checks the exchange too)
        out["track_" + name + "_d_rays"] = (float((ctx.d_rays_d - full.d_rays_d[lo:hi]).abs().max() / full.d_rays_d.abs().max()), True)
        del g
    # ---- mapping: frustum-like masks, packed block, ONE all-reduce
    keys = ("grid_middle", "grid_fine", "grid_color")
    mv = {}
    for k in keys:
        D, H, W = c[k].shape[2:]
        m = torch.zeros(D, H, W, dtype=torch.bool, device=dev)
        m[:, :, : int(0.6 * W)] = True
        mv[k] = MaskedVoxels(c[k], m)
    offs_full = torch.tensor([0, n // 2, n], dtype=torch.int32, device=dev)
    mfull = IterationContext(renderer, n, "color", dev, kind="map", grad_grids=keys, grad_decoders=("color",), masked=mv, n_frames=2)
    mfull.run(c, dec, ro.to(dev), rd.to(dev), gd.to(dev), gc.float().to(dev))
    wantp = mfull.finish_packed(dirs.to(dev), offs_full).clone()
    mctx = IterationContext(renderer, n_shard, "color", dev, kind="map", grad_grids=keys, grad_decoders=("color",), masked=mv, n_frames=2)
    mctx.load_device_inputs(ro[lo:hi].to(dev), rd[lo:hi].to(dev), gd[lo:hi].to(dev), gc[lo:hi].float().to(dev))
    # this rank's rays belong to keyframe `rank` (frame boundaries = shard boundaries here)
    offs = torch.tensor([0, n_shard, n_shard] if rank == 0 else [0, 0, n_shard], dtype=torch.int32, device=dev)
    ms = ShardedMappingIteration(mctx)
    ms.prepare(c, dec, dirs[lo:hi].to(dev), offs)
    gotp = ms.enqueue().clone()
    scale = wantp.abs().max()
    out["map_packed"] = (float((gotp - wantp).abs().max() / scale), True)
    out["map_loss"] = (abs(float(gotp[0] - wantp[0])) / abs(float(wantp[0])), True)
    ms1 = ShardedMappingIteration(mctx)
    ms1.prepare(c, dec, dirs[lo:hi].to(dev), offs, global_gt_depth=gd.to(dev))      # depth maxima from the full batch: ONE collective per iteration
    assert ms1.collectives_per_step == 1
    got1 = ms1.enqueue().clone()
    out["map_packed_one_collective"] = (float((got1 - wantp).abs().max() / scale), True)
    torch.cuda.synchronize()
    if rank == 0:
        q.put(out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_iterations_on_two_gpus_match_the_single_gpu_iterations():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
    for name, (err, _) in out.items():
        assert err < 1e-5, (name, err, out)
    assert "track_auto" in out and "track_nccl" in out and "map_packed" in out
