"""GPU, ONE device: the in-kernel peer-memory exchange protocol of the *_peers kernels (include/nice_slam_b200.h, nsb_seeds.cuh) with two
"ranks" on two streams of the same GPU and two local exchange buffers (hand-built nsb_peers).  Covers what tests/test_gpu_multi.py
covers on a 2-GPU box -- sequence numbers, parity double-buffering, the pooled median, the rank-ordered sums, graph-free re-use of the
buffers over several rounds -- on the single-GPU box the driver's GPU test tier runs on.  The waits are bounded (a missing rank traps the
launch instead of hanging the device)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

from nice_slam_b200 import _lib            # noqa: E402

VP = C.c_void_p
DEV = "cuda"


def _peers(rank, world, bufs, counters, max_rays):
    p = _lib.Peers()
    p.rank, p.world, p.max_rays = rank, world, max_rays
    for r in range(world):
        p.buffer[r] = bufs[r].data_ptr()
    p.counters = counters.data_ptr()
    return p


@pytest.mark.parametrize("world,n", [(2, 150), (2, 400), (3, 64), (4, 300)])
def test_peer_exchange_kernels_two_ranks_on_one_gpu(world, n):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    L = _lib.lib()
    g = torch.Generator().manual_seed(100 * world + n)
    N = world * n
    gt = (torch.rand(N, generator=g) * 3 + 0.5)
    gt[torch.rand(N, generator=g) < 0.05] = 0
    depth = (gt + 0.1 * torch.randn(N, generator=g)).double()
    var = (torch.rand(N, generator=g) * 0.05 + 1e-3).double()
    rgb = torch.rand(N, 3, generator=g)
    gt_rgb = torch.rand(N, 3, generator=g, dtype=torch.float64)
    dirs = torch.randn(N, 3, generator=g)
    dro = torch.randn(N, 3, generator=g)
    drd = torch.randn(N, 3, generator=g)
    to = lambda t: t.to(DEV).contiguous()
    gt, depth, var, rgb, gt_rgb, dirs, dro, drd = map(to, (gt, depth, var, rgb, gt_rgb, dirs, dro, drd))
    # ---- single-rank reference over the whole batch
    dm = torch.zeros(2, device=DEV)
    _lib.check(L.nsb_batch_max_depth(VP(gt.data_ptr()), N, VP(dm.data_ptr()), None), "batch_max")
    gD = torch.empty(N, dtype=torch.float64, device=DEV); gC = torch.empty(N, 3, device=DEV); loss = torch.zeros(1, dtype=torch.float64, device=DEV)
    ws = torch.empty(L.nsb_tracking_seeds_workspace(N), dtype=torch.uint8, device=DEV)
    _lib.check(L.nsb_tracking_seeds(VP(depth.data_ptr()), VP(var.data_ptr()), VP(rgb.data_ptr()), VP(gt.data_ptr()), VP(gt_rgb.data_ptr()), N, 0.5, 1, 1,
                                    None, 0, VP(gD.data_ptr()), VP(gC.data_ptr()), VP(loss.data_ptr()), VP(ws.data_ptr()), ws.numel(), None), "seeds")
    pose = torch.zeros(12, dtype=torch.float64, device=DEV)
    _lib.check(L.nsb_pose_grad(VP(dirs.data_ptr()), VP(dro.data_ptr()), VP(drd.data_ptr()), N, VP(pose.data_ptr()), None), "pose_grad")
    torch.cuda.synchronize()
    # ---- `world` ranks on one GPU: one stream, one exchange buffer, one counter block each
    nbytes = L.nsb_peer_buffer_bytes(n)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=DEV) for _ in range(world)]
    ctrs = [torch.zeros(4, dtype=torch.int64, device=DEV) for _ in range(world)]
    peers = [_peers(r, world, bufs, ctrs[r], n) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    out = [dict(dm=torch.zeros(2, device=DEV), gD=torch.empty(n, dtype=torch.float64, device=DEV), gC=torch.empty(n, 3, device=DEV),
                loss=torch.zeros(1, dtype=torch.float64, device=DEV), ws=torch.empty(max(L.nsb_tracking_seeds_workspace(n), 16), dtype=torch.uint8, device=DEV),
                sum13=torch.zeros(13, dtype=torch.float64, device=DEV)) for _ in range(world)]
    torch.cuda.synchronize()
    for rnd in range(3):                                      # three rounds: sequence numbers advance, both parities are used, nothing is reset
        for r in range(world):                                # all ranks' kernels are enqueued before anything synchronises
            sl = slice(r * n, (r + 1) * n)
            st = VP(streams[r].cuda_stream)
            o = out[r]
            _lib.check(L.nsb_batch_max_depth_peers(VP(gt[sl].data_ptr()), n, VP(o["dm"].data_ptr()), C.byref(peers[r]), st), "batch_max_peers")
            _lib.check(L.nsb_tracking_seeds_peers(VP(depth[sl].data_ptr()), VP(var[sl].data_ptr()), VP(rgb[sl].data_ptr()), VP(gt[sl].data_ptr()),
                                                  VP(gt_rgb[sl].data_ptr()), n, 0.5, 1, 1, C.byref(peers[r]), VP(o["gD"].data_ptr()), VP(o["gC"].data_ptr()),
                                                  VP(o["loss"].data_ptr()), VP(o["ws"].data_ptr()), o["ws"].numel(), st), "seeds_peers")
            _lib.check(L.nsb_pose_grad_peers(VP(dirs[sl].data_ptr()), VP(dro[sl].data_ptr()), VP(drd[sl].data_ptr()), n, VP(o["loss"].data_ptr()),
                                             VP(o["sum13"].data_ptr()), C.byref(peers[r]), st), "pose_grad_peers")
        torch.cuda.synchronize()
        for r in range(world):
            sl = slice(r * n, (r + 1) * n)
            o = out[r]
            assert torch.equal(o["dm"], dm), (rnd, r)                                        # MAX over ranks of [max gt, fl(1.2 max gt)]
            assert torch.equal(o["gD"], gD[sl]) and torch.equal(o["gC"], gC[sl]), (rnd, r)   # same median => same mask => same seeds
            assert torch.equal(o["sum13"], out[0]["sum13"])                                  # rank-ordered sums: identical bits on every rank
        tot = out[0]["sum13"]
        assert abs(float(tot[0] - loss[0])) <= 1e-9 * abs(float(loss[0]))
        assert float((tot[1:] - pose).abs().max()) <= 1e-9 * float(pose.abs().max())
        assert int(ctrs[0][0]) == rnd + 1 and int(ctrs[0][1]) == rnd + 1 and int(ctrs[0][2]) == rnd + 1


@pytest.mark.parametrize("world,n,global_max", [(2, 100, False), (2, 37, False), (2, 100, True), (2, 100, "list"), (4, 160, "list")])
def test_two_launch_sharded_tracking_iteration_on_one_gpu(world, n, global_max):
    """nsb_tracking_iteration_peers (forward + backward launches with the exchanges inside the tile kernels) for `world` ranks on ONE GPU
    (one stream and one set of buffers per rank) == the single-rank tracking iteration over the whole batch: same depth maxima (hence
    bit-identical samples), same median, same seeds, [loss | d c2w] summed over the ranks; replayed three times.  global_max: the depth maxima of the
    FULL batch are handed in (nsb_batch_max_depth over all ranks' depths) and the depth-max exchange (channel 0) is not used."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import scene_util as su
    from gpu_util import make_renderer, rel
    from nice_slam_b200.renderer import _inputs, _linspaces
    from nice_slam_b200.steps import IterationContext
    L = _lib.lib()
    sc = su.load_scenes()["room0"]
    renderer, c, dec = make_renderer(sc, su.make_grids(sc, "soft"), su.load_decoders("soft"), DEV)
    N = world * n
    ro, rd, gd, gc = su.make_rays(sc, N, seed=77)
    dirs = torch.randn(N, 3, generator=torch.Generator().manual_seed(8))
    full = IterationContext(renderer, N, "color", DEV, kind="track")
    full.run(c, dec, ro.to(DEV), rd.to(DEV), gd.to(DEV), gc.double().to(DEV), dirs=dirs.to(DEV))
    torch.cuda.synchronize()
    want = torch.cat([full.loss.reshape(1), full.d_c2w.reshape(-1)]).clone()
    want_rays = torch.cat([full.d_rays_o, full.d_rays_d], 1).clone()
    nbytes = L.nsb_peer_buffer_bytes(n)
    bufs = [torch.zeros(nbytes, dtype=torch.uint8, device=DEV) for _ in range(world)]
    ctrs = [torch.zeros(4, dtype=torch.int64, device=DEV) for _ in range(world)]
    peers = [_peers(r, world, bufs, ctrs[r], n) for r in range(world)]
    streams = [torch.cuda.Stream() for _ in range(world)]
    ranks = []
    for r in range(world):
        sl = slice(r * n, (r + 1) * n)
        x = IterationContext(renderer, n, "color", DEV, kind="track")
        x.load_device_inputs(ro[sl].to(DEV), rd[sl].to(DEV), gd[sl].to(DEV), gc[sl].double().to(DEV))
        vro, vrd, vgd, vgc = x.device_views()
        call, grids, _ = renderer._call(c, dec, "color", vgd, torch.device(DEV))
        t_u, t_s = _linspaces(renderer.N_samples, renderer.N_surface, torch.device(DEV))
        inp = _inputs(call, vro, vrd, None, t_u, t_s, [g.detach() for g in grids])
        if global_max == "list":                                   # the kernel reduces the full batch's depths itself
            gd_all = gd.to(DEV)
            inp.gt_depth_batch, inp.n_batch = gd_all.data_ptr(), N
        elif global_max:
            gd_all = gd.to(DEV)
            _lib.check(L.nsb_batch_max_depth(VP(gd_all.data_ptr()), N, VP(x.depth_max.data_ptr()), VP(torch.cuda.current_stream().cuda_stream)), "batch_max")
            inp.depth_max = x.depth_max.data_ptr()
        bw = x._grads(c)
        d = dirs[sl].to(DEV).contiguous()
        bw.pose_dirs, bw.d_c2w, bw.pose_counter = d.data_ptr(), x.d_c2w.data_ptr(), x.pose_counter.data_ptr()
        ranks.append(dict(x=x, inp=inp, bw=bw, gc=vgc, dirs=d, keep=(call, grids, t_u, t_s, gd_all if global_max else None), out=torch.zeros(13, dtype=torch.float64, device=DEV)))
    torch.cuda.synchronize()
    for rnd in range(3):
        for r in range(world):
            k = ranks[r]
            _lib.check(L.nsb_tracking_iteration_peers(C.byref(k["inp"]), C.byref(k["x"].buf), VP(k["gc"].data_ptr()), 0.5, 1, 1, C.byref(k["bw"]),
                                                      C.byref(peers[r]), VP(k["out"].data_ptr()), VP(streams[r].cuda_stream)), "tracking_iteration_peers")
        torch.cuda.synchronize()
        for r in range(world):
            k = ranks[r]
            assert torch.equal(k["out"], ranks[0]["out"])
            sl = slice(r * n, (r + 1) * n)
            got_rays = torch.cat([k["x"].d_rays_o, k["x"].d_rays_d], 1)
            assert rel(got_rays, want_rays[sl]) < 1e-5, (rnd, r, rel(got_rays, want_rays[sl]))
        assert rel(ranks[0]["out"], want) < 1e-6, (rnd, rel(ranks[0]["out"], want))
        assert all(int(ctrs[r][ch]) == rnd + 1 for r in range(world) for ch in ((1, 2) if global_max else (0, 1, 2)))
