"""GPU: bundle adjustment in the native mapper path (src/Mapper.py:346-379 camera tensors + pose parameter group, :437-467 per-frame rays,
:521-540 write-back) against goldens captured from the REAL Mapper.optimize_map with BA=True on a window of five keyframes + the current
frame (tests/make_golden.py: mapper_ba_grads.pt, mapper_ba_loop.pt), and the pose kernels against torch autograd / torch.optim.Adam."""
import ctypes as C
import os

import pytest
import torch

import scene_util as su
from gpu_util import make_renderer, rel
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu
DEV = "cuda"
VP = C.c_void_p


def _fixed_row(window_c2w, cams):
    poses = tp.camera_from_tensor(cams)
    for r in range(window_c2w.shape[0]):
        if not any(torch.allclose(poses[k], window_c2w[r], atol=1e-5) for k in range(poses.shape[0])):
            return r
    raise AssertionError


def test_window_rays_and_pose_adam_match_torch():
    """nsb_window_rays == get_camera_from_tensor + get_rays_from_uv (oracle restatement, f32 op order); nsb_adam_poses == autograd through
    quad2rotation followed by torch.optim.Adam, over several steps with changing learning rates (incl. 0: state advances, pose does not)."""
    from nice_slam_b200 import _lib
    L = _lib.lib()
    g = torch.Generator().manual_seed(5)
    F, n = 6, 160
    cams = torch.randn(5, 7, generator=g)
    cams[:, :4] = torch.nn.functional.normalize(cams[:, :4], dim=1) * (1 + 0.05 * torch.randn(5, 1, generator=g))      # not exactly unit norm
    fixed_row = 2
    fixed = torch.randn(F, 12, generator=g)
    cam_row = torch.tensor([0, 1, -1, 2, 3, 4], dtype=torch.int32)
    pi = torch.randint(0, 1200, (F * n,), generator=g).float()
    pj = torch.randint(0, 680, (F * n,), generator=g).float()
    fid = torch.arange(F, dtype=torch.int32).repeat_interleave(n)
    fx, fy, cx, cy = 600.0, 600.0, 599.5, 339.5
    d = {k: v.to(DEV) for k, v in dict(cams=cams.clone(), cam_row=cam_row, fixed=fixed, pi=pi, pj=pj, fid=fid).items()}
    c2w = torch.empty(F, 12, device=DEV); ro = torch.empty(F * n, 3, device=DEV); rd = torch.empty(F * n, 3, device=DEV); dirs = torch.empty(F * n, 3, device=DEV)
    _lib.check(L.nsb_window_rays(VP(d["cams"].data_ptr()), VP(d["cam_row"].data_ptr()), VP(d["fixed"].data_ptr()), F, VP(d["pi"].data_ptr()), VP(d["pj"].data_ptr()),
                                 VP(d["fid"].data_ptr()), F * n, fx, fy, cx, cy, VP(c2w.data_ptr()), VP(ro.data_ptr()), VP(rd.data_ptr()), VP(dirs.data_ptr()), None), "window_rays")
    wro, wrd = tp.ba_window_rays(cams, fixed[fixed_row].view(3, 4), fixed_row, pi, pj, fid, fx, fy, cx, cy)
    assert float((ro.cpu() - wro).abs().max()) <= 1e-6 and float((rd.cpu() - wrd).abs().max()) <= 2e-6
    want_dirs = torch.stack([(pi - cx) / fx, -(pj - cy) / fy, -torch.ones_like(pi)], -1)
    assert torch.equal(dirs.cpu(), want_dirs)
    # pose chain backward + Adam vs torch
    ref = cams.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=0.0)
    m = torch.zeros(5, 7, device=DEV); v = torch.zeros(5, 7, device=DEV); dc = torch.zeros(5, 7, device=DEV)
    for step, lr in enumerate([0.0, 0.0, 1e-3, 1e-3, 5e-4], start=1):
        G = torch.randn(F, 3, 4, generator=g) * (100.0 if step % 2 else 1.0)
        opt.param_groups[0]["lr"] = lr
        opt.zero_grad()
        rows = [r for r in range(F) if r != fixed_row]
        (tp.camera_from_tensor(ref) * G[rows]).sum().backward()
        want_grad = ref.grad.clone()
        opt.step()
        Gd = G.reshape(F, 12).to(DEV).contiguous()
        _lib.check(L.nsb_adam_poses(VP(d["cams"].data_ptr()), VP(d["cam_row"].data_ptr()), F, VP(Gd.data_ptr()), VP(m.data_ptr()), VP(v.data_ptr()), VP(dc.data_ptr()),
                                    lr, 0.9, 0.999, 1e-8, step, None), "adam_poses")
        assert rel(dc, want_grad) < 1e-5, (step, rel(dc, want_grad))
        assert float((d["cams"].cpu() - ref.detach()).abs().max()) < 2e-6, step


@pytest.mark.parametrize("stage", ["middle", "fine", "color"])
def test_ba_window_gradients_against_real_mapper(stage):
    """One iteration of the real mapper with BA (golden rays at the renderer boundary): fused mapping iteration -> per-frame d c2w
    (nsb_pose_grad_frames) -> quad2rotation chain (nsb_adam_poses, lr 0) == camera_tensor.grad of the five optimised frames; the compact
    voxel gradients and the colour-decoder gradients of the same iteration match too."""
    from nice_slam_b200 import _lib
    from nice_slam_b200._lib import LEVELS, flat_layout
    from nice_slam_b200.masked import MaskedVoxels
    from nice_slam_b200.steps import IterationContext
    L = _lib.lib()
    case = torch.load(os.path.join(su.GOLDEN, "mapper_ba_grads.pt"), map_location="cpu", weights_only=False)
    sc = su.load_scenes()[case["scene"]]
    renderer, c, dec = make_renderer(sc, su.make_grids(sc, case["variant"]), su.load_decoders(case["variant"]), DEV)
    st = case["stages"][stage]
    cam = sc["cam"]
    F = 6
    fixed = _fixed_row(case["window_c2w"], case["camera_tensors"])
    cam_row = torch.full((F,), -1, dtype=torch.int32)
    for k, r in enumerate([r for r in range(F) if r != fixed]):
        cam_row[r] = k
    cams = case["camera_tensors"].clone().to(DEV)
    cam_row_d, fixed_d = cam_row.to(DEV), case["window_c2w"].reshape(F, 12).contiguous().to(DEV)
    n = st["rays_o"].shape[0]
    pi, pj, fid = st["pix_i"].float().to(DEV).contiguous(), st["pix_j"].float().to(DEV).contiguous(), st["frame_of_ray"].to(DEV).contiguous()
    c2w = torch.empty(F, 12, device=DEV); ro = torch.empty(n, 3, device=DEV); rd = torch.empty(n, 3, device=DEV); dirs = torch.empty(n, 3, device=DEV)
    _lib.check(L.nsb_window_rays(VP(cams.data_ptr()), VP(cam_row_d.data_ptr()), VP(fixed_d.data_ptr()), F, VP(pi.data_ptr()), VP(pj.data_ptr()), VP(fid.data_ptr()), n,
                                 cam["fx"], cam["fy"], cam["cx"], cam["cy"], VP(c2w.data_ptr()), VP(ro.data_ptr()), VP(rd.data_ptr()), VP(dirs.data_ptr()), None), "window_rays")
    assert float((ro.cpu() - st["rays_o"]).abs().max()) <= 1e-6 and float((rd.cpu() - st["rays_d"]).abs().max()) <= 2e-6
    # the render itself takes the GOLDEN rays: the L1 losses make the pose gradient discontinuous in the rays (a 1-ulp change can flip a sign)
    keys = {"middle": ("grid_middle",), "fine": ("grid_middle", "grid_fine"), "color": ("grid_middle", "grid_fine", "grid_color")}[stage]
    mv = {k: MaskedVoxels(c[k], case["masks"][k]) for k in keys}
    ctx = IterationContext(renderer, n, stage, DEV, kind="map", grad_grids=keys, grad_decoders=("color",) if stage == "color" else (), masked=mv, n_frames=F)
    ctx.run(c, dec, st["rays_o"].to(DEV), st["rays_d"].to(DEV), st["gt_depth"].to(DEV), st["gt_color"].to(DEV))
    assert rel(ctx.depth, st["depth"]) < 1e-4 and rel(ctx.rgb, st["rgb"]) < 1e-4
    offs = torch.zeros(F + 1, dtype=torch.int32)
    offs[1:] = torch.bincount(st["frame_of_ray"].long(), minlength=F).cumsum(0).int()
    ctx.finish_packed(dirs, offs.to(DEV))
    m = torch.zeros(5, 7, device=DEV); v = torch.zeros(5, 7, device=DEV); dc = torch.zeros(5, 7, device=DEV)
    _lib.check(L.nsb_adam_poses(VP(cams.data_ptr()), VP(cam_row_d.data_ptr()), F, VP(ctx.d_frames.data_ptr()), VP(m.data_ptr()), VP(v.data_ptr()), VP(dc.data_ptr()),
                                0.0, 0.9, 0.999, 1e-8, 1, None), "adam_poses")
    # The mapping loss is an L1: a ray whose rendered depth (colour) is within rounding of the sensor value can take the other sign on the GPU,
    # which changes ITS gradient by 200 % and the window's pose gradient by ~2/976.  So: per-ray gradients against the oracle on the same rays
    # (all but a handful must agree to 1e-4 of the largest), and the camera gradients to the knife-edge budget of a few such rays.
    lv = {"middle": ("grid_middle",), "fine": ("grid_middle", "grid_fine"), "color": ("grid_middle", "grid_fine", "grid_color")}[stage]
    want = tp.iteration("map", su.make_grids(sc, case["variant"]), su.load_decoders(case["variant"]), st["rays_o"], st["rays_d"], st["gt_depth_loss"], st["gt_color"],
                        stage, su.scene_bound(sc), grad_grids=lv, grad_decoders=())
    per_ray = torch.cat([ctx.d_rays_o, ctx.d_rays_d], 1).cpu() - torch.cat([want["d_rays_o"], want["d_rays_d"]], 1)
    scale = float(torch.cat([want["d_rays_o"], want["d_rays_d"]], 1).abs().max())
    flipped = int((per_ray.abs().amax(1) > 1e-4 * scale).sum())
    assert flipped <= 3, flipped
    assert rel(dc, st["d_cameras"]) < 1e-4 + 4e-3 * flipped, (rel(dc, st["d_cameras"]), flipped)
    tol = 1e-4 + 4e-3 * flipped
    for k, summ in st["masked_grads"].items():
        got = mv[k].to_reference(ctx.d_grid[k]).cpu()
        mine = su.grid_summary(got, n_sample=4096)
        assert abs(mine["norm"] - summ["norm"]) < tol * summ["norm"], k
        assert rel(got[summ["idx"]], summ["val"]) < tol, k
    if stage == "color":
        lay = {nm: (off, cnt) for nm, off, cnt in flat_layout(LEVELS.index("color"))}
        for k, vgrad in st["d_color_decoder"].items():
            off, cnt = lay[k]
            assert rel(ctx.d_flat["color"][off:off + cnt].view_as(vgrad), vgrad) < tol, k


def test_ba_loop_against_real_mapper():
    """Eight joint iterations of the real mapper with BA and the real torch Adam (4 x middle, fine, 3 x color) against the native loop
    (mapping.FusedMappingLoop.iteration_ba): rays regenerated from the current camera tensors each iteration, bbox pre-filter, fused iteration,
    fused Adam on voxels / colour decoder / poses.  Sign-like first Adam steps + L1 losses: the poses must land within a few per cent of the
    reference's update, the fixed frame must not move."""
    from nice_slam_b200.mapping import FusedMappingLoop
    case = torch.load(os.path.join(su.GOLDEN, "mapper_ba_loop.pt"), map_location="cpu", weights_only=False)
    sc = su.load_scenes()[case["scene"]]
    renderer, c, dec = make_renderer(sc, su.make_grids(sc, case["variant"]), su.load_decoders(case["variant"]), DEV)
    depth, _ = su.make_frame(sc, case["frame_seed"])
    c2w_cur = su.make_pose(sc, case["pose_seed"])
    win = case["window_keyframes"]
    fixed = win.index(min(k for k in win if k >= 0))
    loop = FusedMappingLoop(renderer, c, dec, c2w_cur, depth.to(DEV), w_color=case["w_color_loss"])
    loop.enable_ba(case["window_c2w"], fixed, case["BA_cam_lr"], camera_tensors=case["camera_tensors"])
    for stage, lr, dr in zip(case["stages"], case["lrs"], case["draws"]):
        loss = loop.iteration_ba(stage, dr["i"].to(DEV), dr["j"].to(DEV), dr["depth"].to(DEV), dr["color"].to(DEV), lr)
        assert bool(torch.isfinite(loss).all())
    got = loop.window_c2w().cpu()
    want, start = case["final_c2w"], case["window_c2w"]
    assert torch.equal(got[fixed], start[fixed])
    for r in range(6):
        if r == fixed:
            continue
        upd_ref = (want[r] - start[r]).norm()
        assert float((got[r] - want[r]).norm()) < 0.1 * float(upd_ref), (r, float((got[r] - want[r]).norm()), float(upd_ref))
    for key, fin in case["final"].items():
        mv = loop.masked[key]
        after = mv.to_reference(mv.gather(c[key])).cpu()
        got_v, want_v = after[fin["idx"]], fin["val"]
        close = (got_v - want_v).abs() <= 2e-3 * (1 + want_v.abs())
        assert float(close.float().mean()) > 0.99, (key, float(close.float().mean()))
