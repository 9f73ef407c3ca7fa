"""Generate tests/golden/* by running the UNMODIFIED reference (imported from /root/reference through
tests/ref_harness.py) on deterministic synthetic inputs (tests/scene_util.py).

    python tests/make_golden.py          # build container only (needs /root/reference)

Outputs (committed):
    scenes.json        bounds (hex f64) + grid shapes from the reference's own load_bound / grid_init
    decoders.pt        decoder weights: pretrained coarse/middle/fine (pretrained/*.pt) + seed-0 colour decoder
    render_*.pt        Renderer.render_batch_ray outputs + autograd gradients for given output seeds
    tracker_color.pt   one real Tracker.optimize_cam_in_batch iteration, captured at the renderer boundary
    mapper_*.pt        real Mapper.optimize_map iterations (middle / fine / color, coarse mapper), idem
Dense grid gradients are stored as fingerprints (scene_util.grid_summary); the GPU tests additionally compare the
full tensors against the oracle run live.
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import ref_harness as rh        # noqa: E402
import scene_util as su         # noqa: E402
from oracle import torch_port as tp   # noqa: E402

GOLD = su.GOLDEN
SCENES = {"room0": "configs/Replica/room0.yaml", "scene0000": "configs/ScanNet/scene0000.yaml",
          "apartment": "configs/Apartment/apartment.yaml"}


def save(name, obj):
    path = os.path.join(GOLD, name)
    torch.save(obj, path)
    print("wrote %-22s %8.1f KB" % (name, os.path.getsize(path) / 1024))


def gen_scenes():
    out = {}
    for name, yaml_path in SCENES.items():
        cfg = rh.load_cfg(yaml_path)
        slam = rh.build_slam(cfg, seed=0)
        out[name] = dict(
            yaml=yaml_path,
            bound_hex=[[float(v).hex() for v in row] for row in slam.bound.tolist()],
            bound=[[float(v) for v in row] for row in slam.bound.tolist()],
            shapes={k: list(v.shape[2:]) for k, v in slam.shared_c.items()},
            cam=dict(H=slam.H, W=slam.W, fx=slam.fx, fy=slam.fy, cx=slam.cx, cy=slam.cy),
            rendering=dict(cfg["rendering"]), coarse_bound_enlarge=cfg["model"]["coarse_bound_enlarge"],
            tracking=dict(pixels=cfg["tracking"]["pixels"], w_color_loss=cfg["tracking"]["w_color_loss"],
                          ignore_edge_W=cfg["tracking"]["ignore_edge_W"], ignore_edge_H=cfg["tracking"]["ignore_edge_H"]),
            mapping=dict(pixels=cfg["mapping"]["pixels"], w_color_loss=cfg["mapping"]["w_color_loss"]))
    with open(os.path.join(GOLD, "scenes.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote scenes.json")
    return out


def gen_decoders():
    cfg = rh.load_cfg(SCENES["room0"])
    slam = rh.build_slam(cfg, seed=0)
    st = tp.decoders_state(slam.shared_decoders)
    save("decoders.pt", {lvl: {k: v.detach().clone() for k, v in sd.items()} for lvl, sd in st.items()})


def ref_scene(scene_name, variant, grid_seed=0):
    """Reference-side objects (cfg, slam namespace with the reference's NICE module, Renderer) on a scene_util scene."""
    scenes = su.load_scenes()
    sc = scenes[scene_name]
    cfg = rh.load_cfg(sc["yaml"])
    slam = rh.build_slam(cfg, seed=0)
    assert torch.equal(slam.bound, su.scene_bound(sc))
    grids = su.make_grids(sc, variant, grid_seed)
    for k in slam.shared_c:
        assert list(slam.shared_c[k].shape[2:]) == sc["shapes"][k]
        slam.shared_c[k] = grids[k]
    st = su.load_decoders(variant)
    for lvl, sd in st.items():
        getattr(slam.shared_decoders, lvl + "_decoder").load_state_dict(sd)
    return sc, cfg, slam, rh.make_renderer(cfg, slam)


def corner_indices(pts, bound, shape):
    """(ix0,iy0,iz0) of F.grid_sample(align_corners=True, padding_mode='border') for f64 points, computed with
    torch ops in the reference's order: normalize_3d_coordinate (f64) -> .float() -> unnormalise -> clip -> floor."""
    ref = rh.import_reference()
    pn = ref.common.normalize_3d_coordinate(pts.clone(), bound).float()
    D, H, W = shape
    out = []
    for a, size in enumerate((W, H, D)):
        u = ((pn[:, a] + 1) / 2) * (size - 1)
        u = torch.clamp(u, 0, size - 1)          # min(size-1, max(u, 0))
        out.append(torch.floor(u).to(torch.int16))
    return torch.stack(out, -1)


def _render_case(scene_name, variant, stage, n, sc, cfg, slam, renderer, ray_seed=3):
    """One reference render_batch_ray + backward with random output seeds on `n` rays: inputs, outputs, z_vals, voxel-corner indices, all gradients."""
    primary = {"coarse": "grid_coarse", "middle": "grid_middle", "fine": "grid_fine", "color": "grid_fine"}
    ro, rd, gd, gc = su.make_rays(sc, n, seed=ray_seed)
    ro = ro.clone()
    ro[:8, 0] += 5.0               # some rays leave the bound early -> out-of-bound samples
    gt = None if stage == "coarse" else gd
    lv = {"coarse": ["coarse"], "middle": ["middle"], "fine": ["fine", "middle"], "color": ["fine", "color", "middle"]}[stage]
    ro1 = ro.clone().requires_grad_(True)
    rd1 = rd.clone().requires_grad_(True)
    c = {k: v.clone().requires_grad_(k[5:] in lv) for k, v in slam.shared_c.items()}
    for p in slam.shared_decoders.parameters():
        p.grad = None
        p.requires_grad_(True)
    d, u, col = renderer.render_batch_ray(c, slam.shared_decoders, rd1, ro1, "cpu", stage, gt_depth=gt)
    g = torch.Generator().manual_seed(55)
    gD = torch.randn(n, generator=g, dtype=torch.float64)
    gV = torch.randn(n, generator=g, dtype=torch.float64) * 3
    gC = torch.randn(n, 3, generator=g)
    ((d * gD).sum() + (u * gV).sum() + (col * gC).sum()).backward()
    z = tp.sample_z_vals(ro, rd, gt, slam.bound, cfg["rendering"]["N_samples"], cfg["rendering"]["N_surface"], stage)
    pts = (ro[:, None, :] + rd[:, None, :] * z[:, :, None]).reshape(-1, 3)
    bnd = slam.bound * 2 if stage == "coarse" else slam.bound
    cidx = corner_indices(pts, bnd, sc["shapes"][primary[stage]]).reshape(n, -1, 3)
    # sanity: the port reproduces the reference bit for bit on this very case
    d2, u2, c2 = tp.render_batch_ray(slam.shared_c, tp.decoders_state(slam.shared_decoders), rd, ro, stage, gt, slam.bound)
    assert torch.equal(d2, d.detach()) and torch.equal(u2, u.detach()) and torch.equal(c2, col.detach())
    return dict(scene=scene_name, variant=variant, stage=stage, rays_o=ro, rays_d=rd, gt_depth=gt, g_depth=gD, g_var=gV, g_rgb=gC,
                depth=d.detach(), var=u.detach(), rgb=col.detach(), z_vals=z, corner_idx=cidx,
                d_rays_o=ro1.grad, d_rays_d=rd1.grad,
                d_grid={k: su.grid_summary(c[k].grad) for k in c if c[k].grad is not None},
                d_dec={l: {k: v.grad.clone() for k, v in getattr(slam.shared_decoders, l + "_decoder").named_parameters()
                           if v.grad is not None} for l in lv})


def gen_render_cases():
    for variant in ("soft", "init"):
        sc, cfg, slam, renderer = ref_scene("room0", variant)
        for stage in ("coarse", "middle", "fine", "color"):
            if variant == "init" and stage != "color":
                continue
            save("render_%s_%s.pt" % (stage, variant), _render_case("room0", variant, stage, 96, sc, cfg, slam, renderer))


def gen_other_scene_cases():
    """render_color_soft_<scene>.pt: the unmodified reference Renderer + autograd on the ScanNet scene0000 and Apartment volumes (SURVEY.md 8d
    configs 3 and 4: other bounds, grid shapes and intrinsics than room0), stage color, 64 rays, voxel and decoder gradients included."""
    for name in ("scene0000", "apartment"):
        sc, cfg, slam, renderer = ref_scene(name, "soft")
        save("render_color_soft_%s.pt" % name, _render_case(name, "soft", "color", 64, sc, cfg, slam, renderer, ray_seed=5))


class RecOptim:
    """Stand-in optimiser: records gradients at step(), never updates (so every iteration sees the same scene)."""
    instances = []

    def __init__(self, groups, **kw):
        self.param_groups = [dict(g) for g in groups] if isinstance(groups, (list, tuple)) and groups and isinstance(groups[0], dict) \
            else [dict(params=list(groups), lr=kw.get("lr", 0))]
        self.steps = []
        RecOptim.instances.append(self)

    def zero_grad(self, set_to_none=True):
        for g in self.param_groups:
            for p in g["params"]:
                p.grad = None

    def step(self):
        self.steps.append([[None if p.grad is None else p.grad.detach().clone() for p in g["params"]] for g in self.param_groups])


def record_renderer(renderer):
    calls = []
    orig = renderer.render_batch_ray

    def wrapped(c, decoders, rays_d, rays_o, device, stage, gt_depth=None):
        ret = orig(c, decoders, rays_d, rays_o, device, stage, gt_depth=gt_depth)
        calls.append(dict(stage=stage, rays_o=rays_o.detach().clone(), rays_d=rays_d.detach().clone(),
                          gt_depth=None if gt_depth is None else gt_depth.detach().clone(),
                          depth=ret[0].detach().clone(), var=ret[1].detach().clone(), rgb=ret[2].detach().clone()))
        return ret
    renderer.render_batch_ray = wrapped
    return calls


def gen_tracker_case():
    sc, cfg, slam, renderer = ref_scene("room0", "soft")
    ref = rh.import_reference()
    tracker = rh.make_tracker(cfg, slam, renderer)
    calls = record_renderer(renderer)
    depth, color = su.make_frame(sc, 0)
    c2w = su.make_pose(sc, 0)
    cam = ref.common.get_tensor_from_camera(c2w).requires_grad_(True)      # [qw,qx,qy,qz,tx,ty,tz]
    opt = RecOptim([cam], lr=1e-3)
    picked = []
    _ri = torch.randint

    def randint_rec(*a, **k):
        r = _ri(*a, **k)
        picked.append(r.clone())
        return r
    torch.randint = randint_rec
    try:
        torch.manual_seed(123)
        loss = tracker.optimize_cam_in_batch(cam, color, depth, cfg["tracking"]["pixels"], opt)
    finally:
        torch.randint = _ri
    call = calls[0]
    save("tracker_color.pt", dict(scene="room0", variant="soft", camera_tensor=cam.detach().clone(), pixel_idx=picked[0],
                                  frame_seed=0, loss=float(loss), d_camera=opt.steps[0][0][0], **call))


def gen_mapper_cases():
    rh.import_reference()
    for coarse_mapper in (False, True):
        sc, cfg, slam, renderer = ref_scene("room0", "soft")
        mapper = rh.make_mapper(cfg, slam, renderer, coarse_mapper=coarse_mapper)
        calls = record_renderer(renderer)
        depth, color = su.make_frame(sc, 1)
        c2w = su.make_pose(sc, 1)
        RecOptim.instances.clear()
        _adam = torch.optim.Adam
        torch.optim.Adam = RecOptim
        import src.Mapper as mapper_mod
        samples = []
        _gs = mapper_mod.get_samples

        def get_samples_rec(*a, **k):
            r = _gs(*a, **k)
            samples.append([t.detach().clone() for t in r])
            return r
        mapper_mod.get_samples = get_samples_rec
        try:
            torch.manual_seed(321)
            n_it = 1 if coarse_mapper else 5       # 5 iterations: 0-2 middle, 3 fine, 4 color (ratios 0.4 / 0.6)
            mapper.optimize_map(n_it, 1.0, 0, color, depth, c2w, [], [], c2w)
        finally:
            torch.optim.Adam = _adam
            mapper_mod.get_samples = _gs
        opt = RecOptim.instances[-1]
        names = ["decoders", "grid_coarse", "grid_middle", "grid_fine", "grid_color"]
        # frustum masks (per voxel; identical across the 32 channels) recomputed with the reference's own function
        masks = {}
        for key, val in slam.shared_c.items():
            m = mapper.get_mask_from_c2w(c2w, key, val.shape[2:], depth.numpy())
            masks[key] = torch.from_numpy(np.ascontiguousarray(m)).permute(2, 1, 0).contiguous()      # [D,H,W] bool
        for it, call in enumerate(calls):
            stage = call["stage"]
            if not coarse_mapper and it in (1, 2):
                continue                     # keep one middle iteration
            grads = {}
            for gi, nm in enumerate(names):
                gl = opt.steps[it][gi]
                if nm == "decoders":
                    pn = [k for k, _ in slam.shared_decoders.color_decoder.named_parameters()]
                    grads["color_decoder"] = {k: g for k, g in zip(pn, gl) if g is not None}
                elif gl and gl[0] is not None:
                    grads[nm] = gl[0]            # masked parameter vector val[mask] (channel-major order of the bool mask)
            # per-ray ground truth the loss used (Mapper.py:459-481): get_samples output after the bbox pre-filter
            s_o, s_d, s_gd, s_gc = samples[it]
            keep = tp.bbox_prefilter(s_o.float(), s_d.float(), s_gd.float(), slam.bound)
            assert torch.equal(s_o.float()[keep], call["rays_o"]) and torch.equal(s_d.float()[keep], call["rays_d"])
            call = dict(call, gt_depth_loss=s_gd.float()[keep], gt_color=s_gc.float()[keep])
            case = dict(scene="room0", variant="soft", frame_seed=1, coarse_mapper=coarse_mapper, masks=masks,
                        masked_grads={k: v for k, v in grads.items() if k != "color_decoder"},
                        d_color_decoder=grads.get("color_decoder", {}), **call)
            # masked grads can be MBs: keep a fingerprint + a sample
            for k in list(case["masked_grads"].keys()):
                case["masked_grads"][k] = su.grid_summary(case["masked_grads"][k], n_sample=4096)
            save("mapper_%s.pt" % stage, case)


def gen_mapper_loop_case():
    """Five REAL joint iterations of Mapper.optimize_map with the real torch.optim.Adam (3 x middle, fine, color): per-iteration ray
    batches at the renderer boundary + the state the mapper leaves behind (selected voxels of every grid, colour decoder)."""
    rh.import_reference()
    sc, cfg, slam, renderer = ref_scene("room0", "soft")
    mapper = rh.make_mapper(cfg, slam, renderer, coarse_mapper=False)
    calls = record_renderer(renderer)
    depth, color = su.make_frame(sc, 1)
    c2w = su.make_pose(sc, 1)
    import src.Mapper as mapper_mod
    samples = []
    _gs = mapper_mod.get_samples

    def get_samples_rec(*a, **k):
        r = _gs(*a, **k)
        samples.append([t.detach().clone() for t in r])
        return r
    mapper_mod.get_samples = get_samples_rec
    start = {k: v.detach().clone() for k, v in slam.shared_c.items()}
    try:
        torch.manual_seed(321)
        mapper.optimize_map(5, 1.0, 0, color, depth, c2w, [], [], c2w)
    finally:
        mapper_mod.get_samples = _gs
    its = []
    for it, call in enumerate(calls):
        s_o, s_d, s_gd, s_gc = samples[it]
        keep = tp.bbox_prefilter(s_o.float(), s_d.float(), s_gd.float(), slam.bound)
        assert torch.equal(s_o.float()[keep], call["rays_o"])
        st = cfg["mapping"]["stage"][call["stage"]]
        its.append(dict(stage=call["stage"], rays_o=call["rays_o"], rays_d=call["rays_d"], gt_depth=call["gt_depth"],
                        gt_depth_loss=s_gd.float()[keep], gt_color=s_gc.float()[keep],
                        lr=dict(decoders=st["decoders_lr"], middle=st["middle_lr"], fine=st["fine_lr"], color=st["color_lr"])))
    final = {}
    for key in ("grid_middle", "grid_fine", "grid_color"):
        m = mapper.get_mask_from_c2w(c2w, key, slam.shared_c[key].shape[2:], depth.numpy())
        vm = torch.from_numpy(np.ascontiguousarray(m)).permute(2, 1, 0).contiguous()
        m5 = vm.unsqueeze(0).unsqueeze(0).expand_as(slam.shared_c[key])
        after, before = slam.shared_c[key][m5].detach().clone(), start[key][m5]
        g = torch.Generator().manual_seed(99)
        pick = torch.randperm(after.numel(), generator=g)[:8192].clone()          # (a view would drag the whole permutation into the file)
        final[key] = dict(idx=pick, val=after[pick].clone(), delta_norm=float((after - before).double().norm()), norm=float(after.double().norm()),
                          unselected_changed=bool((slam.shared_c[key][~m5] != start[key][~m5]).any()))
    dec_final = {k: v.detach().clone() for k, v in slam.shared_decoders.color_decoder.state_dict().items()}
    save("mapper_loop.pt", dict(scene="room0", variant="soft", frame_seed=1, pose_seed=1, iterations=its, final=final,
                                color_decoder=dec_final, w_color_loss=cfg["mapping"]["w_color_loss"]))


def _ba_setup():
    """Reference mapper with bundle adjustment on a window of 5 keyframes + the current frame (6 x 166 rays)."""
    rh.import_reference()
    sc, cfg, slam, renderer = ref_scene("room0", "soft")
    mapper = rh.make_mapper(cfg, slam, renderer, coarse_mapper=False, BA=True)
    mapper.mapping_window_size = 6                  # 4 overlap-selected keyframes + the last keyframe + the current frame
    mapper.keyframe_selection_method = "overlap"
    kf_dict, kf_list = [], []
    for k in range(5):
        depth, color = su.make_frame(sc, 10 + k)
        c2w = su.make_pose(sc, 10 + k)
        gt = c2w.clone()
        gt[:3, 3] += 0.01 * (k + 1)                 # est != gt, as in a real run (gt is only logged)
        kf_dict.append(dict(gt_c2w=gt, idx=5 * k, color=color, depth=depth, est_c2w=c2w.clone()))
        kf_list.append(5 * k)
    mapper.keyframe_dict, mapper.keyframe_list = kf_dict, kf_list
    depth, color = su.make_frame(sc, 1)
    c2w = su.make_pose(sc, 1)
    return sc, cfg, slam, renderer, mapper, kf_dict, kf_list, depth, color, c2w


def _record_ba(mapper_mod, samples, uv):
    """Wrap get_samples / get_sample_uv of src.Mapper / src.common so that every per-frame draw is recorded."""
    import src.common as common_mod
    _gs, _guv = mapper_mod.get_samples, common_mod.get_sample_uv

    def get_sample_uv_rec(*a, **k):
        r = _guv(*a, **k)
        uv.append((r[0].detach().clone(), r[1].detach().clone()))
        return r

    def get_samples_rec(*a, **k):
        n0 = len(uv)
        r = _gs(*a, **k)
        samples.append(dict(rays_o=r[0].detach().clone(), rays_d=r[1].detach().clone(), depth=r[2].detach().clone(), color=r[3].detach().clone(),
                            i=uv[n0][0], j=uv[n0][1], c2w=a[11].detach().clone()))
        return r
    mapper_mod.get_samples = get_samples_rec
    common_mod.get_sample_uv = get_sample_uv_rec

    def restore():
        mapper_mod.get_samples, common_mod.get_sample_uv = _gs, _guv
    return restore


def gen_mapper_ba_cases():
    """(a) mapper_ba_grads.pt: one window of REAL Mapper.optimize_map iterations with BA=True whose optimiser only records (no updates):
    per stage the ray batch at the renderer boundary, the frame of every ray, and every gradient incl. camera_tensor.grad of the five
    non-fixed frames.  (b) mapper_ba_loop.pt: eight real iterations with the real torch.optim.Adam (4 x middle, fine, 3 x color): the pixel
    draws of every frame and iteration, and what the mapper leaves behind (poses of the window, selected voxels, colour decoder)."""
    ref = rh.import_reference()
    import src.Mapper as mapper_mod
    # ---------------------------------------------------------------- (a) gradients
    sc, cfg, slam, renderer, mapper, kf_dict, kf_list, depth, color, c2w = _ba_setup()
    calls = record_renderer(renderer)
    samples, uv = [], []
    restore = _record_ba(mapper_mod, samples, uv)
    RecOptim.instances.clear()
    _adam = torch.optim.Adam
    torch.optim.Adam = RecOptim
    try:
        torch.manual_seed(777); np.random.seed(777)
        mapper.optimize_map(5, 1.0, 30, color, depth, c2w, kf_dict, kf_list, c2w)
    finally:
        torch.optim.Adam = _adam
        restore()
    opt = RecOptim.instances[-1]
    sel = samples[:1]                                  # first get_samples call = keyframe_selection_overlap's own draw
    per_it = [samples[1 + 6 * it: 1 + 6 * (it + 1)] for it in range(5)]
    assert len(samples) == 1 + 6 * 5 and len(opt.param_groups) == 6 and len(opt.param_groups[5]["params"]) == 5
    masks = {}
    for key, val in slam.shared_c.items():
        m = mapper.get_mask_from_c2w(c2w, key, val.shape[2:], depth.numpy())
        masks[key] = torch.from_numpy(np.ascontiguousarray(m)).permute(2, 1, 0).contiguous()
    # window order and which frame is fixed: re-derive from the recorded per-frame poses of iteration 0
    frames_c2w = [s["c2w"][:3] for s in per_it[0]]
    cases = {}
    names = ["decoders", "grid_coarse", "grid_middle", "grid_fine", "grid_color", "cameras"]
    for it in (0, 3, 4):
        call = calls[it]
        fr = per_it[it]
        ro = torch.cat([f["rays_o"].float() for f in fr]); rd = torch.cat([f["rays_d"].float() for f in fr])
        gd = torch.cat([f["depth"].float() for f in fr]); gc = torch.cat([f["color"].float() for f in fr])
        fid = torch.cat([torch.full((f["rays_o"].shape[0],), k, dtype=torch.int32) for k, f in enumerate(fr)])
        keep = tp.bbox_prefilter(ro, rd, gd, slam.bound)
        assert torch.equal(ro[keep], call["rays_o"]) and torch.equal(rd[keep], call["rays_d"])
        grads = {}
        for gi, nm in enumerate(names):
            gl = opt.steps[it][gi]
            if nm == "decoders":
                pn = [k for k, _ in slam.shared_decoders.color_decoder.named_parameters()]
                grads["color_decoder"] = {k: g for k, g in zip(pn, gl) if g is not None}
            elif nm == "cameras":
                grads["cameras"] = torch.stack([g if g is not None else torch.zeros(7) for g in gl])
            elif gl and gl[0] is not None:
                grads[nm] = su.grid_summary(gl[0], n_sample=4096)
        cases[call["stage"]] = dict(rays_o=call["rays_o"], rays_d=call["rays_d"], gt_depth=call["gt_depth"], gt_depth_loss=gd[keep], gt_color=gc[keep],
                                    frame_of_ray=fid[keep], pix_i=torch.cat([f["i"] for f in fr])[keep], pix_j=torch.cat([f["j"] for f in fr])[keep],
                                    depth=call["depth"], rgb=call["rgb"], d_cameras=grads.pop("cameras"), d_color_decoder=grads.pop("color_decoder"),
                                    masked_grads=grads)
    cam0 = torch.stack([p.detach().clone() for p in opt.param_groups[5]["params"]])
    save("mapper_ba_grads.pt", dict(scene="room0", variant="soft", frame_seed=1, pose_seed=1, keyframe_seeds=list(range(10, 15)), masks=masks,
                                    window_c2w=torch.stack(frames_c2w), camera_tensors=cam0, stages=cases,
                                    selection_pixels=dict(i=sel[0]["i"], j=sel[0]["j"])))
    # ---------------------------------------------------------------- (b) real Adam, poses move
    sc, cfg, slam, renderer, mapper, kf_dict, kf_list, depth, color, c2w = _ba_setup()
    samples, uv = [], []
    restore = _record_ba(mapper_mod, samples, uv)
    start = {k: v.detach().clone() for k, v in slam.shared_c.items()}
    n_it = 8
    cam_hist = []
    _adam = torch.optim.Adam

    class SpyAdam(_adam):                              # the real Adam; only looks at the pose group before every step
        def step(self, *a, **k):
            cams = self.param_groups[5]["params"]
            cam_hist.append(dict(cam=torch.stack([c.detach().clone() for c in cams]), grad=torch.stack([c.grad.detach().clone() for c in cams])))
            return super().step(*a, **k)
    torch.optim.Adam = SpyAdam
    try:
        torch.manual_seed(777); np.random.seed(777)
        new_c2w = mapper.optimize_map(n_it, 1.0, 30, color, depth, c2w, kf_dict, kf_list, c2w)
    finally:
        torch.optim.Adam = _adam
        restore()
    per_it = [samples[1 + 6 * it: 1 + 6 * (it + 1)] for it in range(n_it)]
    window0 = torch.stack([s["c2w"][:3] for s in per_it[0]])        # poses the window started from (row order = optimize_frame order)
    draws = [dict(i=torch.stack([f["i"] for f in fr]), j=torch.stack([f["j"] for f in fr]),
                  depth=torch.stack([f["depth"].float() for f in fr]), color=torch.stack([f["color"].float() for f in fr])) for fr in per_it]
    # which keyframe each window row is: match the starting poses against the keyframe poses
    ident = []
    for r in range(6):
        hit = [k for k in range(5) if torch.allclose(su.make_pose(sc, 10 + k)[:3], window0[r][:3], atol=1e-5)]
        ident.append(hit[0] if hit else -1)
    final_c2w = torch.stack([(kf_dict[k]["est_c2w"] if k >= 0 else new_c2w)[:3].detach().clone() for k in ident])
    final = {}
    for key in ("grid_middle", "grid_fine", "grid_color"):
        m = mapper.get_mask_from_c2w(c2w, key, slam.shared_c[key].shape[2:], depth.numpy())
        vm = torch.from_numpy(np.ascontiguousarray(m)).permute(2, 1, 0).contiguous()
        m5 = vm.unsqueeze(0).unsqueeze(0).expand_as(slam.shared_c[key])
        after, before = slam.shared_c[key][m5].detach().clone(), start[key][m5]
        g = torch.Generator().manual_seed(99)
        pick = torch.randperm(after.numel(), generator=g)[:8192].clone()
        final[key] = dict(idx=pick, val=after[pick].clone(), delta_norm=float((after - before).double().norm()), norm=float(after.double().norm()))
    stages = ["middle" if it <= int(n_it * mapper.middle_iter_ratio) else ("fine" if it <= int(n_it * mapper.fine_iter_ratio) else "color") for it in range(n_it)]
    lrs = [dict(decoders=cfg["mapping"]["stage"][s]["decoders_lr"], middle=cfg["mapping"]["stage"][s]["middle_lr"], fine=cfg["mapping"]["stage"][s]["fine_lr"],
                color=cfg["mapping"]["stage"][s]["color_lr"]) for s in stages]
    save("mapper_ba_loop.pt", dict(scene="room0", variant="soft", frame_seed=1, pose_seed=1, keyframe_seeds=list(range(10, 15)), window_keyframes=ident,
                                   window_c2w=window0.clone(), final_c2w=final_c2w, camera_tensors=cam_hist[0]["cam"], camera_history=cam_hist, stages=stages, lrs=lrs, BA_cam_lr=cfg["mapping"]["BA_cam_lr"],
                                   draws=draws, final=final, color_decoder={k: v.detach().clone() for k, v in slam.shared_decoders.color_decoder.state_dict().items()},
                                   w_color_loss=cfg["mapping"]["w_color_loss"]))


def gen_keyframe_overlap_case():
    """keyframe_overlap.pt: the REAL Mapper.keyframe_selection_overlap (src/Mapper.py:166-228) on twelve synthetic keyframes -- some looking
    the same way as the current frame, some turned away (no overlap) -- with the pixel draw, the per-keyframe percent_inside and the
    selection the reference made under a fixed numpy seed."""
    rh.import_reference()
    import src.Mapper as mapper_mod
    sc, cfg, slam, renderer = ref_scene("room0", "soft")
    mapper = rh.make_mapper(cfg, slam, renderer)
    depth, color = su.make_frame(sc, 1)
    c2w = su.make_pose(sc, 1)
    g = torch.Generator().manual_seed(2024)
    kf = []
    for k in range(12):
        p = su.make_pose(sc, 40 + k)
        if k % 3 == 2:                                    # every third keyframe looks backwards: no overlap
            flip = torch.diag(torch.tensor([-1.0, 1.0, -1.0]))
            p[:3, :3] = p[:3, :3] @ flip
        p[:3, 3] += (torch.rand(3, generator=g) - 0.5) * 2.0
        kf.append(dict(est_c2w=p, gt_c2w=p.clone(), idx=5 * k, depth=depth, color=color))
    samples, recorded = [], []
    _gs = mapper_mod.get_samples

    def get_samples_rec(*a, **k):
        r = _gs(*a, **k)
        samples.append([t.detach().clone() for t in r])
        return r
    mapper_mod.get_samples = get_samples_rec
    # percent_inside is not returned by the reference: recompute it from the list it sorts by wrapping `sorted`
    import builtins
    _sorted = builtins.sorted

    def sorted_rec(it, *a, **k):
        it = list(it)
        if it and isinstance(it[0], dict) and "percent_inside" in it[0]:
            recorded.append([(d["id"], float(d["percent_inside"])) for d in it])
        return _sorted(it, *a, **k)
    mapper_mod.sorted = sorted_rec
    try:
        torch.manual_seed(99); np.random.seed(99)
        sel = mapper.keyframe_selection_overlap(color, depth, c2w, kf, 4)
    finally:
        mapper_mod.get_samples = _gs
        del mapper_mod.sorted
    ro, rd, gd, gc = samples[0]
    save("keyframe_overlap.pt", dict(scene="room0", frame_seed=1, pose_seed=1, keyframe_c2w=torch.stack([d["est_c2w"] for d in kf]),
                                     rays_o=ro, rays_d=rd, gt_depth=gd, percent_inside=[p for _, p in recorded[0]], k=4, numpy_seed=99,
                                     selected=[int(x) for x in sel]))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "scenes":
        gen_other_scene_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "keyframes":
        import warnings
        warnings.filterwarnings("ignore")
        gen_keyframe_overlap_case()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mapper_ba":
        import warnings
        warnings.filterwarnings("ignore")
        gen_mapper_ba_cases()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mapper_loop":
        import warnings
        warnings.filterwarnings("ignore")
        gen_mapper_loop_case()
        sys.exit(0)
    os.makedirs(GOLD, exist_ok=True)
    import warnings
    warnings.filterwarnings("ignore")
    gen_scenes()
    gen_decoders()
    gen_render_cases()
    gen_other_scene_cases()
    gen_tracker_case()
    gen_mapper_cases()
    gen_mapper_loop_case()
    gen_mapper_ba_cases()
    gen_keyframe_overlap_case()
