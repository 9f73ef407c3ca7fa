"""Test-side restatement of the caller glue around the render operator (NOT product code):
  camera tensor -> c2w (get_camera_from_tensor / quad2rotation, src/common.py:137-176),
  pixel indices -> rays (get_rays_from_uv, src/common.py:74-89), bbox pre-filter, losses.
`render` is pluggable so the same glue drives the oracle (CPU) and the fused renderer (GPU)."""
import torch

import scene_util as su
from oracle import torch_port as tp


def quad2rotation(q):
    bs = q.shape[0]
    qr, qi, qj, qk = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    two_s = 2.0 / (q * q).sum(-1)
    R = torch.zeros(bs, 3, 3, device=q.device)
    R[:, 0, 0] = 1 - two_s * (qj ** 2 + qk ** 2)
    R[:, 0, 1] = two_s * (qi * qj - qk * qr)
    R[:, 0, 2] = two_s * (qi * qk + qj * qr)
    R[:, 1, 0] = two_s * (qi * qj + qk * qr)
    R[:, 1, 1] = 1 - two_s * (qi ** 2 + qk ** 2)
    R[:, 1, 2] = two_s * (qj * qk - qi * qr)
    R[:, 2, 0] = two_s * (qi * qk - qj * qr)
    R[:, 2, 1] = two_s * (qj * qk + qi * qr)
    R[:, 2, 2] = 1 - two_s * (qi ** 2 + qj ** 2)
    return R


def camera_from_tensor(t):
    t = t.unsqueeze(0)
    R = quad2rotation(t[:, :4])
    return torch.cat([R, t[:, 4:, None]], 2)[0]


def tracking_rays(sc, case, camera_tensor, device="cpu"):
    """Rays of the captured pixel picks inside the edge-cropped region (get_sample_uv, src/common.py:110-122)."""
    cam, tr = sc["cam"], sc["tracking"]
    H, W = cam["H"], cam["W"]
    He, We = tr["ignore_edge_H"], tr["ignore_edge_W"]
    depth, color = su.make_frame(sc, case["frame_seed"])
    depth, color = depth[He:H - He, We:W - We].to(device), color[He:H - He, We:W - We].to(device)
    idx = case["pixel_idx"].to(device)
    wc = W - 2 * We
    i = (idx % wc).float() + We
    j = (idx // wc).float() + He
    c2w = camera_from_tensor(camera_tensor)
    dirs = torch.stack([(i - cam["cx"]) / cam["fx"], -(j - cam["cy"]) / cam["fy"], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs.reshape(-1, 1, 3) * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d, depth.reshape(-1)[idx], color.reshape(-1, 3)[idx]


def tracking_iteration(sc, case, render, bound, device="cpu"):
    cam = case["camera_tensor"].to(device).clone().requires_grad_(True)
    ro, rd, gd, gc = tracking_rays(sc, case, cam, device)
    keep = tp.bbox_prefilter(ro.cpu(), rd.cpu(), gd.cpu(), bound).to(device)
    ro, rd, gd, gc = ro[keep], rd[keep], gd[keep], gc[keep]
    depth, var, color = render(rd, ro, "color", gd)
    loss = tp.tracking_loss(depth, var, color, gd, gc, sc["tracking"]["w_color_loss"])
    loss.backward()
    return dict(loss=float(loss), d_camera=cam.grad.detach().cpu(), depth=depth.detach().cpu(), rays_o=ro.detach().cpu())


def tracking_iteration_cpu(sc, case, grids, dec):
    bound = su.scene_bound(sc)
    return tracking_iteration(sc, case, lambda rd, ro, stage, gd: tp.render_batch_ray(grids, dec, rd, ro, stage, gd, bound), bound)


def mapping_iteration_cpu(sc, case, grids, dec):
    bound = su.scene_bound(sc)
    stage = case["stage"]
    lv = {"coarse": ["coarse"], "middle": ["middle"], "fine": ["fine", "middle"], "color": ["fine", "color", "middle"]}[stage]
    return tp.iteration("map", grids, dec, case["rays_o"], case["rays_d"], case["gt_depth_loss"], case["gt_color"], stage, bound,
                        grad_grids=["grid_" + x for x in lv], grad_decoders=("color",) if stage == "color" else (),
                        grad_rays=False, w_color=sc["mapping"]["w_color_loss"])
