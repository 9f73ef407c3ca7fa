"""ORACLE (test infrastructure, NOT product code) -- PyTorch-CPU restatement of NICE-SLAM's
render-and-backprop hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file.  The product path (nice_slam_b200/) never does; it fails loudly without its CUDA library.

Parity status: PINNED.  tests/make_golden.py runs this port side by side with the unmodified reference
imported from /root/reference (shims in tests/ref_harness.py) and asserts bit-identical z_vals / outputs /
gradients on CPU for every fixture it writes; tests/golden/*.pt hold those outputs of the real reference,
and tests/test_oracle_golden.py re-checks this port against them wherever /root/reference is absent.

The reference is pure Python over PyTorch; the arithmetic of grid_sample / Linear / sort / cumprod /
autograd lives in the third-party PyTorch wheel (reference pins pytorch=1.11.0, environment.yaml:81;
installed here: 2.11.0).  This port calls the same torch ops in the same order and dtypes, written as
plain functions over a dict of weight tensors instead of the reference's nn.Module classes.

Reference map (file:line under /root/reference):
  sample_z_vals        <- src/utils/Renderer.py:82-170
  in_bound_mask        <- src/utils/Renderer.py:43-46
  normalize_coords     <- src/common.py:269-284
  sample_grid          <- src/conv_onet/models/decoder.py:168-175 (MLP) / :255-260 (MLP_no_xyz)
  mlp_xyz              <- src/conv_onet/models/decoder.py:177-203
  mlp_no_xyz           <- src/conv_onet/models/decoder.py:262-274
  nice_forward         <- src/conv_onet/models/decoder.py:312-342
  eval_points          <- src/utils/Renderer.py:23-61
  composite            <- src/common.py:204-245 (occupancy branch)
  render_batch_ray     <- src/utils/Renderer.py:63-198 (N_importance == 0, perturb == 0, lindisp False)
  bbox_prefilter       <- src/Tracker.py:95-104, src/Mapper.py:471-481
  tracking_loss        <- src/Tracker.py:108-123
  mapping_loss         <- src/Mapper.py:487-493
"""
import torch
import torch.nn.functional as F

STAGES = ("coarse", "middle", "fine", "color")


# ----------------------------------------------------------------------------- weights
def decoder_state(module):
    """Flatten one reference decoder (MLP / MLP_no_xyz) into {name: tensor} (shares storage)."""
    d = {k: v for k, v in module.named_parameters()}
    for k, v in module.named_buffers():
        d[k] = v
    return d


def decoders_state(nice_module):
    out = {}
    for name in ("coarse", "middle", "fine", "color"):
        sub = getattr(nice_module, name + "_decoder", None)
        if sub is not None:
            out[name] = decoder_state(sub)
    return out


# ----------------------------------------------------------------------------- sampling
def sample_z_vals(rays_o, rays_d, gt_depth, bound, n_samples, n_surface, stage):
    """z_vals float64 [N, n_samples(+n_surface)]; dtype flow follows Renderer.py:82-170 exactly:
    near f32, far/t f64 (bound is f64), surface samples f64, merged by torch.sort."""
    if stage == "coarse":
        gt_depth = None
    if gt_depth is None:
        n_surface = 0
        near = 0.01
    else:
        gt_depth = gt_depth.reshape(-1, 1)
        near = gt_depth.repeat(1, n_samples) * 0.01
    with torch.no_grad():
        o = rays_o.detach().unsqueeze(-1)
        d = rays_d.detach().unsqueeze(-1)
        t = (bound.unsqueeze(0) - o) / d                       # (N,3,2) f64
        far_bb, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
        far_bb = far_bb.unsqueeze(-1)
        far_bb += 0.01
    if gt_depth is not None:
        far = torch.clamp(far_bb, 0, torch.max(gt_depth * 1.2))
    else:
        far = far_bb
    if n_surface > 0:
        nz = gt_depth > 0
        g = gt_depth[nz].unsqueeze(-1).repeat(1, n_surface)
        ts = torch.linspace(0., 1., steps=n_surface).double()
        z_nz = 0.95 * g * (1. - ts) + 1.05 * g * ts
        z_surf = torch.zeros(gt_depth.shape[0], n_surface).double()
        nz = nz.squeeze(-1)
        z_surf[nz, :] = z_nz
        z_zero = 0.001 * (1. - ts) + torch.max(gt_depth) * ts
        z_surf[~nz, :] = z_zero
    tv = torch.linspace(0., 1., steps=n_samples)
    z_vals = near * (1. - tv) + far * tv
    if n_surface > 0:
        z_vals, _ = torch.sort(torch.cat([z_vals, z_surf.double()], -1), -1)
    return z_vals


def in_bound_mask(p, bound):
    mx = (p[:, 0] < bound[0][1]) & (p[:, 0] > bound[0][0])
    my = (p[:, 1] < bound[1][1]) & (p[:, 1] > bound[1][0])
    mz = (p[:, 2] < bound[2][1]) & (p[:, 2] > bound[2][0])
    return mx & my & mz


def normalize_coords(p, bound):
    p = p.clone().reshape(-1, 3)
    p[:, 0] = ((p[:, 0] - bound[0, 0]) / (bound[0, 1] - bound[0, 0])) * 2 - 1.0
    p[:, 1] = ((p[:, 1] - bound[1, 0]) / (bound[1, 1] - bound[1, 0])) * 2 - 1.0
    p[:, 2] = ((p[:, 2] - bound[2, 0]) / (bound[2, 1] - bound[2, 0])) * 2 - 1.0
    return p


def sample_grid(p, grid, bound):
    """p [1,P,3] f64 -> [P,C] f32 trilinear features (border padding, align_corners)."""
    vgrid = normalize_coords(p, bound).unsqueeze(0)[:, :, None, None].float()
    c = F.grid_sample(grid, vgrid, padding_mode="border", align_corners=True, mode="bilinear")
    return c.squeeze(-1).squeeze(-1).transpose(1, 2).squeeze(0)


# ----------------------------------------------------------------------------- decoders
def mlp_xyz(p, grids, W, name, bound, concat_middle=False, color=False):
    c = sample_grid(p, grids["grid_" + name], bound)
    if concat_middle:
        with torch.no_grad():
            cm = sample_grid(p, grids["grid_middle"], bound)
        c = torch.cat([c, cm], dim=1)
    pf = p.float().squeeze(0)
    emb = torch.sin(pf @ W["embedder._B"])
    h = emb
    for i in range(5):
        h = F.relu(F.linear(h, W["pts_linears.%d.weight" % i], W["pts_linears.%d.bias" % i]))
        h = h + F.linear(c, W["fc_c.%d.weight" % i], W["fc_c.%d.bias" % i])
        if i == 2:
            h = torch.cat([emb, h], -1)
    out = F.linear(h, W["output_linear.weight"], W["output_linear.bias"])
    return out if color else out.squeeze(-1)


def mlp_no_xyz(p, grids, W, bound_coarse):
    c = sample_grid(p, grids["grid_coarse"], bound_coarse)
    h = c
    for i in range(5):
        h = F.relu(F.linear(h, W["pts_linears.%d.weight" % i], W["pts_linears.%d.bias" % i]))
        if i == 2:
            h = torch.cat([c, h], -1)
    return F.linear(h, W["output_linear.weight"], W["output_linear.bias"]).squeeze(-1)


def nice_forward(p, grids, dec, stage, bound, coarse_enlarge=2):
    """p [1,P,3] f64 -> raw [P,4] f32 = (r,g,b,occ_logit)."""
    if stage == "coarse":
        occ = mlp_no_xyz(p, grids, dec["coarse"], bound * coarse_enlarge)
        raw = torch.zeros(occ.shape[0], 4)
        raw[..., -1] = occ
        return raw
    if stage == "middle":
        occ = mlp_xyz(p, grids, dec["middle"], "middle", bound)
        raw = torch.zeros(occ.shape[0], 4)
        raw[..., -1] = occ
        return raw
    if stage == "fine":
        fine = mlp_xyz(p, grids, dec["fine"], "fine", bound, concat_middle=True)
        raw = torch.zeros(fine.shape[0], 4)
        mid = mlp_xyz(p, grids, dec["middle"], "middle", bound)
        raw[..., -1] = fine + mid
        return raw
    if stage == "color":
        fine = mlp_xyz(p, grids, dec["fine"], "fine", bound, concat_middle=True)
        raw = mlp_xyz(p, grids, dec["color"], "color", bound, color=True)
        mid = mlp_xyz(p, grids, dec["middle"], "middle", bound)
        raw[..., -1] = fine + mid
        return raw
    raise ValueError(stage)


def eval_points(p, grids, dec, stage, bound):
    mask = in_bound_mask(p, bound)
    ret = nice_forward(p.unsqueeze(0), grids, dec, stage, bound)
    ret[~mask, 3] = 100
    return ret


# ----------------------------------------------------------------------------- compositing
def composite(raw, z_vals):
    """Occupancy branch of raw2outputs_nerf_color: alpha = sigmoid(10*occ)."""
    rgb = raw[..., :-1]
    raw[..., 3] = torch.sigmoid(10 * raw[..., -1])
    alpha = raw[..., -1]
    ones = torch.ones((alpha.shape[0], 1)).float()
    weights = alpha.float() * torch.cumprod(torch.cat([ones, (1. - alpha + 1e-10).float()], -1).float(), -1)[:, :-1]
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    tmp = z_vals - depth_map.unsqueeze(-1)
    depth_var = torch.sum(weights * tmp * tmp, dim=1)
    return depth_map, depth_var, rgb_map, weights


def render_batch_ray(grids, dec, rays_d, rays_o, stage, gt_depth, bound, n_samples=32, n_surface=16,
                     return_aux=False):
    n_rays = rays_o.shape[0]
    z_vals = sample_z_vals(rays_o, rays_d, gt_depth, bound, n_samples, n_surface, stage)
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z_vals[..., :, None]
    raw = eval_points(pts.reshape(-1, 3), grids, dec, stage, bound).reshape(n_rays, z_vals.shape[1], -1)
    raw_keep = raw.detach().clone() if return_aux else None
    depth, var, rgb, weights = composite(raw, z_vals)
    if return_aux:
        return depth, var, rgb, dict(z_vals=z_vals, weights=weights, raw=raw_keep)
    return depth, var, rgb


# ----------------------------------------------------------------------------- callers' glue
def bbox_prefilter(rays_o, rays_d, gt_depth, bound):
    with torch.no_grad():
        o = rays_o.detach().unsqueeze(-1)
        d = rays_d.detach().unsqueeze(-1)
        t = (bound.unsqueeze(0) - o) / d
        t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
        return t >= gt_depth


def quad2rotation(quad):
    """[B,4] (qw,qx,qy,qz), not necessarily normalised -> [B,3,3]   (src/common.py:137-160)."""
    qr, qi, qj, qk = quad[:, 0], quad[:, 1], quad[:, 2], quad[:, 3]
    two_s = 2.0 / (quad * quad).sum(-1)
    rot = torch.zeros(quad.shape[0], 3, 3, dtype=quad.dtype)
    rot[:, 0, 0] = 1 - two_s * (qj ** 2 + qk ** 2)
    rot[:, 0, 1] = two_s * (qi * qj - qk * qr)
    rot[:, 0, 2] = two_s * (qi * qk + qj * qr)
    rot[:, 1, 0] = two_s * (qi * qj + qk * qr)
    rot[:, 1, 1] = 1 - two_s * (qi ** 2 + qk ** 2)
    rot[:, 1, 2] = two_s * (qj * qk - qi * qr)
    rot[:, 2, 0] = two_s * (qi * qk - qj * qr)
    rot[:, 2, 1] = two_s * (qj * qk + qi * qr)
    rot[:, 2, 2] = 1 - two_s * (qi ** 2 + qj ** 2)
    return rot


def camera_from_tensor(cam):
    """[7] or [B,7] = [quaternion | translation] -> c2w [3,4] / [B,3,4]   (get_camera_from_tensor, src/common.py:163-176)."""
    one = cam.dim() == 1
    x = cam.unsqueeze(0) if one else cam
    rt = torch.cat([quad2rotation(x[:, :4]), x[:, 4:, None]], 2)
    return rt[0] if one else rt


def rays_from_uv(i, j, c2w, fx, fy, cx, cy):
    """get_rays_from_uv (src/common.py:74-89): pixel coordinates (float tensors) and one pose -> rays_o, rays_d."""
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1).reshape(-1, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    return c2w[:3, -1].expand(rays_d.shape), rays_d


def ba_window_rays(cams, fixed_c2w, fixed_row, pix_i, pix_j, frame_of_ray, fx, fy, cx, cy):
    """Rays of a bundle-adjustment window (src/Mapper.py:437-467 after the bbox pre-filter): window row r uses the pose of camera tensor
    cams[r'] (rows in order, skipping `fixed_row`, whose pose is the constant fixed_c2w -- the oldest frame, Mapper.py:350).  Differentiable
    w.r.t. cams.  Returns rays_o, rays_d [N,3] in the given ray order."""
    n_rows = cams.shape[0] + (0 if fixed_row is None else 1)
    ro = torch.zeros(pix_i.shape[0], 3, dtype=cams.dtype)
    rd = torch.zeros(pix_i.shape[0], 3, dtype=cams.dtype)
    k = 0
    for r in range(n_rows):
        if r == fixed_row:
            c2w = fixed_c2w
        else:
            c2w = camera_from_tensor(cams[k]); k += 1
        sel = torch.nonzero(frame_of_ray == r).reshape(-1)
        o, d = rays_from_uv(pix_i[sel], pix_j[sel], c2w, fx, fy, cx, cy)
        ro = ro.index_put((sel,), o.to(ro.dtype))
        rd = rd.index_put((sel,), d.to(rd.dtype))
    return ro, rd


def tracking_loss(depth, var, color, gt_depth, gt_color, w_color=0.5, handle_dynamic=True, use_color=True):
    unc = var.detach()
    if handle_dynamic:
        tmp = torch.abs(gt_depth - depth) / torch.sqrt(unc + 1e-10)
        mask = (tmp < 10 * tmp.median()) & (gt_depth > 0)
    else:
        mask = gt_depth > 0
    loss = (torch.abs(gt_depth - depth) / torch.sqrt(unc + 1e-10))[mask].sum()
    if use_color:
        loss = loss + w_color * torch.abs(gt_color - color)[mask].sum()
    return loss


def mapping_loss(depth, color, gt_depth, gt_color, stage, w_color=0.2):
    m = gt_depth > 0
    loss = torch.abs(gt_depth[m] - depth[m]).sum()
    if stage == "color":
        loss = loss + w_color * torch.abs(gt_color - color).sum()
    return loss


def iteration(kind, grids, dec, rays_o, rays_d, gt_depth, gt_color, stage, bound,
              n_samples=32, n_surface=16, grad_grids=(), grad_decoders=(), grad_rays=True, w_color=None):
    """One fwd + loss + bwd pass the way Tracker.optimize_cam_in_batch (kind='track') or
    Mapper.optimize_map (kind='map') do it, after ray generation and the bbox pre-filter.
    Returns dict(loss, depth, var, color, d_rays_o, d_rays_d, d_grid_*, d_dec[name][param])."""
    rays_o = rays_o.detach().clone().requires_grad_(grad_rays)
    rays_d = rays_d.detach().clone().requires_grad_(grad_rays)
    g = {k: v.detach().clone().requires_grad_(k in grad_grids) for k, v in grids.items()}
    dw = {n: {k: v.detach().clone().requires_grad_(n in grad_decoders) for k, v in W.items()} for n, W in dec.items()}
    depth, var, color = render_batch_ray(g, dw, rays_d, rays_o, stage,
                                         None if (kind == "map" and stage == "coarse") else gt_depth,
                                         bound, n_samples, n_surface)
    if kind == "track":
        loss = tracking_loss(depth, var, color, gt_depth, gt_color, 0.5 if w_color is None else w_color)
    else:
        loss = mapping_loss(depth, color, gt_depth, gt_color, stage, 0.2 if w_color is None else w_color)
    loss.backward()
    out = dict(loss=loss.detach(), depth=depth.detach(), var=var.detach(), color=color.detach())
    if grad_rays:
        out["d_rays_o"], out["d_rays_d"] = rays_o.grad, rays_d.grad
    for k in grad_grids:
        out["d_" + k] = g[k].grad
    out["d_dec"] = {n: {k: v.grad for k, v in dw[n].items() if v.grad is not None} for n in grad_decoders}
    return out
