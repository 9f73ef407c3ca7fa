"""Deterministic synthetic scenes shared by tests/make_golden.py, the parity tests, smoke() and bench.py.

Nothing here touches /root/reference: bounds / grid shapes computed once with the reference's own formulas
(NICE_SLAM.load_bound / grid_init, src/NICE_SLAM.py:137-157,192-250) are stored in tests/golden/scenes.json by
make_golden.py; grids are regenerated from seeds with CPU torch ops (bit-reproducible for a given torch build).
"""
import json
import os

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")
LEVELS = ("coarse", "middle", "fine", "color")


def load_scenes():
    with open(os.path.join(GOLDEN, "scenes.json")) as f:
        return json.load(f)


def scene_bound(scene):
    """float64 [3,2] bound tensor exactly as the reference computed it (stored as hex floats)."""
    return torch.tensor([[float.fromhex(v) for v in row] for row in scene["bound_hex"]], dtype=torch.float64)


def make_grids(scene, variant="soft", seed=0, keys=("grid_coarse", "grid_middle", "grid_fine", "grid_color")):
    """{key: float32 [1,32,D,H,W]} contiguous NCDHW CPU tensors.
    variant 'init' : the reference's initial state, N(0, 0.01) (fine: 1e-4)  (src/NICE_SLAM.py:223,231,239,247)
    variant 'soft' : smooth low-frequency feature field + small noise -> occupancies vary along rays."""
    out = {}
    for n, key in enumerate(keys):
        D, H, W = scene["shapes"][key]
        g = torch.Generator().manual_seed(1000 * seed + n)
        std = 1e-4 if key == "grid_fine" else 0.01
        noise = torch.randn(1, 32, D, H, W, generator=g) * std
        if variant == "init":
            out[key] = noise
        elif variant == "soft":
            lo = torch.randn(1, 32, max(D // 4, 2), max(H // 4, 2), max(W // 4, 2), generator=g)
            amp = 0.3
            out[key] = (F.interpolate(lo, size=(D, H, W), mode="trilinear", align_corners=True) * amp + noise * 30).contiguous()
        else:
            raise ValueError(variant)
    return out


def load_decoders(variant="soft"):
    """{level: {param name: tensor}} : pretrained coarse/middle/fine + seed-initialised colour decoder
    (tests/golden/decoders.pt, extracted by make_golden.py).  'soft' shifts the middle decoder's output bias so
    that free space is not saturated (alpha = sigmoid(10*occ) stays in its sensitive range)."""
    st = torch.load(os.path.join(GOLDEN, "decoders.pt"), map_location="cpu", weights_only=True)
    st = {lvl: {k: v.clone() for k, v in sd.items()} for lvl, sd in st.items()}
    if variant == "soft":
        st["middle"]["output_linear.bias"] -= 1.05
    return st


def make_frame(scene, seed=0):
    """Synthetic RGB-D frame: depth U(0.5,3.5) with ~2 % zeros (f32 [H,W]), colour U(0,1) (f64 [H,W,3])."""
    H, W = scene["cam"]["H"], scene["cam"]["W"]
    g = torch.Generator().manual_seed(7000 + seed)
    depth = torch.rand(H, W, generator=g) * 3 + 0.5
    depth[torch.rand(H, W, generator=g) < 0.02] = 0
    color = torch.rand(H, W, 3, generator=g, dtype=torch.float64)
    return depth, color


def make_pose(scene, seed=0):
    """c2w [4,4] f32: small rotation about the bound centre plus a seed-dependent offset."""
    b = scene_bound(scene)
    ctr = (b[:, 0] + b[:, 1]) / 2
    g = torch.Generator().manual_seed(9000 + seed)
    ang = (torch.rand(3, generator=g) - 0.5) * 0.6
    cx, cy, cz = torch.cos(ang)
    sx, sy, sz = torch.sin(ang)
    Rx = torch.tensor([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = torch.tensor([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = torch.tensor([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    c2w = torch.eye(4)
    c2w[:3, :3] = Rz @ Ry @ Rx
    c2w[:3, 3] = ctr.float() + (torch.rand(3, generator=g) - 0.5) * 1.0
    return c2w


def rays_from_pixels(scene, c2w, idx):
    """Rays of flat pixel indices idx (int64) -- restatement of get_rays_from_uv (src/common.py:74-89)."""
    cam = scene["cam"]
    W = cam["W"]
    i = (idx % W).float()
    j = (idx // W).float()
    dirs = torch.stack([(i - cam["cx"]) / cam["fx"], -(j - cam["cy"]) / cam["fy"], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs.reshape(-1, 1, 3) * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o.contiguous(), rays_d.contiguous()


def make_rays(scene, n, seed=0, frame_seed=0, pose_seed=0):
    """n random pixels of a synthetic frame -> (rays_o, rays_d, gt_depth f32, gt_color f64), bbox pre-filter NOT applied."""
    depth, color = make_frame(scene, frame_seed)
    c2w = make_pose(scene, pose_seed)
    H, W = scene["cam"]["H"], scene["cam"]["W"]
    g = torch.Generator().manual_seed(11000 + seed)
    idx = torch.randint(H * W, (n,), generator=g)
    ro, rd = rays_from_pixels(scene, c2w, idx)
    return ro, rd, depth.reshape(-1)[idx], color.reshape(-1, 3)[idx]


def prefilter_host(rays_o, rays_d, gt_depth, bound):
    """Host-side data preparation for synthetic batches: keep rays whose bound exit lies beyond the sensor depth
    (same rule as src/Tracker.py:95-104; plain torch, used only to BUILD benchmark/test inputs)."""
    t = (bound.unsqueeze(0) - rays_o.unsqueeze(-1)) / rays_d.unsqueeze(-1)
    t, _ = torch.min(torch.max(t, dim=2)[0], dim=1)
    return t >= gt_depth


def grid_summary(t, seed=0, n_sample=2048):
    """Compact fingerprint of a dense gradient grid: sum, L2 norm, 3 random projections, sampled entries."""
    flat = t.detach().double().reshape(-1).cpu()
    g = torch.Generator().manual_seed(4242 + seed)
    proj = []
    for _ in range(3):
        r = torch.randn(flat.numel(), generator=g, dtype=torch.float32).double()
        proj.append(float((flat * r).sum()))
    nz = torch.nonzero(flat).reshape(-1)
    pick = nz[torch.randperm(nz.numel(), generator=g)[:n_sample]] if nz.numel() > 0 else nz
    return dict(sum=float(flat.sum()), norm=float(flat.norm()), proj=proj, nnz=int(nz.numel()),
                idx=pick.clone(), val=flat[pick].float().clone())


