"""Frustum-selected voxel parameterisation (src/Mapper.py:317-333): during a mapping call only the voxels inside the current
frustum mask are parameters -- `val_grad = val[mask]` -- and every iteration writes them back into the shared grid
(`val[mask] = val_grad`, Mapper.py:393-401 and :511-519).

`MaskedVoxels` keeps that parameter vector as a compact [n_selected, 32] buffer (one 128-byte line per voxel) addressed through a
voxel -> slot table built on the device (nsb_voxel_slots).  The backward kernel scatters voxel gradients straight into a compact
gradient buffer of the same shape (nsb_backward_args.slot_map), which is what gets all-reduced across GPUs.
The reference orders `val[mask]` channel-major ([32][n_selected]); to_reference()/from_reference() convert.
"""
import ctypes as C

import torch

from . import _lib
from .renderer import _VP, _stream, grid_struct


class MaskedVoxels:
    def __init__(self, grid, voxel_mask):
        """grid: [1,32,D,H,W] CUDA tensor; voxel_mask: bool [D,H,W] or the reference's repeated [1,32,D,H,W] mask (Mapper.py:319-320)."""
        L = _lib.lib()
        if voxel_mask.dim() == 5:
            voxel_mask = voxel_mask[0, 0]
        if not grid.is_cuda:
            raise RuntimeError("nice_slam_b200: MaskedVoxels needs a CUDA grid (no CPU fallback)")
        D, H, W = grid.shape[2:]
        assert tuple(voxel_mask.shape) == (D, H, W)
        dev = grid.device
        m8 = voxel_mask.to(device=dev, dtype=torch.uint8).contiguous()
        n = D * H * W
        self.shape = (D, H, W)
        self.slot_map = torch.empty(n, dtype=torch.int32, device=dev)
        cnt = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(L.nsb_voxel_slots_workspace(n), dtype=torch.uint8, device=dev)
        _lib.check(L.nsb_voxel_slots(_VP(m8.data_ptr()), n, _VP(self.slot_map.data_ptr()), _VP(cnt.data_ptr()), _VP(ws.data_ptr()), ws.numel(),
                                     _stream()), "nsb_voxel_slots")
        self.count = int(cnt.item())                     # one host sync per mapping call (the reference does mask.sum() implicitly in val[mask])

    def empty(self, device=None):
        return torch.empty(self.count, 32, dtype=torch.float32, device=device or self.slot_map.device)

    def gather(self, grid, out=None):
        """compact = val[mask]  (slot-major)."""
        out = self.empty() if out is None else out
        g = grid_struct(grid.detach())
        _lib.check(_lib.lib().nsb_masked_gather(C.byref(g), _VP(self.slot_map.data_ptr()), _VP(out.data_ptr()), _stream()), "nsb_masked_gather")
        return out

    def scatter(self, grid, compact):
        """val[mask] = compact  (in place on the shared grid storage, like the reference)."""
        g = grid_struct(grid.detach())
        _lib.check(_lib.lib().nsb_masked_scatter(C.byref(g), _VP(self.slot_map.data_ptr()), _VP(compact.data_ptr()), _stream()), "nsb_masked_scatter")
        return grid

    def _transpose(self, src, to_ref):
        dst = torch.empty(src.numel(), dtype=torch.float32, device=src.device)
        _lib.check(_lib.lib().nsb_compact_transpose(_VP(src.data_ptr()), _VP(dst.data_ptr()), self.count, int(to_ref), _stream()),
                   "nsb_compact_transpose")
        return dst

    def to_reference(self, compact):
        """[n,32] -> the reference's 1-D `val[mask]` order ([32][n] flattened)."""
        return self._transpose(compact.contiguous(), True)

    def from_reference(self, vec):
        return self._transpose(vec.contiguous(), False).view(self.count, 32)


def frustum_voxel_mask(renderer, c2w, key, grid, depth):
    """Mapper.get_mask_from_c2w (src/Mapper.py:93-164) on the device: bool [D,H,W] mask of the voxels of `grid` ([1,32,D,H,W], key
    'grid_middle' / 'grid_fine' / 'grid_color' / 'grid_coarse') that the current frame (pose `c2w` [4,4], sensor `depth` [H,W]) can see.
    `renderer` supplies the scene bound and the intrinsics (the attributes the reference's Mapper copies from slam, Mapper.py:40-60)."""
    L = _lib.lib()
    if not grid.is_cuda:
        raise RuntimeError("nice_slam_b200: frustum_voxel_mask needs CUDA tensors (no CPU fallback)")
    dev = grid.device
    D, H, W = grid.shape[2:]
    if key == "grid_coarse":                                         # Mapper.py:114-116
        return torch.ones(D, H, W, dtype=torch.bool, device=dev)
    b = renderer.bound
    # voxel-centre coordinates exactly as the reference builds them (float32 torch.linspace on the host, Mapper.py:108-110)
    xs = torch.linspace(b[0][0], b[0][1], W).to(dev)
    ys = torch.linspace(b[1][0], b[1][1], H).to(dev)
    zs = torch.linspace(b[2][0], b[2][1], D).to(dev)
    pose = torch.as_tensor(c2w).detach().to("cpu", torch.float32).contiguous()
    c2w_host = (C.c_float * 16)(*pose.reshape(-1).tolist())
    dep = torch.as_tensor(depth).to(dev, torch.float32).contiguous()
    mask = torch.empty(D * H * W, dtype=torch.uint8, device=dev)
    ws = torch.empty(L.nsb_frustum_mask_workspace(D * H * W), dtype=torch.uint8, device=dev)
    _lib.check(L.nsb_frustum_mask(c2w_host, _VP(xs.data_ptr()), _VP(ys.data_ptr()), _VP(zs.data_ptr()), D, H, W,
                                  _VP(dep.data_ptr()), dep.shape[0], dep.shape[1], float(renderer.fx), float(renderer.fy),
                                  float(renderer.cx), float(renderer.cy), _VP(mask.data_ptr()), _VP(ws.data_ptr()), ws.numel(), _stream()),
               "nsb_frustum_mask")
    return mask.view(D, H, W).bool()
