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
