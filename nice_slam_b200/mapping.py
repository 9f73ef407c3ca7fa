"""The joint-iteration loop of Mapper.optimize_map for NICE-SLAM without bundle adjustment (src/Mapper.py:296-333 set-up, :389-520 loop),
assembled from the library's pieces: frustum mask on the device -> voxel slot tables -> per iteration one fused mapping iteration
(compact voxel gradients + colour-decoder gradients, steps.IterationContext) + fused Adam on the selected voxels (in place on the shared
grids) and on the colour decoder (optim.FusedMapperAdam).  Ray sampling stays with the caller (the reference's get_samples, torch RNG)."""
import torch

from ._lib import STAGE_DECODERS
from .masked import MaskedVoxels, frustum_voxel_mask
from .optim import FusedMapperAdam
from .steps import IterationContext

GRID_OF = {"middle": "grid_middle", "fine": "grid_fine", "color": "grid_color", "coarse": "grid_coarse"}


class FusedMappingLoop:
    def __init__(self, renderer, c, decoders, c2w, depth, keys=("grid_middle", "grid_fine", "grid_color"), w_color=0.2):
        """c: the shared grids dict (updated in place); c2w / depth: pose and sensor depth of the current frame (frustum selection)."""
        self.r, self.c, self.dec, self.w_color = renderer, c, decoders, w_color
        self.masked = {k: MaskedVoxels(c[k], frustum_voxel_mask(renderer, c2w, k, c[k], depth)) for k in keys}
        self.adam = FusedMapperAdam()
        self._ctx = {}

    def _context(self, n, stage, device):
        key = (n, stage)
        if key not in self._ctx:
            grids = tuple(GRID_OF[l] for l in STAGE_DECODERS[stage] if GRID_OF[l] in self.masked)
            self._ctx[key] = IterationContext(self.r, n, stage, device, kind="map", grad_grids=grids,
                                              grad_decoders=("color",) if stage == "color" else (), masked=self.masked)
        return self._ctx[key]

    def iteration(self, stage, rays_o, rays_d, gt_depth, gt_color, lr):
        """One joint iteration at `stage`; lr = dict(decoders=, middle=, fine=, color=) as cfg['mapping']['stage'][stage] x lr_factor
        (Mapper.py:412-416).  Returns the loss tensor (device, float64)."""
        ctx = self._context(rays_o.shape[0], stage, rays_o.device)
        loss = ctx.run(self.c, self.dec, rays_o, rays_d, gt_depth, gt_color, w_color=self.w_color)
        for key in ctx.grad_grids:                              # param groups whose parameters received a gradient (Adam skips the others)
            self.adam.step_voxels(key, self.c[key], self.masked[key], ctx.d_grid[key], lr[key[5:]])
        if stage == "color":
            self.adam.step_decoder("color", self.dec, ctx.d_flat["color"], lr["decoders"], renderer=self.r)
        return loss
