"""Helpers for the GPU parity tests / smoke / bench: build a FusedRenderer on a scene_util scene."""
from types import SimpleNamespace

import torch

import scene_util as su
from nice_slam_b200.decoders import NICEDecoders
from nice_slam_b200.renderer import FusedRenderer

LV = {"coarse": ["coarse"], "middle": ["middle"], "fine": ["fine", "middle"], "color": ["fine", "color", "middle"]}


def make_cfg(sc, n_samples=None, n_surface=None):
    r = dict(sc["rendering"])
    if n_samples is not None:
        r["N_samples"] = n_samples
    if n_surface is not None:
        r["N_surface"] = n_surface
    return dict(rendering=r, scale=1, occupancy=True, model=dict(coarse_bound_enlarge=sc["coarse_bound_enlarge"]))


def make_renderer(sc, grids_cpu, dec_state, device="cuda", channels_last=True, n_samples=None, n_surface=None):
    """-> (renderer, c (dict of CUDA grids), decoders (NICEDecoders on CUDA))"""
    c = {k: v.to(device) for k, v in grids_cpu.items()}
    cam = sc["cam"]
    slam = SimpleNamespace(nice=True, bound=su.scene_bound(sc), shared_c=c, H=cam["H"], W=cam["W"], fx=cam["fx"], fy=cam["fy"],
                           cx=cam["cx"], cy=cam["cy"])
    r = FusedRenderer(make_cfg(sc, n_samples, n_surface), SimpleNamespace(nice=True), slam, convert_grids=channels_last)
    dec = NICEDecoders.from_state(dec_state, device)
    return r, slam.shared_c, dec


def rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def l2rel(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
