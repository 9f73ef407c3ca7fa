"""ORACLE (test infrastructure, NOT product code): ctypes front-end of oracle/nsb_oracle.c.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
Takes CPU torch tensors / numpy arrays shaped like the reference's (grids NCDHW or channels-last,
decoders as {name: tensor} dicts from oracle.torch_port.decoders_state) and returns numpy arrays.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
LEVELS = ("coarse", "middle", "fine", "color")
STAGES = {"coarse": 0, "middle": 1, "fine": 2, "color": 3}


class _Grid(C.Structure):
    _fields_ = [("data", C.c_void_p), ("D", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("sc", C.c_longlong), ("sd", C.c_longlong), ("sh", C.c_longlong), ("sw", C.c_longlong)]


class _Inputs(C.Structure):
    _fields_ = [("stage", C.c_int), ("n_rays", C.c_int), ("n_samples", C.c_int), ("n_surface", C.c_int),
                ("bound", C.c_double * 6), ("coarse_bound", C.c_double * 6),
                ("rays_o", C.c_void_p), ("rays_d", C.c_void_p), ("gt_depth", C.c_void_p),
                ("t_uniform", C.c_void_p), ("t_surface", C.c_void_p),
                ("grid", _Grid * 4), ("flat", C.c_void_p * 4)]


def build(force=False):
    so = os.path.join(_HERE, "libnsb_oracle.so")
    src = os.path.join(_HERE, "nsb_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libnsb_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.nsbo_flat_offset.restype = C.c_longlong
        _LIB.nsbo_flat_offset.argtypes = [C.c_int, C.c_int, C.c_int]
        _LIB.nsbo_flat_floats.restype = C.c_longlong
        _LIB.nsbo_flat_floats.argtypes = [C.c_int]
    return _LIB


_KINDS = [("embedder._B", 0, None), ("pts_linears.%d.weight", 1, 5), ("pts_linears.%d.bias", 2, 5),
          ("fc_c.%d.weight", 3, 5), ("fc_c.%d.bias", 4, 5), ("output_linear.weight", 5, None),
          ("output_linear.bias", 6, None)]


def flat_layout(level):
    """[(param name, offset, numel)] of the canonical flat order for decoder `level` (0..3)."""
    L = lib()
    out = []
    total = L.nsbo_flat_floats(level)
    offs = []
    for name, kind, n in _KINDS:
        if level == 0 and kind in (0, 3, 4):
            continue
        for i in (range(n) if n else [0]):
            offs.append((name % i if n else name, L.nsbo_flat_offset(level, kind, i)))
    offs.sort(key=lambda t: t[1])
    for j, (name, off) in enumerate(offs):
        end = offs[j + 1][1] if j + 1 < len(offs) else total
        out.append((name, off, end - off))
    return out


def flatten_decoder(level, state):
    flat = np.zeros(lib().nsbo_flat_floats(level), dtype=np.float32)
    for name, off, n in flat_layout(level):
        v = state[name].detach().cpu().numpy().astype(np.float32).reshape(-1)
        assert v.size == n, (name, v.size, n)
        flat[off:off + n] = v
    return flat


def unflatten_decoder(level, flat, like):
    return {name: torch.from_numpy(np.array(flat[off:off + n])).reshape(like[name].shape)
            for name, off, n in flat_layout(level)}


def _np(t, dtype):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


class Scene:
    """Holds numpy copies + the ctypes struct for one (grids, decoders, bound) scene."""

    def __init__(self, grids, dec, bound, coarse_enlarge=2.0, n_samples=32, n_surface=16):
        self.grids = {}
        for k, v in grids.items():
            a = v.detach().cpu()
            self.grids[k] = a            # keep strides as given (NCDHW or channels-last)
        self.flat = {}
        for lvl, name in enumerate(LEVELS):
            if name in dec:
                self.flat[lvl] = flatten_decoder(lvl, dec[name])
        self.bound = np.asarray(bound.detach().cpu().numpy(), dtype=np.float64).reshape(6)
        self.cbound = self.bound * coarse_enlarge
        self.n_samples, self.n_surface = n_samples, n_surface
        self.t_uniform = torch.linspace(0., 1., steps=n_samples).numpy().copy()
        self.t_surface = torch.linspace(0., 1., steps=max(n_surface, 1)).double().numpy().copy()[:n_surface]

    def _inputs(self, stage, rays_o, rays_d, gt_depth):
        inp = _Inputs()
        inp.stage = STAGES[stage]
        self._keep = [_np(rays_o, np.float32), _np(rays_d, np.float32), _np(gt_depth, np.float32)]
        inp.n_rays = self._keep[0].shape[0]
        inp.n_samples, inp.n_surface = self.n_samples, self.n_surface
        for i in range(6):
            inp.bound[i] = self.bound[i]
            inp.coarse_bound[i] = self.cbound[i]
        inp.rays_o = self._keep[0].ctypes.data
        inp.rays_d = self._keep[1].ctypes.data
        inp.gt_depth = self._keep[2].ctypes.data if self._keep[2] is not None else None
        inp.t_uniform = self.t_uniform.ctypes.data
        inp.t_surface = self.t_surface.ctypes.data if self.n_surface > 0 else None
        for lvl, name in enumerate(LEVELS):
            g = self.grids.get("grid_" + name)
            if g is not None:
                s = g.stride()
                inp.grid[lvl] = _Grid(g.data_ptr(), g.shape[2], g.shape[3], g.shape[4], s[1], s[2], s[3], s[4])
            if lvl in self.flat:
                inp.flat[lvl] = self.flat[lvl].ctypes.data
        return inp

    def n_per_ray(self, stage, gt_depth):
        return self.n_samples + (self.n_surface if (gt_depth is not None and stage != "coarse") else 0)

    def forward(self, stage, rays_o, rays_d, gt_depth):
        inp = self._inputs(stage, rays_o, rays_d, gt_depth)
        n, S = inp.n_rays, self.n_per_ray(stage, gt_depth)
        out = dict(depth=np.zeros(n), var=np.zeros(n), rgb=np.zeros((n, 3), np.float32),
                   z_vals=np.zeros((n, S)), raw=np.zeros((n, S, 4), np.float32),
                   weights=np.zeros((n, S), np.float32), corner_idx=np.zeros((n, S, 3), np.int32))
        rc = lib().nsbo_forward(C.byref(inp), *[C.c_void_p(out[k].ctypes.data) for k in
                                                 ("depth", "var", "rgb", "z_vals", "raw", "weights", "corner_idx")])
        assert rc == 0
        return out

    def backward(self, stage, rays_o, rays_d, gt_depth, g_depth, g_var=None, g_rgb=None,
                 grad_grids=(), grad_decoders=()):
        inp = self._inputs(stage, rays_o, rays_d, gt_depth)
        n = inp.n_rays
        gd, gv, gc = _np(g_depth, np.float64), _np(g_var, np.float64), _np(g_rgb, np.float32)
        out = dict(d_rays_o=np.zeros((n, 3), np.float32), d_rays_d=np.zeros((n, 3), np.float32))
        dgrid = (C.c_void_p * 4)()
        dflat = (C.c_void_p * 4)()
        self._dg = {}
        for lvl, name in enumerate(LEVELS):
            if "grid_" + name in grad_grids:
                g = self.grids["grid_" + name]
                buf = torch.zeros_like(g)       # preserves strides (dense layout)
                assert buf.stride() == g.stride()
                self._dg[name] = buf
                dgrid[lvl] = buf.data_ptr()
                out["d_grid_" + name] = buf
            if name in grad_decoders:
                buf = np.zeros_like(self.flat[lvl])
                dflat[lvl] = buf.ctypes.data
                out["d_flat_" + name] = buf
        rc = lib().nsbo_backward(C.byref(inp), C.c_void_p(gd.ctypes.data),
                                 C.c_void_p(gv.ctypes.data) if gv is not None else None,
                                 C.c_void_p(gc.ctypes.data) if gc is not None else None,
                                 C.c_void_p(out["d_rays_o"].ctypes.data), C.c_void_p(out["d_rays_d"].ctypes.data),
                                 dgrid, dflat)
        assert rc == 0
        return out
