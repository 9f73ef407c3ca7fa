"""Test-side harness that imports the UNMODIFIED reference (cvg/nice-slam) from /root/reference.

Only used in the build container (where /root/reference exists) to
  * pin the oracle (oracle/torch_port.py, oracle/nsb_oracle.c) against the real reference, and
  * generate the golden fixtures under tests/golden/ (tests/make_golden.py).
Nothing here is imported by the product path, bench.py or the `-m gpu` tests.

Shims -- none of them edits the reference:
  1. common.quad2rotation uses `.to(quad.get_device())` (src/common.py:150) which is -1 on CPU.
  2. NICE.forward builds f'cuda:{p.get_device()}' (src/conv_onet/models/decoder.py:316) -> 'cuda:-1' on CPU.
  3. colorama / matplotlib / mathutils / open3d / skimage / trimesh are not installed -> stub modules.
  4. np.bool / np.float aliases (src/Mapper.py:114).
  5. Tracker / Mapper are created with object.__new__ + attribute injection (their __init__ needs datasets).
"""
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF_ROOT = os.environ.get("NSB_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "src"))


_IMPORTED = {}


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference():
    """Import the reference's hot-path modules with the CPU shims installed. Idempotent."""
    if _IMPORTED:
        return SimpleNamespace(**_IMPORTED)
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)

    # (3) stubs for missing third-party modules
    class _Any:
        def __getattr__(self, k):
            return ""
    _stub("colorama", Fore=_Any(), Style=_Any())
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    for n in ("open3d", "trimesh", "skimage", "skimage.measure", "packaging_stub"):
        _stub(n)
    sys.modules["skimage"].measure = sys.modules["skimage.measure"]

    class _Matrix:
        def __init__(self, R):
            self.R = np.asarray(R, dtype=np.float64)

        def to_quaternion(self):
            from scipy.spatial.transform import Rotation
            x, y, z, w = Rotation.from_matrix(self.R).as_quat()
            return np.array([w, x, y, z])
    _stub("mathutils", Matrix=_Matrix)

    # (4) numpy aliases removed in numpy>=1.24
    if not hasattr(np, "bool"):
        np.bool = bool
    if not hasattr(np, "float"):
        np.float = float

    # (2) 'cuda:-1' -> 'cpu' for Tensor.to / Module.to
    _orig_to = torch.Tensor.to

    def _to(self, *a, **k):
        if a and isinstance(a[0], str) and a[0] == "cuda:-1":
            a = ("cpu",) + tuple(a[1:])
        if a and isinstance(a[0], int) and a[0] == -1:   # (1) .to(quad.get_device())
            a = ("cpu",) + tuple(a[1:])
        if k.get("device", None) == "cuda:-1":
            k["device"] = "cpu"
        return _orig_to(self, *a, **k)
    torch.Tensor.to = _to

    from src import common, config  # noqa
    from src.utils import Renderer as renderer_mod
    from src.conv_onet.models import decoder as decoder_mod
    from src import Tracker as tracker_mod
    from src import Mapper as mapper_mod
    from src import NICE_SLAM as slam_mod

    _IMPORTED.update(common=common, config=config, Renderer=renderer_mod.Renderer,
                     decoder=decoder_mod, Tracker=tracker_mod.Tracker, Mapper=mapper_mod.Mapper,
                     NICE_SLAM=slam_mod.NICE_SLAM)
    return SimpleNamespace(**_IMPORTED)


def load_cfg(rel_yaml="configs/Replica/room0.yaml"):
    ref = import_reference()
    cwd = os.getcwd()
    os.chdir(REF_ROOT)            # inherit_from paths are relative to the reference root
    try:
        cfg = ref.config.load_config(rel_yaml, "configs/nice_slam.yaml")
    finally:
        os.chdir(cwd)
    return cfg


def build_slam(cfg, seed=0, grid_scale=None, device="cpu"):
    """Run the reference's own NICE_SLAM.load_bound / load_pretrain / grid_init / update_cam on a
    bare namespace (NICE_SLAM.__init__ needs datasets).  Returns a namespace with bound, shared_c,
    shared_decoders, H,W,fx,fy,cx,cy  -- exactly what Renderer/Tracker/Mapper read from `slam`."""
    ref = import_reference()
    S = ref.NICE_SLAM
    torch.manual_seed(seed)
    np.random.seed(seed)
    slam = SimpleNamespace()
    slam.cfg = cfg
    slam.nice = True
    slam.coarse = cfg["coarse"]
    slam.occupancy = cfg["occupancy"]
    slam.low_gpu_mem = cfg["low_gpu_mem"]
    slam.verbose = False
    slam.coarse_bound_enlarge = cfg["model"]["coarse_bound_enlarge"]
    slam.scale = cfg["scale"]
    slam.H, slam.W, slam.fx, slam.fy, slam.cx, slam.cy = (cfg["cam"][k] for k in ("H", "W", "fx", "fy", "cx", "cy"))
    S.update_cam(slam)
    slam.shared_decoders = ref.config.get_model(cfg, nice=True)
    S.load_bound(slam, cfg)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    _orig_load = torch.load
    torch.load = lambda f, **kw: _orig_load(f, map_location="cpu", weights_only=False)
    try:
        cfg2 = dict(cfg)
        cfg2["mapping"] = dict(cfg["mapping"])
        cfg2["mapping"]["device"] = "cpu"
        S.load_pretrain(slam, cfg2)
    finally:
        torch.load = _orig_load
        os.chdir(cwd)
    S.grid_init(slam, cfg)
    if grid_scale is not None:      # "trained-like" variant: non-trivial occupancies
        for k, v in slam.shared_c.items():
            v.mul_(grid_scale.get(k, 1.0))
    slam.output = "Demo"
    slam.mesher = None
    slam.logger = None
    return slam


def make_renderer(cfg, slam):
    ref = import_reference()
    args = SimpleNamespace(nice=True)
    return ref.Renderer(cfg, args, slam)


def make_tracker(cfg, slam, renderer, device="cpu"):
    ref = import_reference()
    t = object.__new__(ref.Tracker)
    t.cfg = cfg
    t.device = device
    t.nice = True
    t.bound = slam.bound
    t.renderer = renderer
    t.H, t.W, t.fx, t.fy, t.cx, t.cy = slam.H, slam.W, slam.fx, slam.fy, slam.cx, slam.cy
    tr = cfg["tracking"]
    t.ignore_edge_W, t.ignore_edge_H = tr["ignore_edge_W"], tr["ignore_edge_H"]
    t.handle_dynamic = tr["handle_dynamic"]
    t.use_color_in_tracking = tr["use_color_in_tracking"]
    t.w_color_loss = tr["w_color_loss"]
    t.c = {k: v.clone() for k, v in slam.shared_c.items()}
    t.decoders = slam.shared_decoders
    return t


def make_mapper(cfg, slam, renderer, coarse_mapper=False, device="cpu", BA=False):
    ref = import_reference()
    m = object.__new__(ref.Mapper)
    m.cfg = cfg
    m.coarse_mapper = coarse_mapper
    m.nice = True
    m.c = slam.shared_c
    m.bound = slam.bound
    m.output = "Demo"
    m.verbose = False
    m.renderer = renderer
    m.decoders = slam.shared_decoders
    m.device = device
    mp = cfg["mapping"]
    m.fix_fine, m.fix_color = mp["fix_fine"], mp["fix_color"]
    m.BA, m.BA_cam_lr = BA, mp["BA_cam_lr"]
    m.mapping_pixels = mp["pixels"]
    m.w_color_loss = mp["w_color_loss"]
    m.fine_iter_ratio, m.middle_iter_ratio = mp["fine_iter_ratio"], mp["middle_iter_ratio"]
    m.mapping_window_size = mp["mapping_window_size"]
    m.frustum_feature_selection = mp["frustum_feature_selection"]
    m.keyframe_selection_method = "global" if coarse_mapper else mp["keyframe_selection_method"]
    m.save_selected_keyframes_info = False
    m.no_vis_on_first_frame = mp["no_vis_on_first_frame"]
    m.occupancy = cfg["occupancy"]
    m.keyframe_dict, m.keyframe_list = [], []
    m.H, m.W, m.fx, m.fy, m.cx, m.cy = slam.H, slam.W, slam.fx, slam.fy, slam.cx, slam.cy
    return m
