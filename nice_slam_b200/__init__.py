"""nice_slam_b200 -- B200-native render-and-backprop path for NICE-SLAM (drop-in for src/utils/Renderer.py).

Public surface:
    FusedRenderer   mirror of the reference's Renderer (render_batch_ray / eval_points / render_img)
    NICEDecoders    parameter container with the reference's state_dict keys
    to_channels_last, lib (ctypes handle of libnsb.so)
"""
from ._lib import lib, LIB_PATH                       # noqa: F401
from .decoders import NICEDecoders                     # noqa: F401
from .renderer import FusedRenderer, to_channels_last  # noqa: F401

__all__ = ["FusedRenderer", "NICEDecoders", "to_channels_last", "lib", "LIB_PATH"]
