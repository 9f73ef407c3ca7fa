"""Parameter containers for the NICE decoders.

The fused kernels consume the decoders' parameter tensors directly (by device pointer); they accept the
reference's own `NICE` module (src/conv_onet/models/decoder.py:277-342) unchanged.  `NICEDecoders` is a
stand-alone container with the SAME module tree / state_dict keys (embedder._B, pts_linears.i.{weight,bias},
fc_c.i.{weight,bias}, output_linear.{weight,bias}), so checkpoints are interchangeable, for use where the
reference package is not importable.  It holds parameters only; evaluation always goes through the CUDA path.
"""
import torch
import torch.nn as nn

from ._lib import LEVELS

HIDDEN, EMBED, C_DIM = 32, 93, 32


def _xavier_linear(n_in, n_out, gain):
    lin = nn.Linear(n_in, n_out)
    nn.init.xavier_uniform_(lin.weight, gain=gain)      # DenseLayer.reset_parameters, decoder.py:77-82
    nn.init.zeros_(lin.bias)
    return lin


class _Embedder(nn.Module):
    def __init__(self, scale=25.0):
        super().__init__()
        self._B = nn.Parameter(torch.randn(3, EMBED) * scale)     # decoder.py:21-22


class DecoderMLP(nn.Module):
    """Parameters of MLP (xyz=True; middle/fine/color) or MLP_no_xyz (xyz=False; coarse)."""

    def __init__(self, name, xyz=True, c_dim=C_DIM, color=False):
        super().__init__()
        self.name, self.xyz, self.c_dim, self.color = name, xyz, c_dim, color
        relu_gain = nn.init.calculate_gain("relu")
        if xyz:
            self.fc_c = nn.ModuleList([nn.Linear(c_dim, HIDDEN) for _ in range(5)])
            self.embedder = _Embedder()
            ins = [EMBED, HIDDEN, HIDDEN, HIDDEN + EMBED, HIDDEN]
        else:
            ins = [HIDDEN, HIDDEN, HIDDEN, HIDDEN + c_dim, HIDDEN]
        self.pts_linears = nn.ModuleList([_xavier_linear(i, HIDDEN, relu_gain) for i in ins])
        self.output_linear = _xavier_linear(HIDDEN, 4 if color else 1, 1.0)


class NICEDecoders(nn.Module):
    def __init__(self, coarse=True):
        super().__init__()
        if coarse:
            self.coarse_decoder = DecoderMLP("coarse", xyz=False)
        self.middle_decoder = DecoderMLP("middle")
        self.fine_decoder = DecoderMLP("fine", c_dim=2 * C_DIM)
        self.color_decoder = DecoderMLP("color", color=True)

    @classmethod
    def from_state(cls, state, device=None):
        """state: {'coarse'|'middle'|'fine'|'color': {param name: tensor}} (oracle.torch_port.decoders_state format)."""
        m = cls(coarse="coarse" in state)
        with torch.no_grad():
            for lvl, sd in state.items():
                getattr(m, lvl + "_decoder").load_state_dict({k: v.detach().clone() for k, v in sd.items()})
        return m.to(device) if device is not None else m


def decoder_module(decoders, level_name):
    sub = getattr(decoders, level_name + "_decoder", None)
    if sub is None:
        raise RuntimeError("decoders object has no %s_decoder" % level_name)
    return sub


def named_params(decoders, level_name):
    """{reference parameter name: tensor} of one decoder, whatever module class holds it."""
    sub = decoder_module(decoders, level_name)
    out = dict(sub.named_parameters())
    for k, v in sub.named_buffers():
        out.setdefault(k, v)
    if "embedder._B" not in out and hasattr(sub, "embedder") and hasattr(sub.embedder, "_B"):
        out["embedder._B"] = sub.embedder._B          # non-learnable variant keeps _B as a plain tensor
    return out
