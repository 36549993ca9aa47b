"""
network_blocks.py (urnn_b200) -- parameter owners for the head blocks with the reference's key names
(head/network_blocks.py:74-171): BaseConv = {conv.weight, ln.weight, ln.bias}; finalConv = {conv.weight,
conv.bias}.  The arithmetic lives in liburnn_b200's fused head kernels (urnn_head_fwd / urnn_head_bwd).
"""
import torch.nn as nn


class BaseConv(nn.Module):
    """1x1 conv (no bias) -> LayerNorm([C,H,W]) -> SiLU.  Owner only; YOLOXHead.forward runs the math."""

    def __init__(self, in_channels, out_channels, ksize, stride, groups=1, bias=False, act="silu",
                 height=None, width=None):
        super().__init__()
        if ksize != 1 or stride != 1 or groups != 1 or bias or act != "silu" or height is None or width is None:
            raise NotImplementedError("urnn_b200 head blocks are 1x1 conv (no bias) + LayerNorm([C,H,W]) + SiLU")
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0, groups=1, bias=False)
        self.ln = nn.LayerNorm([out_channels, height, width])
        self.act = nn.SiLU(inplace=True)

    def forward(self, x):
        raise RuntimeError("urnn_b200: BaseConv is evaluated inside the fused head kernel, not on its own")


class finalConv(nn.Module):
    """1x1 conv + bias -> activation (sigmoid for cls, LeakyReLU(0.2) for reg).  Owner only."""

    def __init__(self, in_channels, out_channels, ksize, stride, groups=1, bias=False, act="leaky", norm="gn"):
        super().__init__()
        if ksize != 1 or stride != 1 or norm != "" or out_channels != 1:
            raise NotImplementedError("urnn_b200 prediction convs are 1x1, single-channel, un-normalised")
        self.conv = nn.Conv2d(in_channels, out_channels, 1, 1, 0)
        self.norm = None
        self.act = nn.Sigmoid() if act == "sigmoid" else nn.LeakyReLU(0.2, inplace=True)

    def forward(self, x):
        raise RuntimeError("urnn_b200: finalConv is evaluated inside the fused head kernel, not on its own")
