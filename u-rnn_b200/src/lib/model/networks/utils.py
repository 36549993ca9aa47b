"""
utils.py (urnn_b200) -- layer factory with the reference's naming rules (utils.py:73-125): the keys of the
spec dict become sub-module names and therefore state_dict keys ("conv1_leaky_1", "deconv1_leaky_1",
"avgpool").  The returned container only owns parameters; StageStem.forward dispatches the whole stem
(conv/deconv + LeakyReLU(0.2) [+ AvgPool2]) to one fused liburnn_b200 kernel.
"""
from collections import OrderedDict

import torch
from torch import nn

from urnn_b200 import ops


def get_normlization(name, num_features):
    if name == "":
        return None
    raise AttributeError(f"urnn_b200 stems support no normalisation layer (got {name!r}); the reference never uses one")


def get_activation(name="silu", inplace=True):
    table = {"silu": lambda: nn.SiLU(inplace=inplace), "relu": lambda: nn.ReLU(inplace=inplace),
             "lrelu": lambda: nn.LeakyReLU(0.2, inplace=inplace), "gelu": nn.GELU, "sigmoid": nn.Sigmoid}
    if name not in table:
        raise AttributeError(f"Unsupported act type: {name}")
    return table[name]()


class StageStem(nn.Sequential):
    """nn.Sequential (for the reference's key names) whose forward is a single fused CUDA op."""

    def __init__(self, layers, kind, pool):
        super().__init__(layers)
        self.kind = kind      # "conv" | "deconv"
        self.pool = pool      # 1 | 2 (AvgPool2d(2,2) after the activation)
        self.math = None      # "fp32" | "bf16" | None (library default); ED copies it from its cells

    def _conv(self):
        for m in self.children():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                return m
        raise RuntimeError("stem without a convolution")

    def forward(self, x):
        """x (B,C,H,W) -> (B,C',H',W'), batch handled sample by sample (B = 1 on the hot path)."""
        conv = self._conv()
        outs = []
        for b in range(x.size(0)):
            if self.kind == "deconv":
                outs.append(ops.deconv2x2_lrelu(x[b], conv.weight, conv.bias, math=self.math))
            else:
                outs.append(ops.conv1x1_lrelu(x[b], conv.weight, conv.bias, pool=self.pool, math=self.math))
        return outs[0].unsqueeze(0) if len(outs) == 1 else torch.stack(outs)


def make_layers(block, norm_name="", act="lrelu"):
    """block: OrderedDict name -> [cin, cout, k, stride, pad] (conv/deconv) or [k, stride, pad] (avgpool)."""
    if act != "lrelu" or norm_name != "":
        raise NotImplementedError("urnn_b200 stems implement conv/deconv + LeakyReLU(0.2) [+ AvgPool2] only")
    layers, kind, pool = [], None, 1
    for name, v in block.items():
        v = [int(t) for t in v]
        if "avgpool" in name:
            if v != [2, 2, 0] or kind != "conv":
                raise NotImplementedError(f"only AvgPool2d(2,2,0) after a conv stem is supported (got {v})")
            layers.append((name, nn.AvgPool2d(kernel_size=v[0], stride=v[1], padding=v[2])))
            pool = 2
        elif "deconv" in name:
            if v[2:] != [2, 2, 0] or kind is not None:
                raise NotImplementedError(f"only ConvTranspose2d(k=2,s=2,p=0) stems are supported (got {v})")
            layers.append((name, nn.ConvTranspose2d(v[0], v[1], kernel_size=v[2], stride=v[3], padding=v[4])))
            layers.append((act + "_" + name, get_activation(act)))
            kind = "deconv"
        elif "conv" in name:
            if v[2:] != [1, 1, 0] or kind is not None:
                raise NotImplementedError(
                    f"only 1x1/stride-1/pad-0 conv stems are supported (got {v}); the reference's encoder-decoder "
                    "breaks for filter_size > 1 as well (padding=0 shrinks the map before torch.cat)")
            layers.append((name, nn.Conv2d(v[0], v[1], kernel_size=v[2], stride=v[3], padding=v[4])))
            layers.append((act + "_" + name, get_activation(act)))
            kind = "conv"
        else:
            raise NotImplementedError(name)
    return StageStem(OrderedDict(layers), kind, pool)
