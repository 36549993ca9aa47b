"""
flood_head.py (urnn_b200) -- drop-in for the reference's dual-output head (head/flood_head.py:40-202).
forward() hands the 16-channel decoder features and all 19 head tensors to liburnn_b200 (urnn_head_fwd):
four recompute sweeps separated by the three LayerNorm statistic levels; output (S,B,2,H,W) =
[depth * (prob >= cls_thred), prob].
"""
import torch
import torch.nn as nn

from urnn_b200 import ops
from .network_blocks import BaseConv, finalConv
from src.lib.model.networks.encoder import _Alias


class ModuleWrapperIgnores2ndArg_cnn(_Alias):
    pass


class YOLOXHead(nn.Module):
    def __init__(self, cls_thred=0.5, in_channels=64, width=0.25, depthwise=False, use_checkpoint=True,
                 input_height=500, input_width=500):
        super().__init__()
        if depthwise:
            raise NotImplementedError("depthwise head blocks are not part of the shipped network")
        ch = int(in_channels * width)
        if ch != 16:
            raise NotImplementedError(f"urnn_b200 head kernels are specialised for width 16 (got {ch})")
        H, W = input_height, input_width
        self.acts = ["silu"] * 3 + ["sigmoid", "lrelu"]

        def block():
            return BaseConv(ch, ch, ksize=1, stride=1, act="silu", height=H, width=W)

        self.stems = block()
        self.cls_convs = nn.Sequential(block(), block())
        self.reg_convs = nn.Sequential(block(), block())
        self.cls_preds = finalConv(ch, 1, ksize=1, stride=1, act="sigmoid", norm="")
        self.reg_preds = finalConv(ch, 1, ksize=1, stride=1, act="lrelu", norm="")
        self.stems_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.stems)
        self.cls_convs_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.cls_convs)
        self.reg_convs_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.reg_convs)
        self.cls_preds_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.cls_preds)
        self.reg_preds_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.reg_preds)
        self.dummy_tensor = torch.ones(1, dtype=torch.float32, requires_grad=True)
        self.use_checkpoint = use_checkpoint
        self.cls_thred = cls_thred

    def param_dict(self):
        blocks = [self.stems, self.cls_convs[0], self.cls_convs[1], self.reg_convs[0], self.reg_convs[1]]
        return {"conv_w": [b.conv.weight for b in blocks],
                "ln_w": [b.ln.weight for b in blocks], "ln_b": [b.ln.bias for b in blocks],
                "cls_pred_w": self.cls_preds.conv.weight, "cls_pred_b": self.cls_preds.conv.bias,
                "reg_pred_w": self.reg_preds.conv.weight, "reg_pred_b": self.reg_preds.conv.bias}

    def forward(self, inputs):
        """inputs (S,B,16,H,W) -> (S,B,2,H,W)."""
        S, B, C, H, W = inputs.size()
        flat = inputs.reshape(S * B, C, H, W)
        p = self.param_dict()
        outs = [ops.head(flat[i], p, float(self.cls_thred)) for i in range(S * B)]
        out = outs[0].unsqueeze(0) if len(outs) == 1 else torch.stack(outs)
        return out.reshape(S, B, 2, H, W)

    def correction_depth(self, reg_output_t, cls_output_t, flood_thres=0.5):
        """Kept for API parity (flood_head.py:179-202); the fused kernel applies the same mask itself."""
        return reg_output_t * (cls_output_t >= flood_thres).float()
