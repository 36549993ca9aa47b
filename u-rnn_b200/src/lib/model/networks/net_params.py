"""
net_params.py (urnn_b200) -- the plug-in seam: builds the stem specs and the CGRU_cell objects that are
injected into ED, with the reference's signature and return structure (net_params.py:5-141).
"""
from collections import OrderedDict

from src.lib.model.networks.ConvRNN import CGRU_cell

DEFAULT_NET_CFG = {   # configs/network.yaml of the reference (published architecture)
    "encoder": {"conv_out_channels": [16, 64, 96], "gru_channels": [64, 96, 96],
                "downsample_factors": [1, 2, 2], "filter_size": 1},
    "decoder": {"gru_channels": [96, 96, 64], "conv_out_channels": [96, 96, 16],
                "upsample_factors": [2, 2, 1], "filter_size": 1},
    "head": {"in_channels": 64, "width": 0.25},
}


def get_network_params(use_checkpoint, input_height=500, input_width=500, input_channels=63, net_cfg=None,
                       math=None):
    """Returns ([encoder stem specs, encoder cells], [decoder stem specs, decoder cells]).
    `math` is an additive option forwarded to every cell ("fp32" | "f16x3" | "bf16" | None)."""
    cfg = net_cfg if net_cfg is not None else DEFAULT_NET_CFG
    enc, dec = cfg["encoder"], cfg["decoder"]
    e_conv, e_gru, down, e_k = enc["conv_out_channels"], enc["gru_channels"], enc["downsample_factors"], enc["filter_size"]
    d_gru, d_conv, up, d_k = dec["gru_channels"], dec["conv_out_channels"], dec["upsample_factors"], dec["filter_size"]
    n = len(e_gru)

    scale, scales = 1, []
    for f in down:
        scale *= f
        scales.append(scale)
    size_at = [(input_height // s, input_width // s) for s in scales]

    # encoder: stem k maps (input | previous GRU state) -> e_conv[k], optional pool, then a ConvGRU
    stem_in = [input_channels] + list(e_gru[:-1])
    enc_specs = []
    for k in range(n):
        spec = OrderedDict()
        spec[f"conv{k + 1}_leaky_1"] = [stem_in[k], e_conv[k], e_k, 1, 0]
        if down[k] > 1:
            spec["avgpool"] = [down[k], down[k], 0]
        enc_specs.append(spec)
    enc_cells = [CGRU_cell(use_checkpoint=use_checkpoint, shape=size_at[k], input_channels=e_conv[k],
                           filter_size=e_k, num_features=e_gru[k], module="encoder", math=math)
                 for k in range(n)]

    # decoder: index 0 is the deepest scale; its stem consumes the matching GRU state
    dec_specs = []
    for k in range(n):
        spec = OrderedDict()
        cin = e_gru[n - 1 - k]
        if up[k] > 1:
            spec[f"deconv{k + 1}_leaky_1"] = [cin, d_conv[k], d_k + 1, up[k], 0]
        else:
            spec[f"conv{k + 1}_leaky_1"] = [cin, d_conv[k], d_k, 1, 0]
        dec_specs.append(spec)
    x_ch = [d_conv[0]] + [d_conv[k - 1] for k in range(1, n)]
    dec_cells = [CGRU_cell(use_checkpoint=use_checkpoint, shape=size_at[n - 1 - k], input_channels=x_ch[k],
                           filter_size=d_k, num_features=d_gru[k], module="decoder", math=math)
                 for k in range(n)]
    return [enc_specs, enc_cells], [dec_specs, dec_cells]
