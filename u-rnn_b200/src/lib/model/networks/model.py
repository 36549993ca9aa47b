"""
model.py (urnn_b200) -- drop-in for the reference's ED (model.py:22-121): Encoder -> Decoder -> head, one time
step per call with the six recurrent states passed explicitly.  Same constructor, forward signature,
return order and state_dict layout (254 keys).

Two execution routes, numerically identical because they enqueue the same kernels:
  * no gradients needed  -> ONE library call for the whole step (urnn_ed_step_fwd);
  * autograd             -> the per-module ops (cell / stem / head autograd Functions with CUDA backward).
"""
import torch
from torch import nn

from urnn_b200 import ops
from urnn_b200._capi import EdParams
from src.lib.model.networks.decoder import Decoder
from src.lib.model.networks.encoder import Encoder
from src.lib.model.networks.head.flood_head import YOLOXHead


class ED(nn.Module):
    def __init__(self, clstm_flag, encoder_params, decoder_params, cls_thred=0.5, use_checkpoint=True,
                 input_height=500, input_width=500):
        super().__init__()
        self.encoder = Encoder(clstm_flag, encoder_params[0], encoder_params[1], use_checkpoint=use_checkpoint)
        self.decoder = Decoder(clstm_flag, decoder_params[0], decoder_params[1], use_checkpoint=use_checkpoint)
        self.head = YOLOXHead(cls_thred, use_checkpoint=use_checkpoint,
                              input_height=input_height, input_width=input_width)
        self._plan = None
        for stems, cells in zip(self._stems(), self._cells()):      # stems follow the arithmetic mode of the cells
            for stem, cell in zip(stems, cells):
                stem.math = cell.math

    # ---- fused whole-step route --------------------------------------------------------------
    def _cells(self):
        return ([self.encoder.rnn1, self.encoder.rnn2, self.encoder.rnn3],
                [self.decoder.rnn3, self.decoder.rnn2, self.decoder.rnn1])

    def _stems(self):
        return ([self.encoder.stage1, self.encoder.stage2, self.encoder.stage3],
                [self.decoder.stage3, self.decoder.stage2, self.decoder.stage1])

    def ed_desc(self, H, W, Cin, math=None):
        enc_cells, dec_cells = self._cells()
        enc_stems, dec_stems = self._stems()
        math = math or enc_cells[0].math
        return ops.make_ed_desc(H, W, Cin,
                                [s._conv().out_channels for s in enc_stems], [c.num_features for c in enc_cells],
                                [c.num_features for c in dec_cells], [s._conv().out_channels for s in dec_stems],
                                float(self.head.cls_thred), math, ksize=enc_cells[0].filter_size)

    def ed_params(self):
        """ctypes parameter block over the CURRENT parameter storages (rebuild after .to()/load)."""
        enc_cells, dec_cells = self._cells()
        enc_stems, dec_stems = self._stems()
        p = EdParams()
        for k in range(3):
            p.enc_stem_w[k] = enc_stems[k]._conv().weight.data_ptr()
            p.enc_stem_b[k] = enc_stems[k]._conv().bias.data_ptr()
            p.dec_stem_w[k] = dec_stems[k]._conv().weight.data_ptr()
            p.dec_stem_b[k] = dec_stems[k]._conv().bias.data_ptr()
            p.enc_cell[k] = ops.cell_params_struct(*[t.detach() for t in enc_cells[k].param_list()])
            p.dec_cell[k] = ops.cell_params_struct(*[t.detach() for t in dec_cells[k].param_list()])
        p.head = ops.head_params_struct({k: ([t.detach() for t in v] if isinstance(v, list) else v.detach())
                                         for k, v in self.head.param_dict().items()})
        return p

    def _fused_step(self, x, states):
        dev = x.device
        Cin, H, W = x.shape
        math = self._cells()[0][0].math or ops.get_default_math()
        key = (H, W, Cin, dev, math)
        if self._plan is None or self._plan[0] != key:
            # everything that depends only on the geometry is built once: descriptor, workspace, expected shapes
            desc = self.ed_desc(H, W, Cin)
            stem1 = self._stems()[0][0]._conv()
            if Cin != stem1.in_channels:
                raise ValueError(f"input_t has {Cin} channels, encoder.stage1 expects {stem1.in_channels}")
            ln = self.head.param_dict()["ln_w"][0]
            if tuple(ln.shape) != (16, H, W):
                raise ValueError(f"this model's head LayerNorm is built for {tuple(ln.shape[1:])} grids, input_t is {H}x{W} "
                                 "(checkpoints are resolution-specific, network_blocks.py:93-94)")
            from urnn_b200.runner import state_shapes
            enc_cells, dec_cells = self._cells()
            shapes = state_shapes(H, W, [c.num_features for c in enc_cells], [c.num_features for c in dec_cells])
            ws = torch.empty(ops.ed_workspace_bytes(desc), dtype=torch.uint8, device=dev)
            self._plan = (key, desc, ws, shapes)
        _, desc, ws, shapes = self._plan
        for p in self.parameters():
            if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
                raise RuntimeError("urnn_b200: ED parameters must be contiguous float32 CUDA tensors")
        sin = [ops._chk(s[0], f"state {i}", shapes[i]) for i, s in enumerate(states)]
        sout = [torch.empty_like(s) for s in sin]
        out = torch.empty((2, H, W), dtype=torch.float32, device=dev)
        ops.ed_step_fwd(desc, self.ed_params(), ops._chk(x, "input_t"), sin, sout, out, ws)
        return out, sout

    # ---- reference contract ------------------------------------------------------------------
    def forward(self, input_t,
                prev_encoder_state1, prev_encoder_state2, prev_encoder_state3,
                prev_decoder_state1, prev_decoder_state2, prev_decoder_state3):
        """input_t (B,S,C,H,W); states as produced by the reference's initialize_states.  Returns
        (depth (S,B,H,W), e1, e2, e3, d(1/4), d(1/2), d(1x)) -- model.py:115-121 order."""
        enc_prev = [prev_encoder_state1, prev_encoder_state2, prev_encoder_state3]
        dec_prev = [prev_decoder_state1, prev_decoder_state2, prev_decoder_state3]
        states = enc_prev + dec_prev
        needs_grad = torch.is_grad_enabled() and (
            input_t.requires_grad or any(s is not None and s.requires_grad for s in states)
            or any(p.requires_grad for p in self.parameters()))
        if (not needs_grad and input_t.size(0) == 1 and all(s is not None for s in states)):
            out, sout = self._fused_step(input_t[0, 0], states)
            return (out[0][None, None], *[s.unsqueeze(0) for s in sout])

        seq_first = input_t.permute(1, 0, 2, 3, 4)
        enc_state = self.encoder(seq_first, enc_prev)
        feat, dec_state = self.decoder(enc_state, dec_prev)
        out = self.head(feat)                       # (S,B,2,H,W)
        return (out[:, :, 0], *enc_state, *dec_state)
