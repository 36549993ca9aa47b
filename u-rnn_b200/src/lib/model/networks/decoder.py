"""
decoder.py (urnn_b200) -- drop-in for the reference's Skip-ConvGRU decoder (decoder.py:61-217).
Stages run deepest first; each = fused Skip-ConvGRU kernels (encoder state read in place, no torch.cat)
-> fused stem kernel (ConvTranspose2d(k2,s2)+LeakyReLU, or the final 1x1 conv+LeakyReLU).
"""
import torch
from torch import nn

from src.lib.model.networks.utils import make_layers
from src.lib.model.networks.encoder import _Alias


class ModuleWrapperIgnores2ndArg_cnn(_Alias):
    pass


class ModuleWrapperIgnores2ndArg_gru(_Alias):
    pass


class Decoder(nn.Module):
    def __init__(self, clstm, subnets, rnns, use_checkpoint):
        super().__init__()
        if len(subnets) != len(rnns) or len(rnns) != 3:
            raise ValueError("Decoder expects 3 stem specs and 3 recurrent cells")
        self.blocks = len(subnets)
        # index 0 of subnets / rnns is the deepest (1/4 resolution) stage
        self.stage3 = make_layers(subnets[0])
        self.stage2 = make_layers(subnets[1])
        self.stage1 = make_layers(subnets[2])
        self.rnn3, self.rnn2, self.rnn1 = rnns
        self.clstm = clstm
        self.use_checkpoint = use_checkpoint
        self.dummy_tensor = torch.ones(1, dtype=torch.float32, requires_grad=True)
        self.stage1_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.stage1)
        self.stage2_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.stage2)
        self.stage3_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.stage3)
        self.rnn1_wrapper = ModuleWrapperIgnores2ndArg_gru(self.rnn1)
        self.rnn2_wrapper = ModuleWrapperIgnores2ndArg_gru(self.rnn2)
        self.rnn3_wrapper = ModuleWrapperIgnores2ndArg_gru(self.rnn3)

    def forward_by_stage(self, i, inputs, encoder_states, decoder_states=None):
        """inputs (1,B,C,H,W) | None; encoder_states / decoder_states (B,F,H,W).
        Returns (stem output (1,B,C',H'',W''), new decoder state (B,F,H,W))."""
        rnn, stem = getattr(self, f"rnn{i}"), getattr(self, f"stage{i}")
        if decoder_states is None:
            decoder_states = torch.zeros_like(encoder_states)
        B = encoder_states.size(0)
        new = [rnn.step(None if inputs is None else inputs[0, b], encoder_states[b], decoder_states[b])
               for b in range(B)]
        state = new[0].unsqueeze(0) if B == 1 else torch.stack(new)
        return stem(state).unsqueeze(0), state

    def forward(self, encoder_states, decoder_states):
        """encoder_states (e1,e2,e3); decoder_states [d(1/4), d(1/2), d(1x)] -> ((B,1,16,H,W), new states
        in the same deepest-first order)."""
        out, st = self.forward_by_stage(3, None, encoder_states[-1], decoder_states[0])
        states = [st]
        for i in (2, 1):
            out, st = self.forward_by_stage(i, out, encoder_states[i - 1], decoder_states[self.blocks - i])
            states.append(st)
        return out.transpose(0, 1), tuple(states)
