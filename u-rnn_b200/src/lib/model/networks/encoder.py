"""
encoder.py (urnn_b200) -- drop-in for the reference's multi-scale ConvGRU encoder (encoder.py:63-215).
Three stages, each: fused stem kernel (1x1 conv + LeakyReLU [+ AvgPool2]) -> fused ConvGRU cell kernels.
The *_wrapper attributes exist only to reproduce the reference's aliased state_dict keys.
"""
import torch
from torch import nn

from src.lib.model.networks.utils import make_layers


class _Alias(nn.Module):
    """Parameter alias: registers `module` a second time under `<name>_wrapper.module.`."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        raise RuntimeError("urnn_b200: checkpoint wrappers are state_dict aliases and are never called")


class ModuleWrapperIgnores2ndArg_cnn(_Alias):
    pass


class ModuleWrapperIgnores2ndArg_lstm(_Alias):
    pass


class ModuleWrapperIgnores2ndArg_gru(_Alias):
    pass


class Encoder(nn.Module):
    def __init__(self, clstm, subnets, rnns, use_checkpoint):
        super().__init__()
        if len(subnets) != len(rnns) or len(rnns) != 3:
            raise ValueError("Encoder expects 3 stem specs and 3 recurrent cells")
        if clstm:
            raise NotImplementedError("the reference defines no ConvLSTM cell either (SURVEY.md section 2: dead path)")
        self.blocks = len(subnets)
        self.use_checkpoint = use_checkpoint
        self.clstm = clstm
        self.stage1 = make_layers(subnets[0])
        self.stage2 = make_layers(subnets[1])
        self.stage3 = make_layers(subnets[2])
        self.rnn1, self.rnn2, self.rnn3 = rnns
        self.dummy_tensor = torch.ones(1, dtype=torch.float32, requires_grad=True)
        self.stage1_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.stage1)
        self.stage2_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.stage2)
        self.stage3_wrapper = ModuleWrapperIgnores2ndArg_cnn(self.stage3)
        self.rnn1_wrapper = ModuleWrapperIgnores2ndArg_gru(self.rnn1)
        self.rnn2_wrapper = ModuleWrapperIgnores2ndArg_gru(self.rnn2)
        self.rnn3_wrapper = ModuleWrapperIgnores2ndArg_gru(self.rnn3)

    def forward_by_stage(self, i, inputs, hidden_state, subnet, rnn):
        """inputs (S,B,C,H,W) -> (outputs (1,B,F,H',W'), state (B,F,H',W')).  As in the reference
        (encoder.py:180 calls the cell with seq_len=1) only frame 0 of the sequence is consumed."""
        frame = subnet(inputs[0])                      # (B,C',H',W')
        out = rnn(frame.unsqueeze(0), hidden_state)    # (1,B,F,H',W')
        return out, out[0]

    def forward(self, inputs, state_stages):
        states = []
        for i in range(1, self.blocks + 1):
            inputs, st = self.forward_by_stage(i, inputs, state_stages[i - 1],
                                               getattr(self, f"stage{i}"), getattr(self, f"rnn{i}"))
            states.append(st)
        return tuple(states)
