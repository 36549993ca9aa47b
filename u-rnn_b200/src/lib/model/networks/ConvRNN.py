"""
ConvRNN.py (urnn_b200) -- drop-in for the reference's src/lib/model/networks/ConvRNN.py.

Same class names, constructor signature, parameter tree (hence the same 16 state_dict keys per cell,
SURVEY.md appendix A) and forward contract as the reference (ConvRNN.py:47-194), but forward() does not
run nn.Conv2d / nn.GroupNorm: the Conv2d/GroupNorm modules below only OWN the parameters, and the step is
executed by liburnn_b200's fused sm_100a kernels (urnn_cgru_fwd / urnn_cgru_bwd in include/urnn_b200.h).
"""
import torch
import torch.nn as nn

from urnn_b200 import ops
from urnn_b200._capi import URNN_CELL_DECODER, URNN_CELL_ENCODER


class ModuleWrapperIgnores2ndArg(nn.Module):
    """Kept only because it re-registers the wrapped module and thereby creates the aliased
    `*_module_wrapper.module.*` state_dict keys that reference checkpoints contain (ConvRNN.py:30-44)."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, x, dummy_arg=None):
        raise RuntimeError("urnn_b200: the checkpoint wrapper is a parameter alias only and is never called")


class CGRU_cell(nn.Module):
    """(Skip-)ConvGRU cell; interface of the reference's CGRU_cell (ConvRNN.py:47-109).

    module="encoder": hidden_state is h (B,F,H,W).  module="decoder": hidden_state is cat(e, d) (B,2F,H,W),
    only d is updated.  `use_checkpoint` is accepted for signature parity: the CUDA backward always
    recomputes the forward intermediates, which is what reentrant checkpointing does in the reference.
    `math` (additive, optional): "fp32" | "f16x3" | "bf16"; None follows urnn_b200.set_default_math().
    """

    def __init__(self, use_checkpoint, shape, input_channels, filter_size, num_features, module, math=None):
        super().__init__()
        if module not in ("encoder", "decoder"):
            raise ValueError(f"module must be 'encoder' or 'decoder', got {module!r}")
        if int(num_features) % 32 != 0:
            # F < 32 fails in the reference too (GroupNorm(F//32 = 0 groups)); other non-multiples would need
            # GroupNorm groups wider than 32 channels, which the fused kernels do not implement
            raise ValueError(f"CGRU_cell: num_features={num_features} must be a multiple of 32")
        if math is not None and math not in ("fp32", "f16x3", "bf16"):
            raise ValueError(f"CGRU_cell: unknown math mode {math!r} (fp32 | f16x3 | bf16)")
        self.shape = shape
        self.input_channels = int(input_channels)
        self.filter_size = filter_size
        self.num_features = int(num_features)
        self.padding = (filter_size - 1) // 2
        self.module = module
        self.math = math
        hidden_ch = self.num_features * (2 if module == "decoder" else 1)
        cin = self.input_channels + hidden_ch
        # parameter owners; creation order = the reference's, so torch.manual_seed reproduces its init
        self.conv1 = nn.Sequential(
            nn.Conv2d(cin, 2 * self.num_features, filter_size, 1, self.padding),
            nn.GroupNorm(2 * self.num_features // 32, 2 * self.num_features))
        self.conv2 = nn.Sequential(
            nn.Conv2d(cin, self.num_features, filter_size, 1, self.padding),
            nn.GroupNorm(self.num_features // 32, self.num_features))
        self.use_checkpoint = use_checkpoint
        self.dummy_tensor = torch.ones(1, dtype=torch.float32, requires_grad=True)
        self.conv1_module_wrapper = ModuleWrapperIgnores2ndArg(self.conv1)
        self.conv2_module_wrapper = ModuleWrapperIgnores2ndArg(self.conv2)

    # -- helpers -------------------------------------------------------------------------------
    def param_list(self):
        return [self.conv1[0].weight, self.conv1[0].bias, self.conv1[1].weight, self.conv1[1].bias,
                self.conv2[0].weight, self.conv2[0].bias, self.conv2[1].weight, self.conv2[1].bias]

    def step(self, x, e, h):
        """One step on unbatched maps: x (Cx,H,W) | None, e (F,H,W) | None, h (F,H,W) -> (F,H,W)."""
        variant = URNN_CELL_DECODER if self.module == "decoder" else URNN_CELL_ENCODER
        return ops.cgru_cell(x, e, h, self.param_list(), self.filter_size, variant, self.math)

    # -- reference contract --------------------------------------------------------------------
    def forward(self, inputs=None, hidden_state=None, seq_len=1):
        F = self.num_features
        dev = self.conv1[0].weight.device
        if hidden_state is None:
            if inputs is None:
                raise ValueError("CGRU_cell: inputs and hidden_state cannot both be None")
            hidden_state = torch.zeros(inputs.size(1), F * (2 if self.module == "decoder" else 1),
                                       self.shape[0], self.shape[1], device=dev)
        if self.module == "decoder" and seq_len != 1:
            # the reference feeds its F-channel output back as the 2F-channel hidden state and dies in
            # torch.cat on the second iteration (ConvRNN.py:153,192); refuse up front instead
            raise ValueError("CGRU_cell(decoder): seq_len must be 1")
        B = hidden_state.size(0)
        outs = []
        state = hidden_state
        for t in range(seq_len):
            per_sample = []
            for b in range(B):
                x = None if inputs is None else inputs[t, b]
                if self.module == "decoder":
                    e, h = state[b, :F], state[b, F:]
                else:
                    e, h = None, state[b]
                per_sample.append(self.step(x, e, h))
            state = per_sample[0].unsqueeze(0) if B == 1 else torch.stack(per_sample)
            outs.append(state)
        return outs[0].unsqueeze(0) if seq_len == 1 else torch.stack(outs)
