#!/usr/bin/env python
"""Run an UNMODIFIED script of the reference (main.py, test.py, ...) on top of the urnn_b200 drop-in modules.

    python tools/run_with_urnn_b200.py /path/to/U-RNN/code/test.py --exp_config configs/lite.yaml --timestamp T

sys.path gets [<this repo>/u-rnn_b200, <reference code dir>] in that order, so
`src.lib.model.networks.{ConvRNN,encoder,decoder,model,net_params,utils,head.*}` resolve to the B200 package while
everything the package does not provide (`src.lib.model.networks.losses`, `src.lib.model.earlystopping`,
`src.lib.utils.*`, `src.lib.dataset.*`, `config`, `configs/`) falls through to the reference tree (`src` and `src.lib`
are namespace packages there; our `model` / `networks` packages extend their `__path__`).  URNN_MATH=f16x3|fp32|bf16 selects
the arithmetic of the gate contractions: f16x3 (default; tcgen05 with fp16 hi+lo split operands, inside the fp32 tolerance),
fp32 (FFMA parity mode; required for training: the backward kernels are fp32), bf16 (round-1 single-pass mode, short horizons).
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(os.path.dirname(HERE), "u-rnn_b200")


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    code_dir = os.path.dirname(script)
    sys.path[:0] = [PKG, code_dir]
    sys.argv = [script] + sys.argv[2:]
    os.chdir(code_dir)                      # the reference resolves configs/ relative to its code directory
    import urnn_b200
    urnn_b200.set_default_math(os.environ.get("URNN_MATH", "f16x3"))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
