"""make_ref.py -- vendors the UNMODIFIED reference modules of the hot path into oracle/_ref/ (git-ignored; travels to the GPU
box with gpurun like the built .so).  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python, so "building" it is a file copy, done where the sources lie: nothing under
/root/reference is modified and no reference source is committed to this repository.  The copy is what bench.py's
`--impl reference` arm and cpu_baseline leg (kind "reference") execute on the GPU box's host cores, and what
tests/test_gpu_callers.py runs the reference's own loops with.

Copied (paths relative to /root/reference/code):
  src/lib/model/networks/{ConvRNN,encoder,decoder,model,net_params,utils,losses}.py, head/{flood_head,network_blocks}.py
  src/lib/utils/{__init__,general,net_config,torch_utils,distributed_utils}.py, src/lib/dataset/Dynamic2DFlood.py,
  configs/network.yaml
The only adaptation happens at import time in the caller (a one-line `.cuda()` shim), never in the files.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("URNN_REFERENCE", "/root/reference/code")
DST = os.path.join(HERE, "_ref")
FILES = [
    "src/lib/model/networks/ConvRNN.py", "src/lib/model/networks/encoder.py", "src/lib/model/networks/decoder.py",
    "src/lib/model/networks/model.py", "src/lib/model/networks/net_params.py", "src/lib/model/networks/utils.py",
    "src/lib/model/networks/losses.py", "src/lib/model/networks/head/flood_head.py",
    "src/lib/model/networks/head/network_blocks.py", "src/lib/utils/general.py", "src/lib/utils/net_config.py",
    "src/lib/dataset/Dynamic2DFlood.py", "configs/network.yaml",
    "src/lib/utils/__init__.py", "src/lib/utils/torch_utils.py", "src/lib/utils/distributed_utils.py",   # imported by the package __init__
]


def available():
    return os.path.isdir(REF)


def present():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def make(verbose=False):
    """Copies the files (and creates the package __init__.py files); returns the destination or None if the reference
    tree is not mounted (GPU box: the prebuilt copy is used)."""
    if not available():
        return DST if present() else None
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            if verbose:
                print("copied", f, file=sys.stderr)
    d = os.path.join(DST, "src")
    for root, dirs, files in os.walk(d):
        init = os.path.join(root, "__init__.py")
        if not os.path.exists(init):
            src_init = os.path.join(REF, os.path.relpath(init, DST))
            if os.path.exists(src_init):
                shutil.copyfile(src_init, init)
            else:
                open(init, "w").close()
    return DST


if __name__ == "__main__":
    print(make(verbose=True))
