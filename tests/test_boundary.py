"""The drop-in boundary without a GPU: parameter tree / state_dict layout / constructor parity with the
reference, the C ABI symbol table, and loud failure on CPU tensors."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch


def build_ed(H, W, C, **kw):
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params
    torch.manual_seed(0)
    enc, dec = get_network_params(False, H, W, input_channels=C, net_cfg=None, **kw)
    return ED(False, enc, dec, 0.5, False, input_height=H, input_width=W)


def test_state_dict_keys_match_reference(golden_dir):
    want = [ln.rsplit(" (", 1) for ln in open(os.path.join(golden_dir, "state_dict_keys.txt")).read().splitlines()]
    want = [(k, tuple(int(v) for v in re.findall(r"\d+", s))) for k, s in want]
    net = build_ed(32, 32, 9)
    got = [(k, tuple(v.shape)) for k, v in net.state_dict().items()]
    assert len(got) == 254
    assert got == want                      # same keys, same shapes, same order
    assert len({v.data_ptr() for v in net.state_dict().values()}) == 79


def test_seeded_init_reproduces_reference_weights(golden_dir):
    """Same construction order + torch.manual_seed(0) => bit-identical parameters to the reference's."""
    z = np.load(os.path.join(golden_dir, "ed_32x32_c9.npz"))
    net = build_ed(32, 32, 9)
    sd = net.state_dict()
    n = 0
    for k in z.files:
        if k.startswith("w."):
            np.testing.assert_array_equal(sd[k[2:]].numpy(), z[k]); n += 1
    assert n == 79
    z2 = np.load(os.path.join(golden_dir, "ed_lite128.npz"))
    net2 = build_ed(128, 128, 9)
    fp = np.array([float(v.double().sum()) for v in net2.state_dict().values()])
    np.testing.assert_array_equal(fp, z2["w_fingerprint"])


def test_strict_load_of_reference_layout_checkpoint(golden_dir, tmp_path):
    """A checkpoint in the reference's format ({epoch,state_dict,optimizer}, optional 'module.' prefix,
    test.py:398-405) loads strictly, and ours loads back into the same key set."""
    net = build_ed(32, 32, 9)
    sd = {("module." + k): v.clone() for k, v in net.state_dict().items()}
    path = tmp_path / "checkpoint_3_0.123.pth.tar"
    torch.save({"epoch": 3, "state_dict": sd, "optimizer": {}}, path)
    info = torch.load(path, map_location="cpu")
    stripped = {k[7:] if k.startswith("module.") else k: v for k, v in info["state_dict"].items()}
    net2 = build_ed(32, 32, 9)
    net2.load_state_dict(stripped, strict=True)


def test_cell_constructor_contract():
    from src.lib.model.networks.ConvRNN import CGRU_cell
    c = CGRU_cell(False, (64, 64), 16, 1, 64, "encoder")
    assert c.conv1[0].weight.shape == (128, 80, 1, 1) and c.conv2[0].weight.shape == (64, 80, 1, 1)
    d = CGRU_cell(True, (32, 32), 96, 3, 64, "decoder")
    assert d.conv1[0].weight.shape == (128, 224, 3, 3) and d.padding == 1
    assert len(c.state_dict()) == 16
    with pytest.raises(ValueError):
        CGRU_cell(False, (8, 8), 4, 1, 16, "encoder")      # GroupNorm(F//32) needs F % 32 == 0 (as the reference)


def test_c_abi_exports_every_declared_symbol():
    from urnn_b200 import _capi
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "urnn_b200.h")).read()
    declared = set(re.findall(r"\b(urnn_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    loaded = _capi.load()
    assert loaded.urnn_abi_version() == 1
    assert loaded.urnn_launch_count() == 0          # nothing may have launched on a CPU box


def test_cpu_tensors_are_rejected_loudly():
    """No CPU fallback: the product path raises instead of computing on the host."""
    net = build_ed(16, 16, 9).eval()
    st = [torch.zeros(1, c, 16 // s, 16 // s) for c, s in ((64, 1), (96, 2), (96, 4), (96, 4), (96, 2), (64, 1))]
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA tensors"):
        net(torch.rand(1, 1, 9, 16, 16), *st)
    from src.lib.model.networks.ConvRNN import CGRU_cell
    cell = CGRU_cell(False, (16, 16), 16, 1, 64, "encoder")
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA tensors only"):
        cell(torch.rand(1, 1, 16, 16, 16), torch.zeros(1, 64, 16, 16))


def test_argument_validation_without_gpu():
    from urnn_b200 import _capi
    lib = _capi.load()
    d = _capi.CellDesc(16, 16, 16, 48, 1, 0, 0, 1e-5)        # F not a multiple of 32
    p = _capi.CellParams()
    rc = lib.urnn_cgru_fwd(ctypes.byref(d), ctypes.byref(p), None, None, None, None, None, 0, None)
    assert rc == -1 and b"multiple of 32" in lib.urnn_last_error()
    ed = _capi.EdDesc()
    ed.H, ed.W, ed.Cin = 30, 32, 9
    assert lib.urnn_ed_step_workspace_bytes(ctypes.byref(ed)) == 0
    assert b"multiples of 4" in lib.urnn_last_error()


@pytest.mark.skipif(not os.path.isdir("/root/reference/code"), reason="reference tree only exists in the build container")
def test_overlay_resolution_with_reference_tree():
    """With [u-rnn_b200, reference/code] on sys.path the hot-path modules are ours and the rest is the reference's."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.dont_write_bytecode=True; sys.path[:0]=[%r, '/root/reference/code'];"
        "import src.lib.model.networks.model as m, src.lib.model.networks.losses as l, src.lib.utils.net_config as n,"
        " src.lib.model.earlystopping as e, src.lib.model.networks.ConvRNN as c;"
        "print(m.__file__); print(c.__file__); print(l.__file__); print(n.__file__); print(e.__file__)"
    ) % os.path.join(root, "u-rnn_b200")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, check=True).stdout.split()
    assert "u-rnn_b200" in out[0] and "u-rnn_b200" in out[1]
    assert out[2].startswith("/root/reference") and out[3].startswith("/root/reference") and out[4].startswith("/root/reference")
