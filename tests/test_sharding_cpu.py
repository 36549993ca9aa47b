"""N>1 host logic on CPU with gloo (world_size 2): band partitioning, state_dict sharding, and the decomposition the
GPU path relies on -- a row-band-sharded encoder-decoder step whose ONLY communication is an all-reduce of the
normalisation (sum, sumsq, count) triples reproduces the unsharded result (oracle level, fp64)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_band_partition_and_state_dict_sharding():
    sys.path[:0] = [os.path.join(ROOT, "u-rnn_b200")]
    from urnn_b200 import dist as ud
    assert ud.band(32, 2, 0) == (0, 16) and ud.band(32, 2, 1) == (16, 16)
    with pytest.raises(ValueError):
        ud.band(36, 2, 0)                        # bands must be multiples of 4 rows
    t = torch.arange(2 * 8 * 4.).reshape(2, 8, 4)
    assert torch.equal(ud.shard_rows(t, 2, 1, scale=4), t[:, 4:8])
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params
    torch.manual_seed(0)
    enc, dec = get_network_params(False, 32, 16, input_channels=9)
    net = ED(False, enc, dec, 0.5, False, input_height=32, input_width=16)
    sd = ud.shard_state_dict(net.state_dict(), 2, 1)
    assert len(sd) == 254
    assert sd["head.stems.ln.weight"].shape == (16, 16, 16)
    assert sd["head.stems.ln.weight"].data_ptr() == sd["head.stems_wrapper.module.ln.weight"].data_ptr()
    enc2, dec2 = get_network_params(False, 16, 16, input_channels=9)
    local = ED(False, enc2, dec2, 0.5, False, input_height=16, input_width=16)
    local.load_state_dict(sd, strict=True)       # the band model loads the sharded dict strictly


def _worker(rank, world, port, tmp):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "u-rnn_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import urnn_oracle as O
    from urnn_b200 import dist as ud
    z = np.load(os.path.join(ROOT, "tests", "golden", "ed_32x32_c9.npz"))
    w = {k[2:]: z[k].astype(np.float64) for k in z.files if k.startswith("w.")}
    H = W = 32
    xs = O.synthetic_event_inputs(H, W, 2, 3).astype(np.float64)
    rng = np.random.RandomState(3)
    states = [rng.rand(*s.shape) for s in O.zero_states(H, W, np.float64)]
    full = O.ed_step(w, xs[1], states)                        # unsharded reference (no hook yet)

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    O.STATS_ALLREDUCE = allreduce
    r0, rows = ud.band(H, world, rank)
    sc = [1, 2, 4, 4, 2, 1]
    wl = dict(w)
    for k in list(wl):
        if ".ln." in k:
            wl[k] = wl[k][:, r0:r0 + rows]
    loc_states = [s[:, r0 // c:(r0 + rows) // c] for s, c in zip(states, sc)]
    part = O.ed_step(wl, xs[1][:, r0:r0 + rows], loc_states)
    err = max(float(np.abs(part["states"][i] - full["states"][i][:, r0 // c:(r0 + rows) // c]).max()) for i, c in enumerate(sc))
    err = max(err, float(np.abs(part["prob"] - full["prob"][r0:r0 + rows]).max()))
    np.save(os.path.join(tmp, f"err{rank}.npy"), np.array(err))
    dist.destroy_process_group()


def test_sharded_step_needs_only_stat_allreduce(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert float(np.load(tmp_path / f"err{r}.npy")) < 1e-10
