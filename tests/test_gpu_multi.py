"""Multi-GPU (spatially sharded) == single-GPU, on real GPUs: one process per GPU, one row band each, statistics
exchanged inside the kernels over peer-mapped buffers.  World sizes 2, 4 and 8 (each skipped unless that many CUDA
devices are visible); all three arithmetic modes; the 8-rank case runs enough steps to wrap the exchange ring."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, HIST = 48, 3
MODES = ("fp32", "bf16", "f16x3")


def _build(Hl, math):   # noqa: E302
    from src.lib.model.networks.model import ED
    from src.lib.model.networks.net_params import get_network_params
    torch.manual_seed(0)
    enc, dec = get_network_params(False, Hl, W, input_channels=2 * HIST + 3, math=math)
    return ED(False, enc, dec, 0.5, False, input_height=Hl, input_width=W)


def _worker(rank, world, port, tmp, H, T):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "u-rnn_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import urnn_oracle as O
    from urnn_b200 import dist as ud
    ud.init_spatial_sharding()
    dev = f"cuda:{rank}"
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, HIST))
    r0, rows = ud.band(H, world, rank)
    errs = {}
    for math in MODES:
        full_sd = _build(H, math).state_dict()
        net = _build(rows, math)
        net.load_state_dict(ud.shard_state_dict(full_sd, world, rank), strict=True)
        net = net.to(dev).eval()
        st = [torch.zeros(1, *s.shape, device=dev) for s in O.zero_states(rows, W)]
        outs = []
        with torch.no_grad():
            for t in range(T):
                out, *st = net(xs[t][None, None, :, r0:r0 + rows].contiguous().to(dev), *st)
                outs.append(out[0, 0].cpu().numpy())
        ref = np.load(os.path.join(tmp, f"ref_{math}.npz"))
        sc = [1, 2, 4, 4, 2, 1]
        e = 0.0
        for i, c in enumerate(sc):
            e = max(e, float(np.abs(st[i].cpu().numpy()[0] - ref[f"state{i}"][:, r0 // c:(r0 + rows) // c]).max()))
        errs[math] = e
        if math == "f16x3":
            # the sequence entry point: encoder(t+1) || decoder + head (t) on two streams, each half on its own exchange lane
            from urnn_b200.runner import SequenceRunner
            depth, _, fin = SequenceRunner(net, rows, W, 2 * HIST + 3).run_dev(xs[:, :, r0:r0 + rows].contiguous().to(dev), want_prob=False)
            torch.cuda.synchronize()
            eseq = float(np.abs(depth.cpu().numpy() - ref["seq_depth"][:, r0:r0 + rows]).max())
            for i, c in enumerate(sc):
                eseq = max(eseq, float(np.abs(fin[i].cpu().numpy() - ref[f"seq_state{i}"][:, r0 // c:(r0 + rows) // c]).max()))
            errs["seq"] = eseq
    np.save(os.path.join(tmp, f"err{rank}.npy"), np.array([errs[m] for m in MODES] + [errs["seq"]]))
    ud.shutdown_spatial_sharding()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,T", [(2, 3), (4, 3), (8, 6)])
def test_bands_equal_single_gpu(tmp_path, world, T):
    """T steps on a (32 * world) x 48 grid: `world` bands of 32 rows vs the same grid on one GPU.  Every step issues 17
    statistic exchanges, so T = 6 (102 exchanges) wraps the 64-entry exchange ring of the communicator."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import urnn_oracle as O
    H = 32 * world
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, HIST)).cuda()
    for math in MODES:
        net = _build(H, math).cuda().eval()
        st = [torch.zeros(1, *s.shape, device="cuda") for s in O.zero_states(H, W)]
        with torch.no_grad():
            for t in range(T):
                out, *st = net(xs[t][None, None], *st)
        ref = {f"state{i}": s.cpu().numpy()[0] for i, s in enumerate(st)}
        if math == "f16x3":
            from urnn_b200.runner import SequenceRunner
            depth, _, fin = SequenceRunner(net, H, W, 2 * HIST + 3).run_dev(xs, want_prob=False)
            ref["seq_depth"] = depth.cpu().numpy()
            ref.update({f"seq_state{i}": s.cpu().numpy() for i, s in enumerate(fin)})
        np.savez(tmp_path / f"ref_{math}.npz", **ref)
        del net
    port = 29500 + (os.getpid() % 2000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path), H, T), nprocs=world, join=True)
    for r in range(world):
        e = np.load(tmp_path / f"err{r}.npy")
        assert e[0] < 1e-5, e          # fp32: only the summation order of the statistics differs
        assert e[1] < 2e-2, e          # bf16: rounding flips at bf16 boundaries, same bound as the single-GPU tests
        assert e[2] < 2e-5, e          # f16x3: split products + merge order of the statistics
        assert e[3] < 2e-5, e          # f16x3 sequence call (two streams, two exchange lanes) vs the single-GPU sequence call


# ------------------------------------------------------------------------------------------------ sharded backward
def _grad_worker(rank, world, port, tmp, H, T):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "u-rnn_b200")]
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from oracle import urnn_oracle as O
    from urnn_b200 import dist as ud
    ud.init_spatial_sharding()
    dev = f"cuda:{rank}"
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, HIST))
    r0, rows = ud.band(H, world, rank)
    full_sd = _build(H, "fp32").state_dict()
    net = _build(rows, "fp32")
    net.load_state_dict(ud.shard_state_dict(full_sd, world, rank), strict=True)
    net = net.to(dev).train()
    torch.manual_seed(5)
    label = torch.rand(T, H, W)[:, r0:r0 + rows].to(dev)
    st = [torch.zeros(1, *s.shape, device=dev) for s in O.zero_states(rows, W)]
    regs = []
    for t in range(T):
        out, *st = net(xs[t][None, None, :, r0:r0 + rows].contiguous().to(dev), *st)
        regs.append(out)
    reg = torch.cat(regs, dim=1)
    loss = ((reg[0] - label) ** 2).sum() / (T * H * W)          # normalised by the GLOBAL element count
    loss.backward()
    n = ud.allreduce_window_gradients(net)
    assert n == 420146, n                                         # replicated parameters of the published architecture
    norm = ud.clip_grad_norm_sharded(net, 1e9)                    # no clipping, returns the global norm
    ref = np.load(os.path.join(tmp, "ref_grads.npz"))
    worst = 0.0
    for k, v in net.named_parameters():
        g = v.grad.cpu().numpy()
        r = ref[k]
        if ".ln." in k and v.dim() == 3:
            r = r[:, r0:r0 + rows]
        scale = max(float(np.abs(ref[k]).max()), 1e-6)
        worst = max(worst, float(np.abs(g - r).max()) / scale)
    np.save(os.path.join(tmp, f"gerr{rank}.npy"), np.array([worst, float(norm), float(ref["__norm"])]))
    ud.shutdown_spatial_sharding()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_backward_equals_single_gpu(tmp_path, world):
    """SURVEY.md 8e backward: GroupNorm / LayerNorm backward sums exchanged in-kernel, NCCL all-reduce of the replicated
    weight gradients, global clip norm: every gradient of a 2-step window equals the unsharded one (fp32 kernels)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import urnn_oracle as O
    H, T = 32 * world, 2
    xs = torch.from_numpy(O.synthetic_event_inputs(H, W, T, HIST)).cuda()
    net = _build(H, "fp32").cuda().train()
    torch.manual_seed(5)
    label = torch.rand(T, H, W).cuda()
    st = [torch.zeros(1, *s.shape, device="cuda") for s in O.zero_states(H, W)]
    regs = []
    for t in range(T):
        out, *st = net(xs[t][None, None], *st)
        regs.append(out)
    loss = ((torch.cat(regs, dim=1)[0] - label) ** 2).sum() / (T * H * W)
    loss.backward()
    grads = {k: v.grad.cpu().numpy() for k, v in net.named_parameters()}
    grads["__norm"] = np.array(float(torch.sqrt(sum(v.grad.double().pow(2).sum() for v in net.parameters()))))
    np.savez(tmp_path / "ref_grads.npz", **grads)
    del net
    port = 29700 + (os.getpid() % 2000) + world
    mp.spawn(_grad_worker, args=(world, port, str(tmp_path), H, T), nprocs=world, join=True)
    for r in range(world):
        e = np.load(tmp_path / f"gerr{r}.npy")
        assert e[0] < 2e-3, e                                   # atomics in the weight-gradient GEMM: summation order
        assert abs(e[1] - e[2]) <= 1e-3 * e[2], e               # global gradient norm
