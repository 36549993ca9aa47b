"""Spatial sharding of the encoder-decoder step across the GPUs of one box (one process per GPU).

The grid is split into equal row bands aligned to 4 rows, so the two 2x2 poolings and the two 2x2 transposed
convolutions stay band-local; with 1x1 filters the only cross-band quantities are the GroupNorm / LayerNorm
statistics, which liburnn_b200 all-reduces inside its kernels over peer-mapped exchange buffers (see
include/urnn_b200.h, urnn_comm_*).  torch.distributed is used for plumbing only: exchanging the 64-byte IPC handles.
Weights are replicated; the head's LayerNorm([16,H,W]) affine parameters are sharded with the rows.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import _capi


def band(H, world, rank):
    """(first row, number of rows) of this rank's band; H must split into `world` bands of a multiple of 4 rows."""
    if H % (4 * world) != 0:
        raise ValueError(f"H={H} does not split into {world} bands of a multiple of 4 rows")
    rows = H // world
    return rank * rows, rows


def shard_rows(t, world, rank, scale=1):
    """Rows of a (..., H/scale, W/scale) map that belong to `rank` (scale = 1, 2, 4 for the three resolutions)."""
    Hs = t.shape[-2]
    r0, rows = band(Hs * scale, world, rank)
    return t[..., r0 // scale:(r0 + rows) // scale, :].contiguous()


def shard_state_dict(sd, world, rank):
    """Global (reference-layout, 254-key) state_dict -> this rank's: LayerNorm affine maps (16,H,W) are sliced by rows,
    everything else is replicated.  Aliased keys stay aliased (same slice object per storage)."""
    out, cache = {}, {}
    for k, v in sd.items():
        if ".ln." in k and v.dim() == 3:
            key = v.data_ptr()
            if key not in cache:
                cache[key] = shard_rows(v, world, rank)
            out[k] = cache[key]
        else:
            out[k] = v
    return out


def init_spatial_sharding(group=None):
    """Create the in-kernel statistics communicator for this process (call once after init_process_group)."""
    lib = _capi.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return 1
    handle = (C.c_ubyte * 64)()
    _capi.check(lib.urnn_comm_local_init(world, rank, handle), "urnn_comm_local_init")
    mine = torch.tensor(list(handle), dtype=torch.uint8)
    backend = dist.get_backend(group)
    if backend == "nccl":
        mine = mine.cuda()
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine, group=group)
    blob = bytes(torch.cat([g.cpu() for g in gathered]).tolist())
    _capi.check(lib.urnn_comm_connect(blob), "urnn_comm_connect")
    dist.barrier(group)
    return world


def shutdown_spatial_sharding(group=None):
    if dist.is_initialized():
        dist.barrier(group)
    _capi.load().urnn_comm_destroy()


# ---- training under spatial sharding (SURVEY.md 8e "Collectives -- backward") -----------------------------------------
def _is_sharded(name, p):
    """LayerNorm([16,H,W]) affine maps of the head live with the rows of their band; everything else is replicated."""
    return ".ln." in name and p.dim() == 3


def allreduce_window_gradients(net, group=None):
    """After loss.backward() on every rank's band: sum the gradients of the REPLICATED parameters over the ranks with one
    flattened all-reduce (420 146 floats = 1.7 MB for the published architecture; NCCL over NVLink).  The in-kernel
    statistic exchanges already made every rank's backward see the whole grid; what is left is the sum over pixels of the
    weight gradients.  Gradients of the sharded LayerNorm maps stay local.  The loss must be normalised by the GLOBAL
    number of elements (e.g. band mean / world) so that the summed gradient is the gradient of the global loss."""
    named = [(n, p) for n, p in net.named_parameters() if p.grad is not None and not _is_sharded(n, p)]
    if not named or dist.get_world_size(group) == 1:
        return 0
    flat = torch.cat([p.grad.reshape(-1) for _, p in named])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for _, p in named:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return flat.numel()


def clip_grad_norm_sharded(net, max_norm, group=None, eps=1e-6):
    """torch.nn.utils.clip_grad_norm_ (main.py:759-761) when the LayerNorm maps are sharded: the global 2-norm is
    sqrt(|replicated grads|^2 + sum over ranks |local sharded grads|^2) -- one scalar all-reduce."""
    rep = torch.zeros((), device=next(net.parameters()).device, dtype=torch.float64)
    loc = torch.zeros_like(rep)
    for n, p in net.named_parameters():
        if p.grad is None:
            continue
        s = p.grad.double().pow(2).sum()
        if _is_sharded(n, p):
            loc += s
        else:
            rep += s
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(loc, op=dist.ReduceOp.SUM, group=group)
    total = torch.sqrt(rep + loc)
    coef = torch.clamp(max_norm / (total + eps), max=1.0).float()
    for p in net.parameters():
        if p.grad is not None:
            p.grad.mul_(coef)
    return total.float()
