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
