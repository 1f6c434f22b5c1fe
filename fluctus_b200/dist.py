"""Image-space sharding across the GPUs of one box: one process per GPU, full scene replica and private
path state per rank, rows dealt to ranks in interleaved stripes, ONE collective per frame -- the gather of the per-tile
radiance buffers.  torch.distributed is the plumbing (rendezvous, broadcasting the NCCL id, host-side reductions of
scalars); the data-path gather runs inside the library on its own stream (flx_gather_pixels, NCCL send/recv group +
de-interleave kernel).  The helpers below are the host mirror of the device mapping in flx_kernels.cuh
(local_pixel_to_xy) and flx_api.cu (localRows, k_deinterleave); tests/test_dist_cpu.py exercises them over gloo."""
import numpy as np


def tile_rows(height, part, n_parts, stripe_rows):
    """Global row numbers owned by `part`, in local (top-to-bottom) order."""
    rows = []
    s = part
    while s * stripe_rows < height:
        rows.extend(range(s * stripe_rows, min((s + 1) * stripe_rows, height)))
        s += n_parts
    return np.asarray(rows, np.int64)


def tile_pixels(width, height, part, n_parts, stripe_rows):
    return int(len(tile_rows(height, part, n_parts, stripe_rows)) * width)


def local_to_global_pixel(local, width, part, n_parts, stripe_rows):
    """Mirror of local_pixel_to_xy: local pixel index -> index in the full image."""
    local = np.asarray(local, np.int64)
    x, ly = local % width, local // width
    stripe, within = ly // stripe_rows, ly % stripe_rows
    y = (stripe * n_parts + part) * stripe_rows + within
    return y * width + x


def deinterleave(tiles, width, height, n_parts, stripe_rows):
    """tiles: list of (tile_pixels, C) arrays in rank order -> (height*width, C) full image."""
    c = tiles[0].shape[1]
    full = np.zeros((height * width, c), tiles[0].dtype)
    for part, t in enumerate(tiles):
        rows = tile_rows(height, part, n_parts, stripe_rows)
        assert t.shape[0] == len(rows) * width
        full.reshape(height, width, c)[rows] = t.reshape(len(rows), width, c)
    return full


def init(backend=None):
    """Join the job torchrun started (RANK/WORLD_SIZE/MASTER_* from the environment). Returns (rank, world, local_rank)."""
    import os
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def setup_context(ctx, rank, world, stripe_rows=8):
    """Give `ctx` its tile and (world > 1) join the library's NCCL communicator: rank 0 creates the id, torch.distributed
    broadcasts it."""
    ctx.setTile(rank, world, stripe_rows)
    if world > 1:
        import torch.distributed as dist
        uid = [ctx.commUniqueId() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.commInit(uid[0], rank, world)


def gather_host(tile, width, height, stripe_rows, root=0):
    """Backend-agnostic gather of host tiles through torch.distributed (used on CPU/gloo and as a cross-check of
    flx_gather_pixels); returns the full image on root, None elsewhere."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    out = [None] * world if rank == root else None
    dist.gather_object(np.ascontiguousarray(tile), out, dst=root)
    if rank != root:
        return None
    return deinterleave(out, width, height, world, stripe_rows)


def reduce_scalars(values, op="sum"):
    """Sum or max of a few python floats over ranks (bench bookkeeping)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        return list(values)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(list(values), dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]
