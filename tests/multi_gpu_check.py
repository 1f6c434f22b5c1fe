"""Run under torchrun with one rank per GPU (>= 2 GPUs): checks the tiled render and the in-library NCCL gather.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/multi_gpu_check.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fluctus_b200 import CLContext, Tracer, dist as fd  # noqa: E402
from fluctus_b200.scene import make_room_scene, room_params  # noqa: E402


def main():
    import torch.distributed as dist
    rank, world, local = fd.init("nccl")
    assert world >= 2
    scene = make_room_scene(materials="mixed", textured=True)
    W, H, N, S = 96, 70, 8192, 4  # 70 rows / stripes of 4: ragged last stripe
    params = room_params(scene, W, H, max_bounces=4, separate_queues=True)
    ctx = CLContext(N, device=local)
    direct = int(os.environ.get("FLX_GATHER_DIRECT", "0"))  # 0 (default): staged + de-interleave; 1: a send / receive per stripe straight into the full image
    ctx.setTuning(gather_direct=direct)
    fd.setup_context(ctx, rank, world, S)
    ctx.uploadSceneData(scene)
    ctx.setupPixelStorage(W, H)
    assert ctx.tilePixels() == fd.tile_pixels(W, H, rank, world, S)
    tr = Tracer(ctx, params)
    tr.start()
    # every rank runs the same path indices: their random sequences must differ (seed = gid + part * numTasks, k_reset), or the
    # noise pattern would repeat from stripe to stripe of the gathered image
    from fluctus_b200 import SLOT
    seeds = [None] * world
    dist.all_gather_object(seeds, ctx.readTasks()[SLOT.SEED, :64].tolist())
    assert len({tuple(s) for s in seeds}) == world, "ranks share random sequences"
    ctx.render(64)
    tile = ctx.readPixels()
    full = np.zeros((W * H, 4), np.float32) if rank == 0 else None
    ctx.gatherPixels(0, full)
    ctx.finishQueue()
    ref_full = fd.gather_host(tile, W, H, S, root=0)
    # display pass of the gathered frame on the root (flx_read_gathered / saveImage on a tiled context): the same kernel every rank
    # runs over its own tile, so gathering the ranks' own previews on the host must give the same bits
    ctx.enqueuePostprocessKernel()
    ref_preview = fd.gather_host(ctx.readPreview(), W, H, S, root=0)
    if rank != 0:
        try:
            ctx.readGathered(W, H)
            raise AssertionError("readGathered must refuse on a rank that was not the gather's root")
        except Exception as e:
            assert "root" in str(e), e
    if rank == 0:
        assert np.array_equal(full, ref_full), "NCCL gather + de-interleave differs from the host-side gather"
        assert np.array_equal(ctx.readGathered(W, H), full), "flx_read_gathered(accumulators) differs from the gathered image"
        prev = ctx.readGathered(W, H, preview=True)
        assert np.array_equal(prev, ref_preview), "display pass of the gathered image differs from the ranks' own display passes"
        import tempfile
        from PIL import Image
        with tempfile.TemporaryDirectory() as tmp:
            ctx.saveImage(os.path.join(tmp, "full.png"))
            got = np.asarray(Image.open(os.path.join(tmp, "full.png")).convert("RGB"))
            want = (np.float32(255) * np.clip(prev.reshape(H, W, 4)[..., :3], 0.0, 1.0)).astype(np.uint8)[::-1]
            assert got.shape == (H, W, 3) and np.array_equal(got, want), "saveImage on the root of a tiled context"
            ctx.saveImage(os.path.join(tmp, "full.hdr"))
            assert os.path.getsize(os.path.join(tmp, "full.hdr")) > W * H
        print("GATHERED_IMAGE_OK")
        assert (full[:, 3] > 0).all(), "some pixels of the full image were never sampled"
        # statistical agreement with an untiled render of the same scene (different seed->pixel map, same estimator)
        solo = CLContext(N * world, device=local)
        solo.uploadSceneData(scene)
        solo.setupPixelStorage(W, H)
        t2 = Tracer(solo, params)
        t2.start()
        solo.render(64)
        p2 = solo.readPixels()
        m1 = (full[:, :3].sum(axis=0) / full[:, 3].sum())
        m2 = (p2[:, :3].sum(axis=0) / p2[:, 3].sum())
        rel = np.abs(m1 - m2) / m2
        assert (rel < 0.03).all(), (m1, m2)
        print("MULTI_GPU_OK world=%d gather_direct=%d mean radiance tiled %s untiled %s" % (world, direct, m1, m2))
    # asynchronous gathers every iteration (own stream, double-buffered snapshot) must not disturb the render, and the frame
    # delivered is the accumulator as of each call
    ctx.gatherPixels(0)
    for _ in range(5):
        ctx.render(1)
        ctx.gatherPixels(0)
    tile = ctx.readPixels()
    full2 = np.zeros((W * H, 4), np.float32) if rank == 0 else None
    ctx.gatherPixels(0, full2)
    ref2 = fd.gather_host(tile, W, H, S, root=0)
    if rank == 0:
        assert np.array_equal(full2, ref2), "gather after a run of asynchronous gathers differs from the host-side gather"
    # a new image size with the same communicator: ragged 64x9 first, then 64x16 (the full-image buffer kept its
    # old, smaller size); ranks that own no rows of the small image sit that one out
    for (w2, h2) in ((64, 9), (64, 16)):
        p2 = room_params(scene, w2, h2, max_bounces=3, separate_queues=True)
        if fd.tile_pixels(w2, h2, rank, world, 8) == 0 or world > 2:
            continue
        ctx.setTile(rank, world, 8)
        ctx.setupPixelStorage(w2, h2)
        t3 = Tracer(ctx, p2)
        t3.start()
        ctx.render(8)
        tile = ctx.readPixels()
        fullr = np.zeros((w2 * h2, 4), np.float32) if rank == 0 else None
        ctx.gatherPixels(0, fullr)
        refr = fd.gather_host(tile, w2, h2, 8, root=0)
        if rank == 0:
            assert np.array_equal(fullr, refr), "gather after resize to %dx%d" % (w2, h2)
            print("RESIZE_GATHER_OK %dx%d" % (w2, h2))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
