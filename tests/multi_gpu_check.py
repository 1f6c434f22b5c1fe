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
    fd.setup_context(ctx, rank, world, S)
    ctx.uploadSceneData(scene)
    ctx.setupPixelStorage(W, H)
    assert ctx.tilePixels() == fd.tile_pixels(W, H, rank, world, S)
    tr = Tracer(ctx, params)
    tr.start()
    ctx.render(64)
    tile = ctx.readPixels()
    full = np.zeros((W * H, 4), np.float32) if rank == 0 else None
    ctx.gatherPixels(0, full)
    ctx.finishQueue()
    ref_full = fd.gather_host(tile, W, H, S, root=0)
    if rank == 0:
        assert np.array_equal(full, ref_full), "NCCL gather + de-interleave differs from the host-side gather"
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
        print("MULTI_GPU_OK world=%d mean radiance tiled %s untiled %s" % (world, m1, m2))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
