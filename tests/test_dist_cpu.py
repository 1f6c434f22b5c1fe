"""Host-side logic of the multi-GPU path on CPU: stripe partition, local->global pixel map, de-interleave, and a
world_size-2 gloo run of the gather plumbing (torch.distributed, 127.0.0.1 rendezvous)."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest

from fluctus_b200 import dist as fd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("h,n,s", [(1080, 8, 8), (1080, 2, 8), (90, 4, 16), (7, 3, 2), (5, 8, 1), (2160, 8, 8)])
def test_stripes_partition_the_image(h, n, s):
    rows = [fd.tile_rows(h, p, n, s) for p in range(n)]
    allrows = np.concatenate(rows)
    assert sorted(allrows.tolist()) == list(range(h))  # every row exactly once (ragged last stripes included)
    for p in range(n):
        assert fd.tile_pixels(13, h, p, n, s) == len(rows[p]) * 13
        if len(rows[p]):
            loc = np.arange(len(rows[p]) * 13)
            g = fd.local_to_global_pixel(loc, 13, p, n, s)
            assert np.array_equal(g // 13, np.repeat(rows[p], 13)) and np.array_equal(g % 13, np.tile(np.arange(13), len(rows[p])))
    if h >= n * s:  # balanced to within one stripe
        sizes = [len(r) for r in rows]
        assert max(sizes) - min(sizes) <= s


def test_deinterleave_inverts_tiling():
    w, h, n, s = 11, 37, 3, 4
    full = np.arange(w * h * 4, dtype=np.float32).reshape(w * h, 4)
    tiles = [full.reshape(h, w, 4)[fd.tile_rows(h, p, n, s)].reshape(-1, 4) for p in range(n)]
    assert np.array_equal(fd.deinterleave(tiles, w, h, n, s), full)


def test_gloo_world_size_2_gather(tmp_path):
    script = tmp_path / "rank.py"
    script.write_text(textwrap.dedent("""
        import sys, numpy as np
        sys.path.insert(0, %r)
        from fluctus_b200 import dist as fd
        import torch.distributed as dist
        rank, world, local = fd.init("gloo")
        assert world == 2
        w, h, s = 16, 21, 4
        full = np.arange(w * h * 4, dtype=np.float32).reshape(w * h, 4)
        tile = full.reshape(h, w, 4)[fd.tile_rows(h, rank, world, s)].reshape(-1, 4)
        img = fd.gather_host(tile, w, h, s, root=0)
        tot, mx = fd.reduce_scalars([float(rank + 1)]), fd.reduce_scalars([float(rank + 1)], "max")
        assert tot == [3.0] and mx == [2.0]
        if rank == 0:
            assert np.array_equal(img, full)
            print("GATHER_OK")
        else:
            assert img is None
        dist.barrier(); dist.destroy_process_group()
    """ % ROOT))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29600 + os.getpid() % 300), str(script)], capture_output=True, text=True, timeout=240, env=env)
    assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
