"""Host-side multi-GPU logic on CPU: world_size-2 process groups over gloo. The tracer injected into
PartitionedRenderer is the CPU oracle (test infrastructure), so what is under test is the partitioning, the sum /
assembly through the collective and the progressive bookkeeping -- the same code path bench.py drives with NCCL."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_spp_slices_and_row_bands():
    from volren_b200.multigpu import row_bands, spp_slices
    for world in (1, 2, 3, 4, 8):
        for first, n in ((1, 16), (5, 7), (1, 1), (3, 0)):
            sl = spp_slices(first, n, world)
            assert len(sl) == world and sum(k for _, k in sl) == n and sl[0][0] == first
            for (a, ka), (b, _) in zip(sl, sl[1:]):
                assert b == a + ka                                   # contiguous, in rank order
            assert max(k for _, k in sl) - min(k for _, k in sl) <= 1  # balanced
        for h in (1, 3, 4, 64, 270, 1080):
            bands = row_bands(h, world)
            assert bands[0][0] == 0 and bands[-1][1] == h
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 <= a1
            assert all(y0 % 4 == 0 for y0, _ in bands)
    with pytest.raises(ValueError):
        spp_slices(0, 4, 2)


def _worker(rank, world, port, partition, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="2")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from helpers import readme_scene
    from oracle.binding import Oracle
    from volren_b200 import formats
    from volren_b200.multigpu import PartitionedRenderer
    assets = os.path.join(ROOT, "tests", "golden", "assets")
    grid = formats.load_brick(os.path.join(assets, "smoke.brick"))
    env = formats.load_hdr(os.path.join(assets, "table_mountain_2_puresky_1k.hdr"))
    o = Oracle()
    sc = o.make_scene(grid, env, o.env_build(env))
    W, H = 32, 24
    p = readme_scene(grid, W, H, bounces=4)
    color = torch.zeros((H, W, 4), dtype=torch.float32)
    view = color.numpy()                                             # shares memory with the tensor

    def trace_fn(first, n, tile, accum):
        o.trace(sc, p, first, n, color=view, tile=tile, accum_mode=accum)

    deferred = partition.endswith("_deferred")
    pr = PartitionedRenderer(color, trace_fn, partition=partition.split("_")[0])
    pr.reset()
    if deferred:                                                     # one collective per frame
        work = [pr.render(3, reduce=False), pr.render(2, reduce=False)]
        pr.finish()
        pr.finish()                                                  # idempotent
    else:
        work = [pr.render(3), pr.render(2)]                          # two progressive batches: 5 spp in total
    if rank == 0:
        np.save(os.path.join(out_dir, f"{partition}.npy"), view.copy())
    np.save(os.path.join(out_dir, f"{partition}_work{rank}.npy"), np.array(work))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("partition", ["tile", "spp", "tile_deferred", "spp_deferred"])
def test_partitioned_render_world2_gloo(partition, tmp_path, oracle, smoke_grid, env_rgb, env_pyramid):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import readme_scene
    port = 29500 + (os.getpid() % 2000) + ["tile", "spp", "tile_deferred", "spp_deferred"].index(partition)
    mp.spawn(_worker, args=(2, port, partition, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / f"{partition}.npy")
    W, H = 32, 24
    p = readme_scene(smoke_grid, W, H, bounces=4)
    sc = oracle.make_scene(smoke_grid, env_rgb, env_pyramid)
    want, _ = oracle.trace(sc, p, 1, 5)                              # single process, reference running mean over samples 1..5
    w0, w1 = np.load(tmp_path / f"{partition}_work0.npy"), np.load(tmp_path / f"{partition}_work1.npy")
    if partition.startswith("tile"):
        assert np.array_equal(got, want)                             # assembly is exact
        assert w0[0].tolist() == [0, 12] and w1[0].tolist() == [12, 24]
    else:
        # sum / N vs the sequential running mean: fp32 rounding only
        assert np.allclose(got, want, rtol=2e-5, atol=1e-6)
        assert w0.tolist() == [[1, 1], [4, 1]] and w1.tolist() == [[2, 2], [5, 1]]
