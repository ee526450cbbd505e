"""N > 1 path on CPU: world_size 2 over gloo.  The sharding plan (sample slices / interleaved rows) and the
single reduce to rank 0 are exercised with the CPU oracle standing in for the per-GPU render leg: the
oracle seeds its RNG per (pixel, sample), like the GPU's Philox keying, so the shards of a frame must add
up to the unsharded frame."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import raytrace_clj_b200 as rt
from raytrace_clj_b200 import parallel

NX, NY, NS = 48, 32, 6


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _scene():
    sc = rt.scene.make_random_scene(NX, NY, 11, True, random.Random(1))
    return rt.native.marshal_world(sc["world"]), rt.native.marshal_camera(sc["camera"])


def _worker(rank, world, port, mode, out_path):
    import oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        flat, (cam_type, cam) = _scene()
        S = oracle.Scene(flat)

        def render_slice(s_begin, s_count, row_offset, row_stride, out):
            acc, _ = S.render_accumulate(cam_type, cam, NX, NY, s_begin, s_count, 50, seed=7, row_offset=row_offset,
                                         row_stride=row_stride, n_threads=2)
            out += torch.from_numpy(acc.astype(np.float32))

        total = parallel.render_sharded(render_slice, NX, NY, NS, mode=mode)
        if rank == 0:
            np.save(out_path, total.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["samples", "rows"])
def test_two_rank_shards_add_up(tmp_path, mode):
    import oracle

    out = str(tmp_path / f"sum_{mode}.npy")
    mp.spawn(_worker, args=(2, _free_port(), mode, out), nprocs=2, join=True)
    got = np.load(out)
    flat, (cam_type, cam) = _scene()
    full, _ = oracle.Scene(flat).render_accumulate(cam_type, cam, NX, NY, 0, NS, 50, seed=7, n_threads=2)
    assert got.shape == (NY, NX, 3) and got.sum() > 0
    assert np.allclose(got, full.astype(np.float32), rtol=1e-5, atol=1e-5)


def test_shard_plans_tile_the_work():
    for world in (1, 2, 3, 4, 8):
        for ns in (1, 7, 10, 1024):
            spans = [parallel.sample_slice(ns, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == ns
            for (b0, c0), (b1, _) in zip(spans, spans[1:]):
                assert b0 + c0 == b1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
        rows = [parallel.row_interleave(world, r) for r in range(world)]
        covered = sorted(j for off, st in rows for j in range(off, 37, st))
        assert covered == list(range(37))
    with pytest.raises(ValueError):
        parallel.sample_slice(4, 2, 2)


def test_render_sharded_single_process():
    calls = []

    def render_slice(b, c, off, st, out):
        calls.append((b, c, off, st))
        out += 1.0

    t = parallel.render_sharded(render_slice, 4, 3, 10, mode="samples", world=1, rank=0)
    assert calls == [(0, 10, 0, 1)] and t.shape == (3, 4, 3) and float(t.sum()) == 36.0
