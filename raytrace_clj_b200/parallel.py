"""Sharding of one render over the GPUs of a box.

Every (pixel, sample) is independent given the counter-based RNG, which is keyed by the GLOBAL
pixel and sample index; the scene is replicated.  Two partitionings:

  * sample slices (default): rank g renders samples [g*S/G, (g+1)*S/G) of every pixel into a
    full-frame float32 sum buffer — perfect load balance;
  * interleaved rows: rank g renders rows j = g (mod G).

Either way the per-rank sums are combined with ONE reduce to rank 0 (NCCL over NVLink on GPUs,
gloo in the CPU tests) and rank 0 resolves (core.clj:52-57).  There is no other collective on the path.
"""
from __future__ import annotations


def sample_slice(nsamples: int, world: int, rank: int):
    """(begin, count) of rank's sample slice; slices tile [0, nsamples) exactly."""
    if not (0 <= rank < world) or nsamples < 0:
        raise ValueError("bad shard arguments")
    b = nsamples * rank // world
    e = nsamples * (rank + 1) // world
    return b, e - b


def row_interleave(world: int, rank: int):
    """(row_offset, row_stride): rank renders rows j with j % row_stride == row_offset."""
    if not (0 <= rank < world):
        raise ValueError("bad shard arguments")
    return rank, world


def reduce_sums(sums, dst: int = 0):
    """Sum the per-rank float buffers onto rank `dst` (no-op without an initialised process group)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(sums, dst=dst, op=dist.ReduceOp.SUM)
    return sums


def render_sharded(render_slice, nx, ny, nsamples, *, mode="samples", world=None, rank=None, device=None):
    """Render this rank's shard with ``render_slice(sample_begin, sample_count, row_offset, row_stride, out)``
    (which must ADD into ``out``, a [ny, nx, 3] float32 tensor) and reduce to rank 0.
    Returns the reduced sum tensor (meaningful on rank 0)."""
    import torch
    import torch.distributed as dist

    if world is None:
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    out = torch.zeros(ny, nx, 3, dtype=torch.float32, device=device)
    if mode == "samples":
        b, c = sample_slice(nsamples, world, rank)
        render_slice(b, c, 0, 1, out)
    elif mode == "rows":
        off, stride = row_interleave(world, rank)
        render_slice(0, nsamples, off, stride, out)
    else:
        raise ValueError(mode)
    return reduce_sums(out)
