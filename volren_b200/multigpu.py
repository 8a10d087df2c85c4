"""Multi-GPU rendering, one process per GPU (torch.distributed): the scene is replicated on every rank, the work is
split by spp slice or by image tile and the RGBA32F accumulation buffers are combined with ONE exchange
per frame -- spp slices: `reduce(SUM)` of the float4 images; tiles: every rank sends just its own row band to the
destination rank (point-to-point, no arithmetic) -- over NCCL/NVLink on GPUs, gloo in the CPU tests. Every (pixel, sample) depends only
on (seed, pixel, W, sample index, scene) (reference shader/pathtracer_brick.glsl:28), so no other exchange exists.

The tracer is injected (`trace_fn`): on a GPU box it is `Context.trace` rendering into the bound colour tensor; the
CPU tests inject the oracle so that the partition / reduction logic is exercised without a device.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

ACCUM_MEAN, ACCUM_SUM = 0, 1


def spp_slices(first_sample: int, n_samples: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous slices of the 1-based sample range [first, first + n) for each rank: (first_r, n_r), sum n_r == n."""
    if first_sample < 1 or n_samples < 0 or world < 1:
        raise ValueError("bad sample range / world size")
    out, s = [], first_sample
    for r in range(world):
        n = n_samples * (r + 1) // world - n_samples * r // world
        out.append((s, n))
        s += n
    return out


def row_bands(height: int, world: int, align: int = 4) -> List[Tuple[int, int]]:
    """`align`-row-aligned bands [y0, y1) per rank (the tracer walks 8x4 pixel tiles); bands tile [0, height) exactly."""
    if height < 1 or world < 1:
        raise ValueError("bad height / world size")
    units = (height + align - 1) // align
    return [(min(height, units * r // world * align), min(height, units * (r + 1) // world * align)) for r in range(world)]


class PartitionedRenderer:
    """Renders batches of samples across the ranks of a process group into `color` (a (H, W, 4) float32 torch tensor on
    this rank's device). After `render`, rank `dst` holds the finished MEAN image of all samples rendered since `reset`.

    trace_fn(first_sample, n_samples, tile, accum_mode) must fold the samples into `color` in place:
      accum_mode == ACCUM_SUM : color[tile] += sum of L;  ACCUM_MEAN: reference running mean over sample indices.
    """

    def __init__(self, color, trace_fn: Callable[[int, int, Optional[Sequence[int]], int], None], partition: str = "spp", group=None, dst: int = 0):
        import torch.distributed as dist
        if partition not in ("spp", "tile"):
            raise ValueError("partition must be 'spp' or 'tile'")
        self.color, self.trace_fn, self.partition, self.group, self.dst = color, trace_fn, partition, group, dst
        self.dist = dist
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.samples = 0
        self.holds_sum = False
        self.pending = False

    def reset(self):
        self.samples = 0
        self.holds_sum = False          # spp mode: `color` is a local SUM that has not been reduced yet
        self.pending = False            # tile mode: bands traced since the last gather
        self.color.zero_()

    def render(self, n_samples: int, reduce: bool = True):
        """n more samples per pixel (job-wide). Returns the work this rank did: (first, n) or (y0, y1).
        reduce=False defers the collective: the frame is complete (on rank dst) only after finish() -- one reduce per
        frame instead of one per batch."""
        H, W = int(self.color.shape[0]), int(self.color.shape[1])
        first = self.samples + 1
        if self.world == 1:
            self.trace_fn(first, n_samples, None, ACCUM_MEAN)
            self.samples += n_samples
            return (first, n_samples)
        if self.partition == "spp":
            # every rank keeps a running SUM of its own slices; rank dst's buffer doubles as the reduction target, so its
            # mean of the previous batches is turned back into a sum first
            if not self.holds_sum:
                if self.rank == self.dst and self.samples > 0:
                    self.color.mul_(float(self.samples))
                elif self.rank != self.dst:
                    self.color.zero_()
                self.holds_sum = True
            mine = spp_slices(first, n_samples, self.world)[self.rank]
            if mine[1] > 0:
                self.trace_fn(mine[0], mine[1], None, ACCUM_SUM)
            self.samples += n_samples
            if reduce:
                self.finish()
            return mine
        # tiles: every rank owns a band of rows and keeps the reference running mean there (bit-identical to the single-GPU
        # image); the other rows of its buffer are scratch
        y0, y1 = row_bands(H, self.world)[self.rank]
        if y0 < y1:
            self.trace_fn(first, n_samples, (0, y0, W, y1), ACCUM_MEAN)
        self.samples += n_samples
        self.pending = True
        if reduce:
            self.finish()
        return (y0, y1)

    def finish(self):
        """The one collective of a frame: after it rank dst holds the MEAN image of all samples rendered since reset()."""
        if self.world == 1:
            return
        if self.partition == "spp":
            if not self.holds_sum:
                return
            self.dist.reduce(self.color, dst=self.dst, op=self.dist.ReduceOp.SUM, group=self.group)
            if self.rank == self.dst:
                self.color.mul_(1.0 / float(self.samples))
            self.holds_sum = False
            return
        if not self.pending:
            return
        # gather the bands: rank r sends rows [y0_r, y1_r) to dst, which receives them in place (contiguous row slices of
        # the (H, W, 4) image) -- (world - 1) / world of one image moves instead of a full-image reduce of mostly zeros
        H = int(self.color.shape[0])
        bands = row_bands(H, self.world)
        ops = []
        if self.rank == self.dst:
            for r, (y0, y1) in enumerate(bands):
                if r != self.dst and y0 < y1:
                    ops.append(self.dist.P2POp(self.dist.irecv, self.color[y0:y1], r, self.group))
        else:
            y0, y1 = bands[self.rank]
            if y0 < y1:
                ops.append(self.dist.P2POp(self.dist.isend, self.color[y0:y1], self.dst, self.group))
        if ops:
            for req in self.dist.batch_isend_irecv(ops):
                req.wait()
        self.pending = False
