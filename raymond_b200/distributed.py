"""One process per GPU: partition a render across ranks and combine the accumulators with NCCL.

The path shards naturally — every (pixel, sample) is independent — and the only exchange
is one sum of the per-rank radiance accumulators (W*H*3 f64) onto rank 0.  torch is plumbing here:
it owns the accumulator memory, the stream and the process group (`torch.distributed`, backend "nccl"
over NVLink on the GPU box, "gloo" in CPU tests); the rendering itself is the CUDA library.

    sample partition   rank g renders global samples g, g+G, g+2G, ... of every pixel; the RNG is keyed
                       by the GLOBAL sample index, so the image does not depend on G (FP association aside)
    tile partition     tiles dealt round-robin in the reference's queue order (src/trace.rs:146-172);
                       each rank's accumulator is zero outside its tiles

Every rank keeps its own running sums for the whole render (`tile.data += sample`, src/trace.rs:203); a
checkpoint sums a COPY of them onto rank 0, so rendering can go on afterwards — the reference's progressive
passes (`TileProgressed` every `samples_per_iteration`, src/trace.rs:207-219) across ranks.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np

from . import api as A


def world() -> Tuple[int, int]:
    """(rank, world_size) of the default process group, (0, 1) without one."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return 0, 1


def sample_share(sample_count: int, rank: int, world_size: int) -> Tuple[int, int, int]:
    """(first, count, stride) of the global sample indices rank `rank` renders under the sample partition."""
    if world_size <= 1:
        return 0, sample_count, 1
    count = (sample_count - rank + world_size - 1) // world_size if sample_count > rank else 0
    return rank, count, world_size


def tile_owner(n_tiles: int, world_size: int) -> np.ndarray:
    """Owner rank of every tile (reference queue order) under the tile partition."""
    return np.arange(n_tiles) % max(world_size, 1)


def reduce_sums(accum, dst: int = 0):
    """Sum `accum` of every rank onto `dst`, IN PLACE (ncclReduce over NVLink on GPUs, gloo on CPU tensors).
    On the other ranks the contents of `accum` are unspecified afterwards — pass a copy of anything still needed
    (AccumulatorExchange does).  Asynchronous with respect to the host on CUDA tensors: enqueued on the current stream."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        if accum.is_cuda and dist.get_backend() == "gloo":
            # gloo has no CUDA reduce: the single-GPU tests of the multi-rank logic stage through the host
            host = accum.cpu()
            dist.reduce(host, dst=dst, op=dist.ReduceOp.SUM)
            if dist.get_rank() == dst:
                accum.copy_(host)
        else:
            dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


_pinned_frames = {}


def pinned_frame(shape):
    """A page-locked (H, W, 3) f64 host buffer, one per shape and process, kept for the life of the process: pinning 50 MB
    costs ~10 ms, more than the copy it serves."""
    import torch
    key = tuple(shape)
    if key not in _pinned_frames:
        _pinned_frames[key] = torch.empty(key, dtype=torch.float64, pin_memory=True)
    return _pinned_frames[key]


class AccumulatorExchange:
    """The non-destructive accumulator exchange: `checkpoint(accum)` sums a scratch copy of every rank's running sums
    onto rank 0 and leaves `accum` alone, so it can be called any number of times during a render."""

    def __init__(self, like):
        import torch
        self.scratch = torch.empty_like(like)

    def checkpoint(self, accum, dst: int = 0):
        """The summed running sums on `dst` (a tensor that is overwritten by the next checkpoint), None on the other ranks."""
        self.scratch.copy_(accum)
        reduce_sums(self.scratch, dst)
        return self.scratch if world()[0] == dst else None


def slice_tiles(sums: np.ndarray, layout: np.ndarray, kind: str, sample_count: int, owned=None):
    """The reference's messages for one pass over the tile queue (src/trace.rs:211-219): an independent Tile copy of every
    (owned) tile, in queue order.  `sums` is the (H, W, 3) frame of running sums."""
    for i, (left, top, w, h) in enumerate(layout):
        if owned is not None and not owned[i]:
            continue
        yield A.Message(kind, A.Tile(int(sample_count), int(w), int(h), int(left), int(top), sums[top:top + h, left:left + w].copy()))


class DistributedRenderer:
    """This rank's share of a frame + the accumulator exchange.  Keeps the scene resident between frames.

        dr.render(n)            enqueue this rank's share of the next n global samples (accumulates; nothing is exchanged)
        dr.checkpoint()         sum of all ranks' running sums on rank 0 (device tensor), ranks keep accumulating
        dr.frame(n)             checkpoint + the averaged (H, W, 3) frame on rank 0's host
        dr.render_progressive   the reference's progressive render: TileProgressed every samples_per_iteration, TileFinished
    """

    def __init__(self, scene, settings: A.Settings, *, device: int = 0, seed: int = 0, partition: int = A.PARTITION_SAMPLES,
                 batch_spp: int = 0, flags: int = 0, precision: int = A.PRECISION_F64):
        import torch
        self.torch = torch
        self.rank, self.world_size = world()
        self.settings = settings
        self.partition = partition
        self.device = device
        self.samples_done = 0           # global samples per pixel rendered since the last clear()
        cs = settings.camera_settings
        with torch.cuda.device(device):
            self.stream = torch.cuda.Stream(device=device)
            self.accum = torch.zeros((cs.backbuffer_height, cs.backbuffer_width, 3), dtype=torch.float64, device=f"cuda:{device}")
            self.exchange = AccumulatorExchange(self.accum)
            self._host = pinned_frame(self.accum.shape) if self.rank == 0 else None
        opts = A.GpuOptions(device=device, rank=self.rank, world_size=self.world_size, partition=partition, seed=seed,
                            stream=self.stream.cuda_stream, accum_device=self.accum.data_ptr(), batch_spp=batch_spp, flags=flags,
                            precision=precision)
        self.renderer = A.Renderer(scene, settings, opts)

    def render(self, sample_count: Optional[int] = None, first_sample: Optional[int] = None) -> None:
        """Enqueue this rank's share of the global samples [first_sample, first_sample + sample_count) on self.stream.
        `first_sample` defaults to the samples rendered since the last clear(), so consecutive calls continue the frame."""
        n = self.settings.sample_count if sample_count is None else sample_count
        base = self.samples_done if first_sample is None else first_sample
        if self.partition == A.PARTITION_SAMPLES:
            first, count, stride = sample_share(n, self.rank, self.world_size)
        else:
            first, count, stride = 0, n, 1
        if count:
            self.renderer.render(base + first, count, stride)
        self.samples_done = base + n

    def checkpoint(self):
        """All ranks' running sums added up on rank 0 (a (H, W, 3) f64 device tensor, valid until the next checkpoint; None on
        the other ranks).  Enqueued on self.stream behind the rendering; no rank's accumulator is modified."""
        with self.torch.cuda.stream(self.stream):
            return self.exchange.checkpoint(self.accum, 0)

    def clear(self) -> None:
        self.renderer.clear()
        self.samples_done = 0

    def synchronize(self) -> None:
        self.stream.synchronize()

    def sums(self) -> Optional[np.ndarray]:
        """checkpoint() brought to rank 0's host: the (H, W, 3) running sums of the whole job (None on the other ranks).
        The array is a view of a pinned buffer shared by the renderers of this process: copy it to keep it past the next call."""
        total = self.checkpoint()
        if total is None:
            self.stream.synchronize()
            return None
        with self.torch.cuda.stream(self.stream):
            self._host.copy_(total, non_blocking=True)
        self.stream.synchronize()
        return self._host.numpy()

    def frame(self, sample_count: Optional[int] = None) -> Optional[np.ndarray]:
        """The averaged frame (H, W, 3) on rank 0 (tile.data / sample_count, src/trace.rs:95), None elsewhere.  The division
        runs on the device on the exchanged sums (the same IEEE division) and the result lands in the pinned frame buffer of
        this process: like sums(), the array is a view of that buffer — copy it to keep it past the next call."""
        n = self.samples_done if sample_count is None else sample_count
        total = self.checkpoint()
        if total is None:
            self.stream.synchronize()
            return None
        with self.torch.cuda.stream(self.stream):
            total.div_(float(n))                   # the exchange scratch: overwritten by the next checkpoint anyway
            self._host.copy_(total, non_blocking=True)
        self.stream.synchronize()
        return self._host.numpy()

    def render_progressive(self, on_message: Optional[Callable[[A.Message], None]] = None) -> Optional[np.ndarray]:
        """settings.sample_count samples with a checkpoint every settings.samples_per_iteration (0 = only the final one): rank 0
        receives the reference's messages — TileProgressed per tile per checkpoint, TileFinished per tile at the end, tiles in
        queue order, data = running sums (src/trace.rs:207-219) — and the averaged frame is returned on rank 0."""
        total = self.settings.sample_count
        chunk = self.settings.samples_per_iteration or max(total, 1)
        layout = A.tile_layout(self.settings) if (on_message is not None and self.rank == 0) else None
        self.clear()
        sums = None
        while True:
            n = min(chunk, total - self.samples_done)
            self.render(n)
            last = self.samples_done >= total
            if on_message is None and not last:
                self.checkpoint()              # the exchange still happens at every checkpoint (a viewer could attach later)
                continue
            sums = self.sums()
            if sums is not None and on_message is not None:
                for m in slice_tiles(sums, layout, "TileFinished" if last else "TileProgressed", self.samples_done):
                    on_message(m)
            if last:
                break
        return None if sums is None else sums / float(max(total, 1))

    def stats(self) -> dict:
        return self.renderer.stats()

    def close(self) -> None:
        self.renderer.close()
