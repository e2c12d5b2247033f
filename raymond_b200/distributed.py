"""One process per GPU: partition a render across ranks and combine the accumulators with NCCL.

The path shards naturally — every (pixel, sample) is independent — and the only exchange
is one sum of the per-rank radiance accumulators (W*H*3 f64) onto rank 0.  torch is plumbing here:
it owns the accumulator memory, the stream and the process group (`torch.distributed`, backend "nccl"
over NVLink on the GPU box, "gloo" in CPU tests); the rendering itself is the CUDA library.

    sample partition   rank g renders global samples g, g+G, g+2G, ... of every pixel; the RNG is keyed
                       by the GLOBAL sample index, so the image does not depend on G (FP association aside)
    tile partition     tiles dealt round-robin in the reference's queue order (src/trace.rs:146-172);
                       each rank's accumulator is zero outside its tiles
"""
from __future__ import annotations

from typing import Optional, Tuple

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
    """Sum the per-rank accumulators onto `dst` (ncclReduce over NVLink on GPUs, gloo on CPU tensors).
    Asynchronous with respect to the host on CUDA tensors: it is enqueued on the current stream."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(accum, dst=dst, op=dist.ReduceOp.SUM)
    return accum


class DistributedRenderer:
    """This rank's share of a frame + the accumulator exchange.  Keeps the scene resident between frames."""

    def __init__(self, scene, settings: A.Settings, *, device: int = 0, seed: int = 0, partition: int = A.PARTITION_SAMPLES,
                 batch_spp: int = 0, flags: int = 0):
        import torch
        self.torch = torch
        self.rank, self.world_size = world()
        self.settings = settings
        self.partition = partition
        self.device = device
        cs = settings.camera_settings
        with torch.cuda.device(device):
            self.stream = torch.cuda.Stream(device=device)
            self.accum = torch.zeros((cs.backbuffer_height, cs.backbuffer_width, 3), dtype=torch.float64, device=f"cuda:{device}")
        opts = A.GpuOptions(device=device, rank=self.rank, world_size=self.world_size, partition=partition, seed=seed,
                            stream=self.stream.cuda_stream, accum_device=self.accum.data_ptr(), batch_spp=batch_spp, flags=flags)
        self.renderer = A.Renderer(scene, settings, opts)

    def render(self, sample_count: Optional[int] = None, first_sample: int = 0) -> None:
        """Enqueue this rank's share of `sample_count` samples per pixel, then the reduce onto rank 0, on self.stream."""
        n = self.settings.sample_count if sample_count is None else sample_count
        if self.partition == A.PARTITION_SAMPLES:
            first, count, stride = sample_share(n, self.rank, self.world_size)
        else:
            first, count, stride = 0, n, 1
        if count:
            self.renderer.render(first_sample + first, count, stride)
        with self.torch.cuda.stream(self.stream):
            reduce_sums(self.accum, 0)

    def clear(self) -> None:
        self.renderer.clear()

    def synchronize(self) -> None:
        self.stream.synchronize()

    def frame(self, sample_count: Optional[int] = None) -> Optional[np.ndarray]:
        """The averaged frame (H, W, 3) on rank 0 (tile.data / sample_count, src/trace.rs:95), None elsewhere."""
        n = self.settings.sample_count if sample_count is None else sample_count
        self.stream.synchronize()
        if self.rank != 0:
            return None
        return self.renderer.read_frame(n)

    def stats(self) -> dict:
        return self.renderer.stats()

    def close(self) -> None:
        self.renderer.close()
