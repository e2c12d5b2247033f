"""raymond_b200.distributed.DistributedRenderer on the GPU with world_size 2.

NCCL refuses two ranks on one device, and the driver's test box has one GPU, so the two ranks share cuda:0 and exchange
through a `gloo` group (reduce_sums stages CUDA tensors through the host for gloo).  What is under test is the product's
multi-rank logic on the real CUDA renderer: share of the samples / tiles per rank, accumulation across render() calls,
the non-destructive checkpoint, and the progressive message sequence (src/trace.rs:207-219)."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, out_dir: str):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from raymond_b200 import api as A
        from raymond_b200 import distributed as D
        from raymond_b200 import fixtures as F
        from util import product_scene, settings
        device = rank % max(torch.cuda.device_count(), 1)
        cam, spp = F.camera(96, 64), 6
        scene = product_scene(F.gold_dragon(F.dragon_standin(160, 40)))
        out = {}
        for name, partition in (("samples", A.PARTITION_SAMPLES), ("tiles", A.PARTITION_TILES)):
            st = settings(cam, spp, spi=2)
            dr = D.DistributedRenderer(scene, st, device=device, seed=21, partition=partition)
            # two render() calls, a checkpoint after each, no clear() in between
            dr.render(4)
            first = dr.sums()
            first = None if first is None else first.copy()
            dr.render(2)
            second = dr.sums()
            second = None if second is None else second.copy()
            # a second checkpoint of the same state gives the same sums (nothing was consumed by the first)
            again = dr.sums()
            if rank == 0:
                assert np.array_equal(again, second)
                out[name + "_4"], out[name + "_6"] = first, second
            # the reference's progressive render: messages on rank 0
            msgs = []
            frame = dr.render_progressive(msgs.append if rank == 0 else None)
            if rank == 0:
                layout = A.tile_layout(st)
                kinds = [m.kind for m in msgs]
                assert kinds == ["TileProgressed"] * (2 * len(layout)) + ["TileFinished"] * len(layout)
                assert [m.tile.sample_count for m in msgs] == [2] * len(layout) + [4] * len(layout) + [6] * len(layout)
                assert [(m.tile.left, m.tile.top, m.tile.width, m.tile.height) for m in msgs[-len(layout):]] == [tuple(r) for r in layout]
                out[name + "_frame"] = frame
                built = np.zeros_like(frame)
                for m in msgs[-len(layout):]:
                    t = m.tile
                    built[t.top:t.top + t.height, t.left:t.left + t.width] = t.data / t.sample_count
                assert np.array_equal(built, frame)
            else:
                assert frame is None
            dr.close()
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), **out)
    finally:
        dist.destroy_process_group()


def test_two_ranks_accumulate_and_checkpoint(tmp_path):
    import torch.multiprocessing as mp
    from raymond_b200 import api as A
    from raymond_b200 import fixtures as F
    from util import product_scene, settings
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "out.npz")
    cam = F.camera(96, 64)
    r = A.Renderer(product_scene(F.gold_dragon(F.dragon_standin(160, 40))), settings(cam, 6), A.GpuOptions(seed=21))
    r.render(0, 4)
    want4 = r.read_sums()
    r.render(4, 2)
    want6 = r.read_sums()
    r.close()
    for name in ("samples", "tiles"):
        exact = name == "tiles"           # disjoint tiles: no summation-order freedom
        for key, want in ((name + "_4", want4), (name + "_6", want6), (name + "_frame", want6 / 6.0)):
            if exact:
                assert np.array_equal(got[key], want), key
            else:
                assert np.allclose(got[key], want, rtol=1e-12, atol=1e-12), key
