"""Host-side multi-GPU logic on CPU: world_size-2 `gloo` process group.

The path shards with no data-path collective (every (pixel, sample) is independent); the only exchange
is the sum of the per-rank radiance accumulators onto rank 0.  Here each rank's share is rendered by the oracle
(there is no GPU in this container) with exactly the (first, count, stride) / tile ownership the product's
raymond_b200.distributed hands its CUDA renderer, and the product's reduce_sums() combines them over gloo."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, mode: str, out_dir: str):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        from raymond_b200 import api as A
        from raymond_b200 import distributed as D
        from raymond_b200 import fixtures as F
        from util import oracle_scene, settings
        assert D.world() == (rank, world)
        cam, spp, seed = F.camera(64, 48), 5, 17
        sc = oracle_scene(F.reflective_spheres())
        if mode == "samples":
            first, count, stride = D.sample_share(spp, rank, world)
            sums, cnt = O.render(sc, cam, count, seed=seed, first_sample=first, sample_stride=stride, worker_count=2) if count else (np.zeros((48, 64, 3)), {"samples": 0})
        else:
            layout = A.tile_layout(settings(cam, spp, tile=(16, 16)))
            owner = D.tile_owner(len(layout), world)
            full, cnt = O.render(sc, cam, spp, seed=seed, tile_size=(16, 16), worker_count=2)
            sums = np.zeros_like(full)
            for (l, t, w, h), o in zip(layout, owner):
                if o == rank:
                    sums[t:t + h, l:l + w] = full[t:t + h, l:l + w]
        if mode == "progressive":
            # two passes accumulate into the rank's own running sums; a checkpoint after each, and nothing is cleared in between
            # (TileProgressed semantics, src/trace.rs:207-219): the exchange must not disturb any rank's accumulator
            acc = torch.zeros((48, 64, 3), dtype=torch.float64)
            ex = D.AccumulatorExchange(acc)
            shots = []
            for done, n in ((0, 2), (2, 3)):
                first, count, stride = D.sample_share(n, rank, world)
                if count:
                    part, _ = O.render(sc, cam, count, seed=seed, first_sample=done + first, sample_stride=stride, worker_count=2)
                    acc += torch.from_numpy(part)
                before = acc.clone()
                total = ex.checkpoint(acc, 0)
                assert torch.equal(acc, before), "a checkpoint modified the rank's own running sums"
                assert (total is None) == (rank != 0)
                if rank == 0:
                    shots.append(total.numpy().copy())
            if rank == 0:
                np.save(os.path.join(out_dir, "progressive.npy"), np.stack(shots))
            return
        acc = torch.from_numpy(sums.copy())
        D.reduce_sums(acc, 0)
        if rank == 0:
            np.save(os.path.join(out_dir, f"{mode}.npy"), acc.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["samples", "tiles"])
def test_two_rank_partition_and_reduce(tmp_path, mode):
    from oracle import oracle as O
    from raymond_b200 import fixtures as F
    from util import oracle_scene
    port = _free_port()
    mp.spawn(_worker, args=(2, port, mode, str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / f"{mode}.npy")
    want, _ = O.render(oracle_scene(F.reflective_spheres()), F.camera(64, 48), 5, seed=17, tile_size=(16, 16) if mode == "tiles" else (32, 32))
    if mode == "tiles":
        assert np.array_equal(got, want)                     # disjoint tiles: the sum is exact
    else:
        assert np.allclose(got, want, rtol=1e-12, atol=1e-12)   # same samples, different association


def test_two_rank_progressive_checkpoints_do_not_double_count(tmp_path):
    """Two passes, a checkpoint after each, no clear() in between: checkpoint k is the 1-rank image of the samples so far."""
    from oracle import oracle as O
    from raymond_b200 import fixtures as F
    from util import oracle_scene
    mp.spawn(_worker, args=(2, _free_port(), "progressive", str(tmp_path)), nprocs=2, join=True)
    shots = np.load(tmp_path / "progressive.npy")
    sc, cam = oracle_scene(F.reflective_spheres()), F.camera(64, 48)
    first2, _ = O.render(sc, cam, 2, seed=17)
    all5, _ = O.render(sc, cam, 5, seed=17)
    assert np.allclose(shots[0], first2, rtol=1e-12, atol=1e-12)
    assert np.allclose(shots[1], all5, rtol=1e-12, atol=1e-12)


def test_slice_tiles_is_the_reference_message_sequence():
    from raymond_b200 import api as A
    from raymond_b200 import distributed as D
    from raymond_b200 import fixtures as F
    from util import settings
    st = settings(F.camera(70, 50), 3, tile=(32, 32))
    layout = A.tile_layout(st)
    sums = np.arange(50 * 70 * 3, dtype=np.float64).reshape(50, 70, 3)
    msgs = list(D.slice_tiles(sums, layout, "TileProgressed", 2))
    assert [(m.tile.left, m.tile.top, m.tile.width, m.tile.height) for m in msgs] == [tuple(r) for r in layout]
    assert [(m.tile.left, m.tile.top) for m in msgs[:3]] == [(0, 0), (0, 32), (32, 0)]      # y advances first (src/trace.rs:146-172)
    for m in msgs:
        t = m.tile
        assert m.kind == "TileProgressed" and t.sample_count == 2
        assert np.array_equal(t.data, sums[t.top:t.top + t.height, t.left:t.left + t.width])
    own = D.tile_owner(len(layout), 2) == 1
    assert len(list(D.slice_tiles(sums, layout, "TileFinished", 3, owned=own))) == int(own.sum())


def test_sample_share_covers_every_sample_once():
    from raymond_b200 import distributed as D
    for n in (0, 1, 5, 8, 500, 4096):
        for world in (1, 2, 3, 4, 8):
            seen = []
            for r in range(world):
                first, count, stride = D.sample_share(n, r, world)
                seen += [first + i * stride for i in range(count)]
            assert sorted(seen) == list(range(n)), (n, world)
    assert D.tile_owner(7, 3).tolist() == [0, 1, 2, 0, 1, 2, 0]
