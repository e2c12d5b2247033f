"""Several GPUs driven from ONE process through the C ABI (rm_gpu_options.device_count): samples or tiles are split between
the devices and the accumulators summed onto the first one with peer copies.  Needs >= 2 GPUs (gpurun --gpus 2)."""
import numpy as np
import pytest

from raymond_b200 import api as A
from raymond_b200 import fixtures as F

from util import product_scene, settings

pytestmark = pytest.mark.gpu


def _gpu_count():
    import torch
    return torch.cuda.device_count()


needs_two = pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs")


@needs_two
@pytest.mark.parametrize("partition", [A.PARTITION_SAMPLES, A.PARTITION_TILES])
def test_device_count_two_matches_one(partition):
    objs, cam, spp = F.gold_dragon(F.dragon_standin(160, 40)), F.camera(160, 96), 7
    st = settings(cam, spp, spi=3)
    one = A.render_tiled(product_scene(objs), st, A.GpuOptions(seed=5))
    ref = one.await_()
    two = A.render_tiled(product_scene(objs), st, A.GpuOptions(seed=5, device_count=2, partition=partition))
    stats = two.stats()
    msgs = []
    while (m := two.poll()) is not None:
        msgs.append(m)
    assert stats["samples"] == 160 * 96 * spp
    prog = sorted({m.tile.sample_count for m in msgs if m.kind == "TileProgressed"})
    assert prog == [3, 6]
    got = np.zeros_like(ref)
    for m in msgs:
        if m.kind == "TileFinished":
            t = m.tile
            got[t.top:t.top + t.height, t.left:t.left + t.width] = t.data / t.sample_count
    if partition == A.PARTITION_TILES:
        assert np.array_equal(got, ref)                        # disjoint tiles: identical sums
    else:
        assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)   # same samples, different association across GPUs


@needs_two
def test_device_count_rejects_single_device_options():
    sc = product_scene(F.reflective_spheres())
    with pytest.raises(A.RaymondError) as e:
        A.render_tiled(sc, settings(F.camera(16, 16), 2), A.GpuOptions(device_count=2, world_size=2, rank=0))
    assert e.value.status == A.RM_ERR_INVALID_ARGUMENT
    with pytest.raises(A.RaymondError) as e:
        A.render_tiled(sc, settings(F.camera(16, 16), 2), A.GpuOptions(device_count=64))
    assert e.value.status == A.RM_ERR_CUDA
