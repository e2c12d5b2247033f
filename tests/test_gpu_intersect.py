"""GPU parity, bit-exact scope: Scene::intersect (scene.rs:54-74) through the C ABI vs the oracle.

Object index, triangle index and the distance bits must be identical (north_star (a):
"hit object/triangle index bit-exact, t within 1e-5 relative" — identical bits is the stronger
statement and is what the f64 / no-FMA kernels are built to deliver)."""
import numpy as np
import pytest

from oracle import oracle as O
from raymond_b200 import api as A
from raymond_b200 import fixtures as F

from util import assert_hits_equal, oracle_scene, product_scene, small_meshes

pytestmark = pytest.mark.gpu


def _mixed_rays(cam, n_random=20000, seed=0xD1CE, box=((-2.0, 2.0), (-1.0, 2.0), (-2.0, 5.0))):
    return np.concatenate([O.primary_rays(cam), F.random_rays(n_random, seed, box)])


@pytest.mark.parametrize("name", ["cube", "bumpy", "bumpy_fine", "tube"])
def test_single_grid_bit_exact(name):
    tris = F.translate(small_meshes()[name], (0.1, -0.2, 3.0))
    objs = F.soup_scene(tris)
    rays = _mixed_rays(F.camera(160, 120))
    want = oracle_scene(objs).intersect(rays)
    assert (want[0] >= 0).sum() > 100
    assert_hits_equal(product_scene(objs).intersect(rays), want, name)


def test_grid_quirks_reproduced():
    """+side origins miss (A3), OOB start cells miss, first-hit-cell return (A2) — same as the oracle, not 'fixed'."""
    tris = F.cube()
    objs = F.soup_scene(tris)
    rays = np.array([[0.1, 0.05, -3, 0, 0, 1], [-2, 0.03, 0.02, 1, 0, 0], [0.1, -3, 0.05, 0, 1, 0],
                     [2, 0, 0, -1, 0, 0], [0.1, 0.05, 3, 0, 0, -1], [0.1, 3, 0.05, 0, -1, 0],
                     [0.0, 0.0, 0.0, 0.0, 0.0, 1.0], [0.0, 0.0, 0.0, -0.0, 0.0, 1.0], [0.3, 0.3, -2.0, 0.0, -0.0, 1.0]], dtype=np.float64)
    want = oracle_scene(objs).intersect(rays)
    got = product_scene(objs).intersect(rays)
    assert_hits_equal(got, want, "cube quirks")
    assert list(got[0][:6]) == [0, 0, 0, -1, -1, -1]
    assert got[2][0] == 2.4789250000000003 and got[2][1] == 1.4789249999999998


def test_incoherent_rays_aliased_grid():
    """Random rays from all around an anisotropic mesh: exercises fix-up starts, +side misses and index aliasing (A1)."""
    tris = small_meshes()["bumpy_fine"]
    objs = F.soup_scene(tris)
    rays = F.random_rays(60000, 7, ((-3.0, 3.0), (-3.0, 3.0), (-3.0, 3.0)))
    # aim half of them at the mesh
    target = F.random_rays(60000, 11, ((-0.8, 0.8), (-0.5, 0.5), (-0.3, 0.3)))[:, :3]
    d = target - rays[:, :3]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays[::2, 3:] = d[::2]
    want = oracle_scene(objs).intersect(rays)
    assert (want[0] >= 0).sum() > 5000
    assert_hits_equal(product_scene(objs).intersect(rays), want, "aliased grid")


@pytest.mark.parametrize("scene_name", ["reflective_spheres", "gold_dragon_small"])
def test_benchmark_scene_primaries(scene_name):
    """Primary rays (pixel centres and jittered) of the benchmark scenes: spheres, planes and the grid, ties by object order."""
    if scene_name == "reflective_spheres":
        objs = F.reflective_spheres()
    else:
        objs = F.gold_dragon(F.dragon_standin(240, 60))
    cam = F.camera(320, 180)
    jit = np.random.default_rng(3).random((320 * 180, 2))
    rays = np.concatenate([O.primary_rays(cam), O.primary_rays(cam, jit), F.random_rays(20000, 5, ((-1.9, 1.9), (-0.9, 1.9), (-1.9, 4.9)))])
    want = oracle_scene(objs).intersect(rays)
    assert_hits_equal(product_scene(objs).intersect(rays), want, scene_name)
    assert len(set(want[0].tolist())) >= 5


def test_triangle_soup_primaries():
    """C4 at a size the oracle finishes in seconds: 200k-triangle soups, cubic (B1) and flat/aliased (B2) boxes."""
    cam = F.camera(480, 270)
    for box in (F.SOUP_BOX_CUBIC, F.SOUP_BOX_FLAT):
        tris = F.triangle_soup(200_000, box)
        objs = F.soup_scene(tris)
        rays = np.concatenate([O.primary_rays(cam), F.random_rays(1 << 16)])
        want = oracle_scene(objs).intersect(rays, threads=8)
        assert (want[0] >= 0).mean() > 0.1
        assert_hits_equal(product_scene(objs).intersect(rays), want, f"soup {box}")


def test_primary_rays_bit_exact():
    """generate_primary_ray (src/trace.rs:322-333) with the jitter forced to 0: identical bits."""
    import torch
    for (w, h, fov, pos) in ((320, 240, 55.0, (0, 0, 0)), (1920, 1080, 55.0, (0.5, -0.25, -4.0)), (97, 61, 90.0, (1, 2, 3))):
        cam = F.camera(w, h, fov_vert=fov, position=pos)
        want = O.primary_rays(cam)
        buf = torch.empty((w * h, 6), dtype=torch.float64, device="cuda")
        A.primary_rays_device(A.CameraSettings.from_fixture(cam), 0, buf.data_ptr(), torch.cuda.current_stream().cuda_stream)
        got = buf.cpu().numpy()
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64))


def test_device_resident_query_matches_host_query():
    import torch
    objs = F.gold_dragon(F.dragon_standin(120, 30))
    sc = product_scene(objs)
    rays = O.primary_rays(F.camera(200, 100))
    want = sc.intersect(rays)
    ds = A.DeviceScene(sc, 0)
    r = torch.from_numpy(rays).cuda()
    obj = torch.full((rays.shape[0],), -7, dtype=torch.int64, device="cuda")
    sub = torch.zeros(rays.shape[0], dtype=torch.int64, device="cuda")
    t = torch.zeros(rays.shape[0], dtype=torch.float64, device="cuda")
    ds.intersect_device(r.data_ptr(), rays.shape[0], obj.data_ptr(), sub.data_ptr(), t.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(obj.cpu().numpy(), want[0])
    hit = want[0] >= 0
    assert np.array_equal(sub.cpu().numpy()[hit], want[1][hit].astype(np.int64))
    assert np.array_equal(t.cpu().numpy()[hit], want[2][hit])


def test_empty_and_single_ray():
    sc = product_scene(F.reflective_spheres())
    obj, sub, t = sc.intersect(np.zeros((0, 6)))
    assert obj.shape == (0,)
    obj, sub, t = sc.intersect(np.array([[0, 0, 0, 0, 0, 1.0]]))
    want = oracle_scene(F.reflective_spheres()).intersect(np.array([[0, 0, 0, 0, 0, 1.0]]))
    assert obj[0] == want[0][0] and t[0] == want[2][0]
