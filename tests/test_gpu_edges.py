"""GPU edge cases of the path: several grids in one scene (one traversal pass per grid object), a grid shared by two objects
(Arc<AccGrid>), empty scenes, tiny and ragged frames, deep bounce limits, tiles larger than the frame."""
import numpy as np
import pytest

from oracle import oracle as O
from raymond_b200 import api as A
from raymond_b200 import fixtures as F

from test_gpu_render import compare_same_stream, gpu_render
from util import assert_hits_equal, oracle_scene, product_scene, settings

pytestmark = pytest.mark.gpu


def two_mesh_scene():
    a = F.translate(F.bumpy_sphere(30, 60, 0.6, 0.1, (1.0, 0.8, 0.6)), (-0.7, -0.3, 3.0))
    b = F.translate(F.dragon_standin(120, 30), (0.5, -0.2, 3.4))
    return [F.RED_SPHERE, ("grid", a, ("Metal", (0.9, 0.9, 0.9), 0.05)), ("grid", b, F.DRAGON_MATERIAL)] + F.BOX_PLANES


def test_two_grids_bit_exact_and_render():
    objs = two_mesh_scene()
    cam = F.camera(200, 120)
    rays = np.concatenate([O.primary_rays(cam), F.random_rays(30000, 4, ((-1.9, 1.9), (-0.9, 1.9), (-1.9, 4.9)))])
    want = oracle_scene(objs).intersect(rays)
    assert {1, 2} <= set(want[0].tolist())
    assert_hits_equal(product_scene(objs).intersect(rays), want, "two grids")
    g, gs = gpu_render(objs, F.camera(128, 72), 4, seed=3)
    o, oc = O.render(oracle_scene(objs), F.camera(128, 72), 4, seed=3)
    compare_same_stream(g, o, 4, "two grids render")


def test_shared_grid_two_objects():
    """The same Arc<AccGrid> pushed twice with different materials: ties go to the first object (scene.rs:61)."""
    tris = F.translate(F.bumpy_sphere(20, 40, 0.5), (0.0, 0.0, 3.0))
    grid = A.AccGrid.build_from_mesh(A.Mesh.new(tris))
    s = A.Scene()
    s.push_grid(grid, A.Material.Metal((1, 1, 0.1), 0.15))
    s.push_grid(grid, A.Material.Diffuse((0.2, 0.9, 0.2), 0.5))
    osc = O.Scene()
    og = O.AccGrid.build_from_mesh(O.Mesh.from_triangles(tris))
    osc.add_grid(og, ("Metal", (1, 1, 0.1), 0.15))
    osc.add_grid(og, ("Diffuse", (0.2, 0.9, 0.2), 0.5))
    rays = O.primary_rays(F.camera(160, 120))
    want = osc.intersect(rays)
    got = s.intersect(rays)
    assert_hits_equal(got, want, "shared grid")
    assert set(got[0].tolist()) == {-1, 0}                    # equal distances: the first object wins


def test_empty_scene_and_too_many_objects():
    s = A.Scene()
    obj, sub, t = s.intersect(O.primary_rays(F.camera(16, 8)))
    assert (obj == -1).all()
    r = A.Renderer(s, settings(F.camera(16, 8), 2), A.GpuOptions(seed=1))
    r.render(0, 2)
    assert not r.read_sums().any()
    big = A.Scene()
    for i in range(65):
        big.push_sphere((i * 0.1, 0, 5), 0.01, A.Material.Diffuse((1, 1, 1), 0.5))
    with pytest.raises(A.RaymondError) as e:
        big.intersect(np.array([[0, 0, 0, 0, 0, 1.0]]))
    assert e.value.status == A.RM_ERR_UNSUPPORTED


@pytest.mark.parametrize("w,h,tile", [(1, 1, (32, 32)), (33, 7, (32, 32)), (50, 40, (64, 64)), (37, 29, (5, 3))])
def test_ragged_frames_and_tiles(w, h, tile):
    objs, cam, spp = F.reflective_spheres(), F.camera(w, h), 3
    st = settings(cam, spp, tile=tile, spi=1)
    task = A.render_tiled(product_scene(objs), st, A.GpuOptions(seed=2))
    task.stats()
    msgs = []
    while (m := task.poll()) is not None:
        msgs.append(m)
    layout = A.tile_layout(st)
    assert [tuple(r) for r in layout] == [tuple(r) for r in O.tile_layout(cam, tile).tolist()]
    fin = [m for m in msgs if m.kind == "TileFinished"]
    assert len(fin) == len(layout) and sum(m.tile.width * m.tile.height for m in fin) == w * h
    want, _ = O.render(oracle_scene(objs), cam, spp, seed=2, tile_size=tile)
    got = np.zeros((h, w, 3))
    for m in fin:
        t = m.tile
        got[t.top:t.top + t.height, t.left:t.left + t.width] = t.data
    assert np.abs(got - want).max() <= 1e-9 * max(np.abs(want).max(), 1.0) or (np.abs(got - want).max(axis=-1) > 1e-9).mean() < 0.02


def test_deep_bounce_limit():
    objs, cam, spp = F.reflective_spheres(), F.camera(48, 32), 2
    g, gs = gpu_render(objs, cam, spp, seed=5, bounce_limit=24)
    o, oc = O.render(oracle_scene(objs), cam, spp, seed=5, bounce_limit=24)
    compare_same_stream(g, o, spp, "bounce_limit 24")
    assert gs["rays"] > 0


def test_extreme_rays_bit_exact():
    """Rays a caller may pass to Scene::intersect that stress the conservative pre-test and the DDA set-up: non-unit and tiny /
    huge directions, far-away origins aimed at the mesh, axis-aligned directions with +0 / -0 components, NaN / Inf inputs."""
    tris = F.translate(F.bumpy_sphere(40, 80, 1.0, 0.1, (1.0, 0.7, 0.45)), (0.1, -0.2, 3.0))
    objs = F.soup_scene(tris)
    rng = np.random.default_rng(12)
    base = F.random_rays(4000, 31, ((-3.0, 3.0), (-3.0, 3.0), (0.0, 6.0)))
    target = np.array([0.1, -0.2, 3.0]) + rng.uniform(-0.7, 0.7, size=(4000, 3)) * np.array([1.0, 0.7, 0.45])
    aimed = base.copy()
    aimed[:, 3:] = target - aimed[:, :3]                                  # un-normalised directions, |d| ~ 1..6
    sets = [aimed]
    for s in (1e-6, 1e-3, 37.5, 1e6, 1e12):
        r = aimed.copy(); r[:, 3:] *= s; sets.append(r)
    far = aimed.copy()
    d = far[:, 3:] / np.linalg.norm(far[:, 3:], axis=1, keepdims=True)
    for dist in (1e3, 1e6, 1e9):
        r = far.copy(); r[:, :3] = target - d * dist; r[:, 3:] = d; sets.append(r)
    axis = np.zeros((600, 6))
    axis[:, :3] = np.array([0.1, -0.2, 3.0]) + rng.uniform(-0.9, 0.9, size=(600, 3)) * np.array([1.0, 0.7, 0.45])
    axis[:, 2] -= 4.0
    axis[:, 5] = 1.0
    axis[::2, 3] = -0.0; axis[::3, 4] = -0.0
    sets.append(axis)
    ax2 = axis.copy(); ax2[:, :3] = axis[:, [2, 0, 1]]; ax2[:, 3:] = axis[:, [5, 3, 4]]; ax2[:, 0] += 4.1; ax2[:, 1] -= 0.3; ax2[:, 2] += 7.2
    sets.append(ax2)
    weird = aimed[:64].copy()
    weird[0, 0] = np.nan; weird[1, 4] = np.nan; weird[2, 3] = np.inf; weird[3, 1] = -np.inf; weird[4, 3:] = 0.0; weird[5, 3:] = (0.0, 0.0, 1e-300)
    weird[6, :3] = 1e300; weird[7, 3:] = 1e300
    sets.append(weird)
    rays = np.ascontiguousarray(np.concatenate(sets))
    want = oracle_scene(objs).intersect(rays, threads=8)
    got = product_scene(objs).intersect(rays)
    assert_hits_equal(got, want, "extreme rays")
    assert (want[0] >= 0).sum() > 10000


def test_release_cached_memory_then_render_again():
    objs, cam, spp = F.reflective_spheres(), F.camera(96, 64), 3
    a, _ = gpu_render(objs, cam, spp, seed=2)
    A.release_cached_memory()
    b, _ = gpu_render(objs, cam, spp, seed=2)
    assert np.array_equal(a, b)


def test_grid_image_is_reused_across_uploads_and_scenes():
    """The flattened device image of an (immutable) grid is built at its first upload and kept pinned with the grid: later
    uploads — the same scene again, another scene holding the same Arc<AccGrid>, an upload after the caches were dropped — must
    give the same bits, for queries and for frames."""
    tris = F.translate(F.dragon_standin(96, 24), (0.0, -0.2, 3.2))
    grid = A.AccGrid.build_from_mesh(A.Mesh.new(tris))
    s1 = A.Scene()
    s1.push_grid(grid, A.Material.Metal((1, 0.8, 0.3), 0.2))
    for ob in F.BOX_PLANES:
        s1.push_plane(ob[1], ob[2], A.Material.from_fixture(ob[3]))
    cam = F.camera(160, 96)
    rays = O.primary_rays(cam)
    first = s1.intersect(rays)
    again = s1.intersect(rays)
    A.release_cached_memory()
    after_release = s1.intersect(rays)
    s2 = A.Scene()
    s2.push_sphere((0.0, 0.0, 30.0), 0.5, A.Material.Diffuse((1, 1, 1), 0.5))      # object indices shift by one
    s2.push_grid(grid, A.Material.Diffuse((0.2, 0.9, 0.2), 0.5))
    other = s2.intersect(rays)
    grid_hits = first[0] == 0
    assert grid_hits.sum() > 1000
    for got, what in ((again, "second upload"), (after_release, "upload after release")):
        assert_hits_equal(got, first, what)
    assert np.array_equal(other[0][grid_hits], np.ones(grid_hits.sum(), dtype=other[0].dtype))
    assert np.array_equal(other[1][grid_hits], first[1][grid_hits]) and np.array_equal(other[2][grid_hits].view(np.uint64), first[2][grid_hits].view(np.uint64))
    st = A.Settings(A.CameraSettings.from_fixture(cam), 3)
    frames = [A.render_tiled(s1, st, A.GpuOptions(seed=5)).await_() for _ in range(2)]
    assert np.array_equal(frames[0], frames[1])


def test_read_frame_is_sums_over_count():
    objs, cam, spp = F.reflective_spheres(), F.camera(640, 480), 3        # above the threaded read-back threshold
    sc = product_scene(objs)
    r = A.Renderer(sc, A.Settings(A.CameraSettings.from_fixture(cam), spp), A.GpuOptions(seed=4))
    r.render(0, spp)
    sums, frame = r.read_sums(), r.read_frame(spp)
    r.close()
    assert np.array_equal(frame, sums / float(spp))


def test_fp64_rate_probe_is_plausible():
    """rm_measure_fp64_rate: a B200 has 148 SMs x 64 f64 lanes; without FMA that is ~18.6 Tera-op/s at 1.965 GHz."""
    g = A.measure_fp64_rate(0)
    assert 5e3 < g < 4e4, g


def test_grid_skip_margin_on_a_large_box_with_grazing_rays():
    """k_setup skips a grid traversal when an analytic object is hit before the grid's box is entered, by a margin that bounds
    the rounding error of a computed triangle distance (|o - v0| <= max(tmin, 0) + diagonal).  A 2000-unit box holding slivers
    whose |a| is close to Triangle::intersects' 1e-8 threshold, a sphere grazing the box surface and rays that clip the box
    corners (small tmax - tmin, far v0): the skip must never change which object Scene::intersect reports."""
    rng = np.random.default_rng(5)
    big = F.translate(F.bumpy_sphere(24, 48, 900.0, 0.1, (1.0, 0.9, 0.6)), (0.0, 0.0, 2500.0))
    # near-degenerate slivers: long thin triangles scattered through the box
    n = 4000
    c = rng.uniform(-800, 800, (n, 3)) * np.array([1.0, 1.0, 0.4]) + np.array([0.0, 0.0, 2500.0])
    e1 = rng.normal(size=(n, 3)) * np.array([300.0, 300.0, 60.0])      # res.y >= res.z: the only regime the reference's cell index survives (A1)
    e2 = e1 * rng.uniform(0.2, 0.9, (n, 1)) + rng.normal(size=(n, 3)) * rng.choice([1e-6, 1e-4, 1e-2], (n, 1))
    sliv = F.make_triangles(c, c + e1, c + e2)
    tris = np.concatenate([big, sliv])
    lo, hi = F.positions(tris).reshape(-1, 3).min(axis=0), F.positions(tris).reshape(-1, 3).max(axis=0)
    objs = [("sphere", (lo[0] - 40.0, 0.0, 2500.0), 39.999, ("Diffuse", (0.8, 0.2, 0.2), 0.5)),        # grazes the box's -x face
            ("sphere", (0.0, 0.0, 1200.0), 300.0, ("Metal", (0.2, 0.2, 0.9), 0.1)),                     # in front of the box
            ("grid", tris, F.DRAGON_MATERIAL),
            ("plane", (0.0, lo[1] - 1e-3, 0.0), (0.0, 1.0, 0.0), ("Diffuse", (0.5, 0.5, 0.5), 0.5))]    # just under the box
    m = 200_000
    corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    target = corners[rng.integers(0, 8, m)] + rng.normal(size=(m, 3)) * rng.choice([1e-3, 1.0, 50.0], (m, 1))
    origin = rng.uniform(-3000, 3000, (m, 3)) + np.array([0.0, 1500.0, 0.0])
    d = target - origin
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays = np.concatenate([np.concatenate([origin, d], axis=1), O.primary_rays(F.camera(320, 200))])
    want = oracle_scene(objs).intersect(rays, threads=8)
    assert {0, 1, 2, 3} <= set(want[0].tolist())
    assert_hits_equal(product_scene(objs).intersect(rays), want, "large box, grazing rays")
