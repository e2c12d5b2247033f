"""GPU parity of BASELINE.json's configs at their own sizes, through the reference-facing calls.

    configs[2]  C3  ReflectiveSpheres + aperture sampling 1920x1080, tiles dealt to 8 shares   (src/trace.rs:335-360, :142-173)
    configs[3]  C4  4 M and 10 M triangle soups, both boxes, bit-exact hit indices              (acc_grid.rs:89-185)
    configs[4]  C5  GoldDragon 3840x2160, progressive (samples_per_iteration)                   (src/trace.rs:207-219)

The full sample counts (1000 / 4096 spp) are 2-34 G paths: the CPU oracle renders the same frames at 2-4 spp in seconds,
and the comparison is sample for sample (same counter-based RNG stream), so it does not get weaker with fewer samples.
On top of the fraction of paths that may differ at all, `compare_paths` bounds how MUCH they may differ:
every differing path must still be a plausible path value, and the differences must not be biased."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from raymond_b200 import api as A
from raymond_b200 import fixtures as F

from util import assert_hits_equal, oracle_scene, product_scene, settings

pytestmark = pytest.mark.gpu

THREADS = os.cpu_count() or 8
LUMA = np.array([0.2126, 0.7152, 0.0722])


def gpu_sample_frames(ps, st, spp, seed, **opt):
    """One frame of radiance per global sample index (the running sum is read and cleared between samples)."""
    r = A.Renderer(ps, st, A.GpuOptions(seed=seed, **opt))
    frames = []
    for s in range(spp):
        r.clear()
        r.render(s, 1)
        frames.append(r.read_sums())
    stats = r.stats()
    r.close()
    return frames, stats


def oracle_sample_frames(osc, cam, spp, seed):
    return [O.render(osc, cam, 1, seed=seed, first_sample=s, worker_count=THREADS)[0] for s in range(spp)]


def compare_paths(gpu_frames, ora_frames, what, max_differing=1e-5, max_gross=2e-6):
    """Path-for-path comparison of two same-stream renders (one frame per sample).

    A path "differs" when its radiance is off by more than 1e-9 relative.  Measured on B200: NO path differs on either full
    1920x1080 x 2 frame (4 147 200 paths each; largest relative difference 4e-11) — CUDA's sincos and the closed forms used
    for cos(acos(x)) differ from glibc by <= 1 ulp, and nothing on these frames amplifies that past 1e-9.  A last-ulp
    difference flipping a lobe / hit decision remains possible in principle, so a handful of paths are allowed, and for
    those the magnitude and the bias are bounded too.  Asserted:
      * at most `max_differing` of the paths differ at all, at most `max_gross` by more than 1e-3 relative;
      * every differing GPU path value is finite, non-negative and no larger than twice the largest path value the oracle
        produced anywhere in these frames (it is a path the reference could have traced, not garbage);
      * the differences are unbiased: |mean luminance difference| over the differing paths <= 4 standard errors;
      * the mean luminance of the whole frame agrees to 0.2 %."""
    g, o = np.stack(gpu_frames), np.stack(ora_frames)
    assert np.isfinite(g).all(), f"{what}: non-finite path value"
    diff = np.abs(g - o).max(axis=-1)
    scale = np.maximum(np.abs(o).max(axis=-1), 1e-3)
    rel = diff / scale
    differing = rel > 1e-9
    gross = rel > 1e-3
    frac, gfrac = differing.mean(), gross.mean()
    bound = 2.0 * o.max()
    gd, od = g[differing], o[differing]
    report = f"{what}: {int(differing.sum())} of {differing.size} paths differ = {frac:.5%} (by more than 1e-3: {int(gross.sum())}; largest relative difference {rel.max():.3g})"
    print(report)
    assert frac <= max_differing, report
    assert gfrac <= max_gross, report
    if gd.size:
        assert gd.min() >= 0.0 and gd.max() <= bound, f"{what}: a differing path value {gd.max()} is outside [0, {bound}]"
        dl = np.clip(gd, 0, 10) @ LUMA - np.clip(od, 0, 10) @ LUMA
        se = dl.std() / np.sqrt(dl.size)
        assert abs(dl.mean()) <= 4.0 * se + 1e-12, f"{what}: differing paths are biased: mean {dl.mean()} vs standard error {se} over {dl.size} paths"
    gl, ol = (np.clip(g, 0, 10) @ LUMA).mean(), (np.clip(o, 0, 10) @ LUMA).mean()
    assert abs(gl - ol) <= 0.002 * ol, f"{what}: mean luminance {gl} vs {ol}"
    return frac


def noise_floor(gpu_mean, ora_a, ora_b, what):
    """RMSE(GPU, oracle seed A) <= 1.15 RMSE(oracle seed B, oracle seed A) per channel on linear radiance clamped to
    [0, 10] and on the tonemapped image; |mean luminance difference| <= max(0.5 %, 3 standard errors).  Different RNG streams."""
    clip = lambda x: np.clip(x, 0, 10)
    rmse0 = np.sqrt(((clip(ora_a) - clip(ora_b)) ** 2).mean(axis=(0, 1)))
    rmse = np.sqrt(((clip(gpu_mean) - clip(ora_a)) ** 2).mean(axis=(0, 1)))
    assert (rmse <= 1.15 * rmse0).all(), f"{what}: rmse {rmse} vs floor {rmse0}"
    la, lg = clip(ora_a) @ LUMA, clip(gpu_mean) @ LUMA
    se = np.sqrt(la.var() / la.size + lg.var() / lg.size)
    assert abs(la.mean() - lg.mean()) <= max(0.005 * la.mean(), 3 * se), f"{what}: mean luminance {lg.mean()} vs {la.mean()}"
    ta, tb, tg = (F.tonemap(x).astype(float) for x in (ora_a, ora_b, gpu_mean))
    assert np.sqrt(((tg - ta) ** 2).mean()) <= 1.15 * np.sqrt(((tb - ta) ** 2).mean()), f"{what}: tonemapped rmse above the floor"


PRECISIONS = [A.PRECISION_F64, A.PRECISION_F32_SHADING]


# ------------------------------------------------------------------------------------------ C3

C3_CAMERA = dict(focal_length=2.5, aperture_radius=0.5)      # the reference's only non-zero values, server/src/main.rs:148-149


def test_c3_dof_full_frame_paths():
    """generate_primary_ray_with_dof at 1920x1080, path for path against the oracle."""
    objs, cam, spp = F.reflective_spheres(), F.camera(1920, 1080, **C3_CAMERA), 2
    g, stats = gpu_sample_frames(product_scene(objs), settings(cam, spp), spp, seed=31)
    o = oracle_sample_frames(oracle_scene(objs), cam, spp, seed=31)
    compare_paths(g, o, "C3 DoF 1920x1080x2")
    assert stats["samples"] == 1920 * 1080 * spp and stats["nonfinite_samples"] == 0


def test_c3_tiles_over_eight_shares():
    """configs[2] as stated — tiles split across 8 — through render_tiled with an 8-entry device list (GPU 0 eight times on a
    1-GPU box): every share owns tiles i, i+8, ... of the reference's column-major tile order and the combined frame equals
    the single-share frame exactly (disjoint tiles: no summation-order freedom)."""
    objs, cam, spp = F.reflective_spheres(), F.camera(1920, 1080, **C3_CAMERA), 2
    st = settings(cam, spp)
    ps = product_scene(objs)
    eight = A.render_tiled(ps, st, A.GpuOptions(seed=31, device_list=[0] * 8, partition=A.PARTITION_TILES))
    frame = eight.await_()
    stats = eight.stats()
    assert stats["samples"] == 1920 * 1080 * spp
    one = A.render_tiled(ps, st, A.GpuOptions(seed=31)).await_()
    assert np.array_equal(frame, one)
    want, _ = O.render(oracle_scene(objs), cam, spp, seed=31, worker_count=THREADS)
    rel = np.abs(frame * spp - want).max(axis=-1) / np.maximum(np.abs(want).max(axis=-1), 1e-3 * spp)
    print(f"C3 tiles over eight shares: {int((rel > 1e-9).sum())} of {rel.size} pixels differ beyond 1e-9 relative (largest {float(rel.max()):.3g})")
    assert (rel > 1e-9).mean() < 1e-4          # measured: none
    # tile ownership of share 3 (a rank of a multi-process job renders exactly these pixels)
    layout = A.tile_layout(st)
    assert len(layout) == 60 * 34
    r = A.Renderer(ps, st, A.GpuOptions(seed=31, rank=3, world_size=8, partition=A.PARTITION_TILES))
    r.render(0, spp)
    part = r.read_sums()
    r.close()
    owner = np.full((1080, 1920), -1)
    for i, (l, t, w, h) in enumerate(layout):
        owner[t:t + h, l:l + w] = i % 8
    assert not part[owner != 3].any()
    assert np.array_equal(part[owner == 3], (one * spp)[owner == 3])


@pytest.mark.parametrize("precision", PRECISIONS, ids=["f64", "f32shade"])
def test_c3_dof_noise_floor(precision):
    objs, cam, spp = F.reflective_spheres(), F.camera(480, 270, **C3_CAMERA), 64
    osc = oracle_scene(objs)
    a = O.render(osc, cam, spp, seed=41, worker_count=THREADS)[0] / spp
    b = O.render(osc, cam, spp, seed=42, worker_count=THREADS)[0] / spp
    r = A.Renderer(product_scene(objs), settings(cam, spp), A.GpuOptions(seed=43, precision=precision))
    r.render(0, spp)
    g = r.read_frame(spp)
    r.close()
    noise_floor(g, a, b, f"C3 DoF 480x270x{spp} precision {precision}")


# ------------------------------------------------------------------------------------------ C2 / C5 (GoldDragon stand-in)

@pytest.fixture(scope="module")
def dragon():
    objs = F.gold_dragon(F.dragon_standin())
    return objs, product_scene(objs), oracle_scene(objs)


def test_c2_gold_dragon_full_frame_paths(dragon):
    """configs[1]'s frame, path for path (the 871 200-triangle stand-in, 1920x1080)."""
    objs, ps, osc = dragon
    cam, spp = F.camera(1920, 1080), 2
    g, stats = gpu_sample_frames(ps, settings(cam, spp), spp, seed=2026)
    o = oracle_sample_frames(osc, cam, spp, seed=2026)
    compare_paths(g, o, "C2 GoldDragon 1920x1080x2")
    assert stats["nonfinite_samples"] == 0


@pytest.mark.parametrize("precision", PRECISIONS, ids=["f64", "f32shade"])
def test_c2_gold_dragon_noise_floor(dragon, precision):
    objs, ps, osc = dragon
    cam, spp = F.camera(480, 270), 32
    a = O.render(osc, cam, spp, seed=51, worker_count=THREADS)[0] / spp
    b = O.render(osc, cam, spp, seed=52, worker_count=THREADS)[0] / spp
    r = A.Renderer(ps, settings(cam, spp), A.GpuOptions(seed=53, precision=precision))
    r.render(0, spp)
    g = r.read_frame(spp)
    r.close()
    noise_floor(g, a, b, f"GoldDragon 480x270x{spp} precision {precision}")


def test_c5_4k_progressive(dragon):
    """configs[4]'s shape: 3840x2160, TileProgressed every samples_per_iteration passes, TileFinished at the end, over two
    shares of the samples (device list [0, 0]); 8160 tiles per round in the reference's queue order."""
    objs, ps, osc = dragon
    cam, spp, spi = F.camera(3840, 2160), 4, 2
    st = settings(cam, spp, spi=spi)
    task = A.render_tiled(ps, st, A.GpuOptions(seed=77, device_list=[0, 0]))
    stats = task.stats()
    assert stats["samples"] == 3840 * 2160 * spp
    layout = A.tile_layout(st)
    assert len(layout) == 120 * 68
    progressed = np.zeros((2160, 3840, 3))
    finished = np.zeros((2160, 3840, 3))
    n_prog = n_fin = 0
    order = []
    while (m := task.poll()) is not None:
        t = m.tile
        if m.kind == "TileProgressed":
            assert t.sample_count == spi
            progressed[t.top:t.top + t.height, t.left:t.left + t.width] = t.data
            n_prog += 1
        else:
            assert t.sample_count == spp
            finished[t.top:t.top + t.height, t.left:t.left + t.width] = t.data
            order.append((t.left, t.top, t.width, t.height))
            n_fin += 1
    assert n_prog == n_fin == len(layout)
    assert order == [tuple(r) for r in layout]
    # the checkpoint holds the first `spi` samples of both shares; the final tiles all of them (association across shares aside)
    want_half, _ = O.render(osc, cam, spi, seed=77, worker_count=THREADS)
    want_full, cnt = O.render(osc, cam, spp, seed=77, worker_count=THREADS)
    for got, want, n in ((progressed, want_half, spi), (finished, want_full, spp)):
        rel = np.abs(got - want).max(axis=-1) / np.maximum(np.abs(want).max(axis=-1), 1e-3 * n)
        print(f"C5 4K after {n} samples: {int((rel > 1e-9).sum())} of {rel.size} pixels differ beyond 1e-9 relative (largest {float(rel.max()):.3g})")
        assert (rel > 1e-9).mean() < 1e-4, f"{(rel > 1e-9).mean():.3%} of the 4K pixels differ after {n} samples"       # measured: none
        gl, ol = (np.clip(got / n, 0, 10) @ LUMA).mean(), (np.clip(want / n, 0, 10) @ LUMA).mean()
        assert abs(gl - ol) <= 0.002 * ol
    assert stats["nonfinite_samples"] == cnt["nonfinite"]


# ------------------------------------------------------------------------------------------ C4

@pytest.mark.parametrize("box", ["cubic", "flat"])
@pytest.mark.parametrize("n", [4_000_000, 10_000_000], ids=["4M", "10M"])
def test_c4_soup_hit_indices(n, box):
    """configs[3] at 4 M and 10 M triangles: 1920x1080 pixel-centre primaries + 2^21 random rays — object index, triangle index
    and distance bits identical to the oracle.  The product's grid is built by the CUDA grid build (the host build takes seconds
    at this size) and must be the oracle's grid."""
    tris = F.triangle_soup(n, F.SOUP_BOX_CUBIC if box == "cubic" else F.SOUP_BOX_FLAT)
    og = O.AccGrid.build_from_mesh(O.Mesh.from_triangles(tris))
    osc = O.Scene()
    osc.add_grid(og, F.DRAGON_MATERIAL)
    pg = A.AccGrid.build_from_mesh(A.Mesh.new(tris), device=0)
    del tris
    oi, pi = og.info(), pg.info()
    assert oi["resolution"] == pi["resolution"] and oi["reference_count"] == pi["reference_count"]
    if box == "flat":
        assert pi["resolution"][1] > pi["resolution"][2]          # the aliased regime (A1)
    ps = A.Scene()
    ps.push_grid(pg, A.Material.from_fixture(F.DRAGON_MATERIAL))
    rays = np.concatenate([O.primary_rays(F.camera(1920, 1080)), F.random_rays(1 << 21)])
    want = osc.intersect(rays, threads=THREADS)
    got = ps.intersect(rays)
    assert_hits_equal(got, want, f"{n // 1_000_000}M soup {box}")
    assert (want[0] >= 0).mean() > 0.1
    del ps, pg, osc, og
    A.release_cached_memory()
