"""BASELINE.json configs[3]: synthetic triangle soup (1M-10M tris), primary-ray hit-index microbench through the C ABI.

    python scripts/c4_soup.py [--sizes 1000000,4000000,10000000] [--no-check]

Per size and box (B1 cubic grid, B2 flat/aliased grid): host grid build time, scene upload, device-resident
Scene::intersect over the 1920x1080 pixel-centre primaries (rm_primary_rays_device + rm_device_scene_intersect), timed with
CUDA events (3 warm-ups, best of 5; the soup's traversal set is far larger than L2 at 4M/10M), and a bit-exact check of
object index, triangle index and distance bits against the oracle on the same rays (+ 2^21 random rays at 1M).
One JSON line per case on stdout."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from raymond_b200 import api as A, fixtures as F

sizes = [1_000_000, 4_000_000, 10_000_000]
if "--sizes" in sys.argv:
    sizes = [int(x) for x in sys.argv[sys.argv.index("--sizes") + 1].split(",")]
check = "--no-check" not in sys.argv
cam = F.camera(1920, 1080)
cs = A.CameraSettings.from_fixture(cam)
n = 1920 * 1080
stream = torch.cuda.current_stream().cuda_stream
for N in sizes:
    for box_name, box in (("B1-cubic", F.SOUP_BOX_CUBIC), ("B2-flat", F.SOUP_BOX_FLAT)):
        tris = F.triangle_soup(N, box)
        t0 = time.perf_counter()
        scene = A.Scene.from_fixture(F.soup_scene(tris))
        build_s = time.perf_counter() - t0
        info = scene._grids[0].info()
        # the same grid built by the CUDA kernels (rm_grid_build_on_device), timed warm
        dev_build_s, same = None, None
        for _ in range(2):
            m = A.Mesh.new(tris)
            t0 = time.perf_counter()
            dg = A.AccGrid.build_from_mesh(m, device=0)
            dev_build_s = time.perf_counter() - t0
        hs, hr = scene._grids[0].cells()
        ds_, dr = dg.cells()
        same = bool(np.array_equal(hs, ds_) and np.array_equal(hr, dr))
        del dg, m, hs, hr, ds_, dr
        t0 = time.perf_counter()
        ds = A.DeviceScene(scene, 0)
        upload_s = time.perf_counter() - t0
        rays = torch.empty((n, 6), dtype=torch.float64, device="cuda")
        obj = torch.empty(n, dtype=torch.int64, device="cuda")
        sub = torch.empty(n, dtype=torch.int64, device="cuda")
        dist = torch.zeros(n, dtype=torch.float64, device="cuda")
        A.primary_rays_device(cs, 0, rays.data_ptr(), stream)
        for _ in range(3):
            ds.intersect_device(rays.data_ptr(), n, obj.data_ptr(), sub.data_ptr(), dist.data_ptr(), stream)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ds.intersect_device(rays.data_ptr(), n, obj.data_ptr(), sub.data_ptr(), dist.data_ptr(), stream)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        line = {"config": "C4 triangle soup", "triangles": N, "box": box_name, "resolution": info["resolution"], "cells": info["cell_count"],
                "references": info["reference_count"], "host_grid_build_s": round(build_s, 2), "device_grid_build_s": round(dev_build_s, 3), "device_grid_identical": same, "upload_s": round(upload_s, 3), "rays": n,
                "ms": best, "mrays_per_s": n / best / 1e3, "hit_fraction": float((obj >= 0).float().mean())}
        if check:
            from oracle import oracle as O
            osc = O.Scene()
            osc.add_grid(O.AccGrid.build_from_mesh(O.Mesh.from_triangles(tris)), F.DRAGON_MATERIAL)
            hr = rays.cpu().numpy()
            t0 = time.perf_counter()
            wobj, wsub, wt, cnt = osc.intersect(hr, threads=os.cpu_count() or 8)
            cpu_s = time.perf_counter() - t0
            gobj, gsub, gt = obj.cpu().numpy(), sub.cpu().numpy().astype(np.uint64), dist.cpu().numpy()
            hit = wobj >= 0
            ok = bool(np.array_equal(gobj, wobj) and np.array_equal(gsub[hit], wsub[hit]) and np.array_equal(gt[hit].view(np.uint64), wt[hit].view(np.uint64)))
            line.update({"bit_exact_vs_oracle": ok, "oracle_mrays_per_s": n / cpu_s / 1e6, "oracle_threads": os.cpu_count(),
                         "cells_per_ray": cnt["cells"] / n, "tests_per_ray": cnt["tri_tests"] / n,
                         "alg_bytes_per_ray": 64 + 8 * cnt["cells"] / n + 76 * cnt["tri_tests"] / n})
            line["alg_gbs"] = line["alg_bytes_per_ray"] * n / (best * 1e-3) / 1e9
            del osc
        print(json.dumps(line), flush=True)
        del ds, scene, tris
