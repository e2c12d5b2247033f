"""The HBM-bound regime: Scene::intersect on the 10 M-triangle soup, whose traversal set (positions 960 MB, spheres
+ references 1 GB, cells 240 MB) is far larger than the 126 MB L2.

    python scripts/c4_profile.py [triangles] [box: cubic|flat]

Two ray sets through rm_device_scene_intersect, timed with CUDA events (3 warm-ups, best of 5):
  primaries   the 1920x1080 pixel-centre rays of configs[3] (2 073 600 rays: one short launch, the persistent kernel's fill and
              drain are a large part of it)
  x16         16 jittered rays per pixel (33 177 600 rays: steady state)
Run under `ncu -k regex:k_traverse` to capture the two traversal launches (the 4th and the last one)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from raymond_b200 import api as A, fixtures as F

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
box = F.SOUP_BOX_FLAT if (len(sys.argv) > 2 and sys.argv[2] == "flat") else F.SOUP_BOX_CUBIC
quick = "--quick" in sys.argv          # one warm-up, one timed call per ray set (for ncu)
tris = F.triangle_soup(N, box)
t0 = time.perf_counter()
grid = A.AccGrid.build_from_mesh(A.Mesh.new(tris), device=0)
build_s = time.perf_counter() - t0
del tris
scene = A.Scene()
scene.push_grid(grid, A.Material.from_fixture(F.DRAGON_MATERIAL))
info = grid.info()
ds = A.DeviceScene(scene, 0)
cam = F.camera(1920, 1080)
cs = A.CameraSettings.from_fixture(cam)
W, H = 1920, 1080
stream = torch.cuda.current_stream().cuda_stream
prim = torch.empty((W * H, 6), dtype=torch.float64, device="cuda")
A.primary_rays_device(cs, 0, prim.data_ptr(), stream)
# 16 jittered rays per pixel: the pixel-centre direction moved inside the pixel's footprint, re-normalised (timing only)
g = torch.Generator(device="cuda"); g.manual_seed(1)
tan_half = np.tan(55.0 / 2 * np.pi / 180.0)
px = 2.0 * tan_half / H
many = prim.repeat(16, 1)
many[:, 3:5] += (torch.rand((many.shape[0], 2), generator=g, device="cuda", dtype=torch.float64) - 0.5) * px * many[:, 5:6]
many[:, 3:6] /= many[:, 3:6].norm(dim=1, keepdim=True)
out = {}
for name, rays in (("primaries", prim), ("x16", many)):
    n = rays.shape[0]
    obj = torch.empty(n, dtype=torch.int64, device="cuda")
    sub = torch.empty(n, dtype=torch.int64, device="cuda")
    dist = torch.zeros(n, dtype=torch.float64, device="cuda")
    for _ in range(1 if quick else 3):
        ds.intersect_device(rays.data_ptr(), n, obj.data_ptr(), sub.data_ptr(), dist.data_ptr(), stream)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(1 if quick else 5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ds.intersect_device(rays.data_ptr(), n, obj.data_ptr(), sub.data_ptr(), dist.data_ptr(), stream)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    out[name] = {"rays": n, "ms": best, "mrays_per_s": n / best / 1e3, "hit_fraction": float((obj >= 0).float().mean())}
print(json.dumps({"config": "C4 soup, HBM regime", "triangles": N, "box": "flat" if box is F.SOUP_BOX_FLAT else "cubic", "resolution": info["resolution"],
                  "references": info["reference_count"], "device_grid_build_s": round(build_s, 3), **out}))
