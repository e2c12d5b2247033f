"""A small end-to-end run for compute-sanitizer: mesh scene, bit-exact query, render, grid build, tonemap."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raymond_b200 import api as A, fixtures as F
objs = F.gold_dragon(F.dragon_standin(96, 24))
sc = A.Scene.from_fixture(objs)
rays = F.random_rays(5000, 3, ((-1.9, 1.9), (-0.9, 1.9), (-1.9, 4.9)))
obj, sub, t = sc.intersect(rays)
st = A.Settings(A.CameraSettings.from_fixture(F.camera(64, 36)), 3)
r = A.Renderer(sc, st, A.GpuOptions(seed=1, flags=A.FLAG_COUNT_WORK))
r.render(0, 3)
s = r.read_sums()
rgb = r.read_rgb8(3)
g = A.AccGrid.build_from_mesh(A.Mesh.new(F.bumpy_sphere(20, 40)), device=0)
print("ok", int((obj >= 0).sum()), float(s.mean()), rgb.shape, g.info()["reference_count"], r.stats()["rays"])
