"""Where the end-to-end time of render_tiled(...).await() goes (GoldDragon stand-in, 1920x1080)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raymond_b200 import api as A, fixtures as F

spp = int(sys.argv[1]) if len(sys.argv) > 1 else 64
# argv[2]: device list of the render_tiled leg, e.g. "0,0,0,0" (four shares on GPU 0) or "0,1,2,3,4,5,6,7"
devices = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else None
objs = F.gold_dragon(F.dragon_standin())
cam = F.camera(1920, 1080)
t = time.perf_counter(); scene = A.Scene.from_fixture(objs); print(f"host scene build {time.perf_counter()-t:.3f}s")
st = A.Settings(A.CameraSettings.from_fixture(cam), spp)
for rep in range(3):
    t0 = time.perf_counter(); ds = A.DeviceScene(scene, 0); t1 = time.perf_counter()
    r = A.Renderer(ds, st, A.GpuOptions(seed=rep)); t2 = time.perf_counter()
    r.render(0, spp); r.sync(); t3 = time.perf_counter()
    out = r.read_frame(spp); t4 = time.perf_counter()
    r.close(); t5 = time.perf_counter()
    del ds; t6 = time.perf_counter()
    print(f"rep {rep}: device scene {t1-t0:.3f}  renderer create {t2-t1:.3f}  render {t3-t2:.3f}  read_frame {t4-t3:.3f}  renderer close {t5-t4:.3f}  scene destroy {t6-t5:.3f}")
for rep in range(3):
    t0 = time.perf_counter(); task = A.render_tiled(scene, st, A.GpuOptions(seed=rep, device_list=devices)); t1 = time.perf_counter()
    s = task.stats(); t2 = time.perf_counter()
    out = task.await_(); t3 = time.perf_counter()
    del task; t4 = time.perf_counter()
    print(f"render_tiled rep {rep}: call {t1-t0:.3f}  until finished {t2-t1:.3f} (device {s['device_ms']:.0f} ms)  await {t3-t2:.3f}  destroy {t4-t3:.3f}  total {t4-t0:.3f}")
