"""How much do several wavefront streams on ONE GPU overlap?  render_tiled with a device list that repeats ordinal 0
(every share renders its samples on its own stream with its own queues) against the single-stream render of the same frame.
    python scripts/concurrency_probe.py [spp]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raymond_b200 import api as A, fixtures as F
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 500
scene = A.Scene.from_fixture(F.gold_dragon(F.dragon_standin()))
st = A.Settings(A.CameraSettings.from_fixture(F.camera(1920, 1080)), spp)
for shares, batch in ((1, 0), (1, 16), (2, 32), (2, 16), (2, 8), (3, 16), (3, 8), (4, 16), (4, 8), (4, 4), (6, 8), (8, 4)):
    best = None
    for rep in range(3):
        t0 = time.perf_counter()
        task = A.render_tiled(scene, st, A.GpuOptions(seed=rep, device_list=[0] * shares, batch_spp=batch))
        s = task.stats()
        dt = time.perf_counter() - t0
        del task
        if rep:
            best = dt if best is None else min(best, dt)
    print(f"shares {shares} batch_spp {batch:2d}: {best*1e3:7.1f} ms until finished = {1920*1080*spp/best/1e6:7.1f} Msamples/s (device_ms max {s['device_ms']:.1f})", flush=True)
