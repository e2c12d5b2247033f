"""Quick device-side throughput probe (not the bench): renders a few spp of the benchmark scenes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from raymond_b200 import api as A, fixtures as F

def run(name, objs, cam, spp, batch=0):
    t0 = time.time()
    sc = A.Scene.from_fixture(objs)
    t1 = time.time()
    st = A.Settings(A.CameraSettings.from_fixture(cam), spp)
    r = A.Renderer(sc, st, A.GpuOptions(seed=1, batch_spp=batch))
    t2 = time.time()
    r.render(0, min(spp, 2)); r.sync(); r.clear(); r.sync()
    s0 = r.stats()
    t3 = time.time()
    r.render(0, spp); r.sync()
    t4 = time.time()
    s = r.stats()
    ms = s['device_ms'] - s0['device_ms']
    n = cam['width'] * cam['height'] * spp
    print(f"{name}: {n/ms/1e3:.1f} Msamples/s  {(s['rays'])/ms/1e3:.1f} Mrays/s  device {ms:.1f} ms wall {1e3*(t4-t3):.1f} ms  rays/path {s['rays']/s['samples']:.2f} nonfinite {s['nonfinite_samples']} launches {s['kernel_launches']} host-build {t1-t0:.2f}s upload+init {t2-t1:.2f}s", flush=True)
    fr = r.read_frame(spp)
    print("   mean", fr.mean(axis=(0, 1)))
    r.close()

if __name__ == "__main__":
    spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    run("ReflectiveSpheres 1920x1080", F.reflective_spheres(), F.camera(1920, 1080), spp)
    run("ReflectiveSpheres+DoF 1920x1080", F.reflective_spheres(), F.camera(1920, 1080, aperture_radius=0.5), spp)
    tris = F.dragon_standin()
    run("GoldDragon(stand-in) 1920x1080", F.gold_dragon(tris), F.camera(1920, 1080), spp)
    run("GoldDragon(stand-in) 1920x1080 batch1", F.gold_dragon(tris), F.camera(1920, 1080), spp, batch=1)
