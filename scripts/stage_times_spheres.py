"""Per-kernel device time of one ReflectiveSpheres (+DoF) render (stage timing on)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raymond_b200 import api as A, fixtures as F
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sc = A.Scene.from_fixture(F.reflective_spheres())
for name, cam in (("sharp", F.camera(1920, 1080)), ("dof", F.camera(1920, 1080, focal_length=2.5, aperture_radius=0.5))):
    st = A.Settings(A.CameraSettings.from_fixture(cam), spp)
    r = A.Renderer(sc, st, A.GpuOptions(seed=1, flags=A.FLAG_STAGE_TIMING))
    r.render(0, spp); r.sync()
    s0 = r.stage_stats(); t0 = r.stats()["device_ms"]
    r.render(0, spp); r.sync()
    s1 = r.stage_stats(); t1 = r.stats()["device_ms"]
    tot = {k: sum(b - a for a, b in zip(s0["ms"][k], s1["ms"][k])) for k in A.KERNEL_KINDS}
    sd = [round(b - a, 2) for a, b in zip(s0["ms"]["setup"], s1["ms"]["setup"])][1:6]
    sh = [round(b - a, 2) for a, b in zip(s0["ms"]["shade"], s1["ms"]["shade"])][1:6]
    print(f"{name}: total {t1 - t0:7.2f} ms ({1920*1080*spp/(t1-t0)/1e3:.0f} Msamples/s) | " + "  ".join(f"{k} {v:6.2f}" for k, v in tot.items()) + f" | setup {sd} shade {sh}")
    r.close()
