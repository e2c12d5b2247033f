"""Per-kernel device time of one GoldDragon render (stage timing on), for the library named by RAYMOND_CUDA_LIB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from raymond_b200 import api as A, fixtures as F
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
sc = A.Scene.from_fixture(F.gold_dragon(F.dragon_standin()))
st = A.Settings(A.CameraSettings.from_fixture(F.camera(1920, 1080)), spp)
r = A.Renderer(sc, st, A.GpuOptions(seed=1, flags=A.FLAG_STAGE_TIMING))
r.render(0, spp); r.sync()
s0 = r.stage_stats(); t0 = r.stats()["device_ms"]
r.render(0, spp); r.sync()
s1 = r.stage_stats(); t1 = r.stats()["device_ms"]
tot = {k: sum(b - a for a, b in zip(s0["ms"][k], s1["ms"][k])) for k in A.KERNEL_KINDS}
per_depth = [round(b - a, 2) for a, b in zip(s0["ms"]["traverse"], s1["ms"]["traverse"])][1:6]
print(f"total {t1 - t0:7.2f} ms | " + "  ".join(f"{k} {v:7.2f}" for k, v in tot.items()) + f" | traverse by depth {per_depth}")
