"""Time tuning variants of libraymond_cuda.so (build_variants/*.so, made with raymond_b200.build.build(defines=...)).

    python scripts/tune.py [spp]          # parent: one subprocess per variant
"""
import glob, hashlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(spp):
    import numpy as np
    from raymond_b200 import api as A, fixtures as F
    sc = A.Scene.from_fixture(F.gold_dragon(F.dragon_standin()))
    st = A.Settings(A.CameraSettings.from_fixture(F.camera(1920, 1080)), spp)
    r = A.Renderer(sc, st, A.GpuOptions(seed=1, batch_spp=int(os.environ.get("RM_BATCH_SPP", "0"))))
    r.render(0, 4); r.sync(); r.clear(); r.sync()
    best = 1e9
    for _ in range(3):
        r.clear(); r.sync()
        s0 = r.stats(); r.render(0, spp); r.sync(); s1 = r.stats()
        best = min(best, s1["device_ms"] - s0["device_ms"])
    fr = r.read_sums()
    print(f"{1920*1080*spp/best/1e3:8.1f} Msamples/s  {best:7.2f} ms  sha {hashlib.sha1(fr.tobytes()).hexdigest()[:10]}", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        spp = sys.argv[1] if len(sys.argv) > 1 else "16"
        libs = [os.path.join(ROOT, "raymond_b200", "libraymond_cuda.so")] + sorted(glob.glob(os.path.join(ROOT, "build_variants", "*.so")))
        for lib in libs:
            for batch in os.environ.get("RM_BATCH_SWEEP", "0").split(","):
                env = dict(os.environ, RAYMOND_CUDA_LIB=lib, RM_BATCH_SPP=batch)
                out = subprocess.run([sys.executable, __file__, "--child", spp], env=env, capture_output=True, text=True)
                print(f"{os.path.basename(lib):28s} batch_spp={batch:3s} {out.stdout.strip() or out.stderr.strip()[-300:]}", flush=True)
