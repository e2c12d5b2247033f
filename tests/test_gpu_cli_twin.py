"""The C++ twin of the reference's front-end (examples/cli_old.cpp) drives the C ABI from compiled code: scene build,
PLY load + bake_transform + grid build, render_tiled + await, display transform, output.png."""
import os
import subprocess

import numpy as np
import pytest

from raymond_b200 import api as A
from raymond_b200 import fixtures as F

from util import settings

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cli_old_twin(tmp_path):
    from PIL import Image
    exe = str(tmp_path / "cli_old")
    lib_dir = os.path.join(ROOT, "raymond_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cli_old.cpp"),
                    "-L", lib_dir, "-lraymond_cuda", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    mesh = F.dragon_standin(160, 40)
    ply = str(tmp_path / "dragon.ply")
    F.write_ply(ply, mesh)
    out = str(tmp_path / "output.png")
    res = subprocess.run([exe, "--mesh", ply, "--out", out, "--width", "160", "--height", "90", "--spp", "6", "--seed", "11"],
                         check=True, capture_output=True, text=True)
    assert "Total render time" in res.stdout
    got = np.asarray(Image.open(out))
    # the same through the Python mirror: same library, same seed -> same bytes
    m = A.Mesh.load_ply(ply)
    m.bake_transform(F.DRAGON_TRANSLATE)
    scene = A.Scene()
    scene.push_sphere(*F.RED_SPHERE[1:3], A.Material.from_fixture(F.RED_SPHERE[3]))
    scene.push_grid(A.AccGrid.build_from_mesh(m), A.Material.from_fixture(F.DRAGON_MATERIAL))
    for p in F.BOX_PLANES:
        scene.push_plane(p[1], p[2], A.Material.from_fixture(p[3]))
    frame = A.render_tiled(scene, settings(F.camera(160, 90), 6), A.GpuOptions(seed=11)).await_()
    assert np.array_equal(got, A.tonemap(frame))
    assert (got.reshape(-1, 3).max(axis=0) > 200).all() and got.shape == (90, 160, 3)
