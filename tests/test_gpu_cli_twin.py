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


def test_cli_old_twin_streams_tile_messages(tmp_path):
    """--serve: the reference's intended server role (server/src/main.rs:174-192, protocol.rs:9-14) — every Message of a
    progressive render on a TCP socket, one JSON text per "\\r\\n"-terminated line, in the order TaskHandle::poll delivers them."""
    import json
    import socket
    import time
    exe = str(tmp_path / "cli_old")
    lib_dir = os.path.join(ROOT, "raymond_b200")
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "cli_old.cpp"),
                    "-L", lib_dir, "-lraymond_cuda", f"-Wl,-rpath,{lib_dir}", "-o", exe], check=True)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    W, H, spp, spi, seed = 96, 64, 6, 2, 4
    out = str(tmp_path / "served.png")
    proc = subprocess.Popen([exe, "--out", out, "--width", str(W), "--height", str(H), "--spp", str(spp), "--spi", str(spi), "--seed", str(seed),
                             "--serve", str(port)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    conn = None
    for _ in range(200):
        try:
            conn = socket.create_connection(("127.0.0.1", port), timeout=60)
            break
        except OSError:
            time.sleep(0.05)
    assert conn is not None, "the twin never listened"
    data = b""
    while chunk := conn.recv(1 << 20):
        data += chunk
    conn.close()
    assert proc.wait(timeout=120) == 0, proc.stderr.read()
    lines = data.split(b"\r\n")
    assert lines[-1] == b""
    msgs = [json.loads(l) for l in lines[:-1]]
    st = settings(F.camera(W, H), spp, spi=spi)
    layout = [tuple(r) for r in A.tile_layout(st)]
    assert [m["type"] for m in msgs] == ["TileProgressed"] * (2 * len(layout)) + ["TileFinished"] * len(layout)
    assert [m["data"]["sample_count"] for m in msgs] == [2] * len(layout) + [4] * len(layout) + [6] * len(layout)
    assert [(m["data"]["left"], m["data"]["top"], m["data"]["width"], m["data"]["height"]) for m in msgs[-len(layout):]] == layout
    assert list(msgs[0]["data"].keys()) == ["sample_count", "width", "height", "left", "top", "data"]       # core/src/tile.rs:6-14 field order
    # the tiles on the wire are the running sums of the same render through the Python mirror (exact f64 round trip)
    r = A.Renderer(A.Scene.from_fixture(F.reflective_spheres()), st, A.GpuOptions(seed=seed))
    r.render(0, spp)
    want = r.read_sums()
    r.close()
    for m in msgs[-len(layout):]:
        t = m["data"]
        got = np.array([[p["x"], p["y"], p["z"]] for p in t["data"]]).reshape(t["height"], t["width"], 3)
        assert np.array_equal(got, want[t["top"]:t["top"] + t["height"], t["left"]:t["left"] + t["width"]])
    assert os.path.exists(out)
