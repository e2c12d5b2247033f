"""The data formats either side of the path: JSON project -> Scene, tile message -> JSON, 8-bit PNG.
Host code: runs without a GPU.  (The GPU half — tracing a project-loaded scene, the tonemap kernel — is in
tests/test_gpu_formats.py.)"""
import json
import os

import numpy as np
import pytest

from raymond_b200 import api as A
from raymond_b200 import fixtures as F


def _vec(v):
    return {"x": float(v[0]), "y": float(v[1]), "z": float(v[2])}


def write_project(path, objects, mesh_dir):
    """Serialise fixture objects the way serde_json writes core::project::Project (externally tagged enums)."""
    out = []
    for i, o in enumerate(objects):
        m = o[-1]
        mat = {m[0]: [_vec(m[1]), m[2]]} if m[0] != "Emission" else {"Emission": [_vec(m[1]), _vec(m[2]), m[3], m[4]]}
        if o[0] == "sphere":
            geo = {"Sphere": {"origin": _vec(o[1]), "radius": o[2]}}
        elif o[0] == "plane":
            geo = {"Plane": {"origin": _vec(o[1]), "normal": _vec(o[2])}}
        else:
            ply = os.path.join(mesh_dir, f"mesh{i}.ply")
            F.write_ply(ply, o[1])
            geo = {"Mesh": ply}
        out.append({"geometry": geo, "material": mat})
    with open(path, "w") as f:
        json.dump({"objects": out}, f)


def test_project_loader(tmp_path):
    objs = F.gold_dragon(F.dragon_standin(48, 12))
    path = str(tmp_path / "scene.json")
    write_project(path, objs, str(tmp_path))
    s = A.Scene.load_project(path)
    assert len(s) == len(objs) == 8


def test_project_loader_errors(tmp_path):
    def load(text):
        p = tmp_path / "p.json"
        p.write_text(text)
        return A.Scene.load_project(str(p))

    with pytest.raises(A.RaymondError) as e:
        A.Scene.load_project(str(tmp_path / "missing.json"))
    assert e.value.status == A.RM_ERR_IO
    for bad in ('{"objects": [', '{"things": []}', '{"objects":[{"geometry":{"Cone":{}},"material":{"Diffuse":[{"x":1,"y":1,"z":1},0.5]}}]}',
                '{"objects":[{"geometry":{"Sphere":{"origin":{"x":0,"y":0,"z":0}}},"material":{"Diffuse":[{"x":1,"y":1,"z":1},0.5]}}]}',
                '{"objects":[{"geometry":{"Sphere":{"origin":{"x":0,"y":0,"z":0},"radius":1}},"material":{"Metal":[0.5]}}]}'):
        with pytest.raises(A.RaymondError) as e:
            load(bad)
        assert e.value.status == A.RM_ERR_PROJECT, bad
    # a Mesh that cannot be read is the PLY loader's IO error (Mesh::load_ply unwraps the read, mesh.rs:59)
    with pytest.raises(A.RaymondError) as e:
        load('{"objects":[{"geometry":{"Mesh":"/nonexistent/x.ply"},"material":{"Metal":[{"x":1,"y":1,"z":0.1},0.15]}}]}')
    assert e.value.status == A.RM_ERR_IO
    assert len(load('{"objects": []}')) == 0
    assert len(load(' {"objects":[{"material":{"Emission":[{"x":1.5,"y":1.5,"z":1.5},{"x":1,"y":1,"z":1},0.27,0]},'
                    '"geometry":{"Plane":{"origin":{"x":0,"y":2e0,"z":0},"normal":{"x":0,"y":-1,"z":0}}}}]}\n')) == 1


def test_tile_message_json_round_trip():
    rng = np.random.default_rng(5)
    data = rng.random((3, 5, 3)) * 100
    data[0, 0] = (0.0, 1.0, 1e-300)
    data[1, 2] = (np.nan, np.inf, -2.5e17)
    tile = A.Tile(7, 5, 3, 32, 64, data)
    for kind in ("TileProgressed", "TileFinished"):
        text = A.message_to_json(kind, tile)
        msg = json.loads(text)
        assert list(msg) == ["type", "data"] and msg["type"] == kind           # serde tag = "type", content = "data"
        d = msg["data"]
        assert list(d) == ["sample_count", "width", "height", "left", "top", "data"]   # Tile's field order, core/src/tile.rs:7-14
        assert (d["sample_count"], d["width"], d["height"], d["left"], d["top"]) == (7, 5, 3, 32, 64)
        got = np.array([[np.nan if v is None else v for v in (p["x"], p["y"], p["z"])] for p in d["data"]], dtype=np.float64).reshape(3, 5, 3)
        finite = np.isfinite(data)
        assert np.array_equal(got[finite].view(np.uint64), data[finite].view(np.uint64))      # f64 text round-trips exactly
        assert np.isnan(got[~finite]).all()                                                     # serde_json writes non-finite as null
    assert '"x":0.0,"y":1.0' in A.message_to_json("TileFinished", tile)                         # floats keep a float mark, like serde_json


def test_png_writer(tmp_path):
    from PIL import Image
    img = (np.random.default_rng(1).random((77, 131, 3)) * 256).astype(np.uint8)
    path = str(tmp_path / "out.png")
    A.write_png(path, img)
    back = Image.open(path)
    assert back.mode == "RGB" and back.size == (131, 77)
    assert np.array_equal(np.asarray(back), img)
    with pytest.raises(A.RaymondError) as e:
        A.write_png(str(tmp_path / "no" / "dir.png"), img)
    assert e.value.status == A.RM_ERR_IO
