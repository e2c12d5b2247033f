"""GPU parity of the device grid build (rm_grid_build_on_device) — AccGrid::build_from_mesh, acc_grid.rs:36-83.

The grid decides which triangle a ray reports, so the bar is identity with the oracle's grid: resolution, cell-size
bits, and every cell's triangle list in the same (ascending) order; the reference's panics are the same statuses."""
import numpy as np
import pytest

from oracle import oracle as O
from raymond_b200 import api as A
from raymond_b200 import fixtures as F

from util import assert_hits_equal

pytestmark = pytest.mark.gpu

MESHES = {
    "cube": F.cube,
    "bumpy": F.bumpy_sphere,
    "bumpy_fine_aliased": lambda: F.bumpy_sphere(60, 120, 1.0, 0.1, (1.0, 0.7, 0.45)),
    "tube": lambda: F.dragon_standin(96, 24),
    "soup_cubic_200k": lambda: F.triangle_soup(200_000, F.SOUP_BOX_CUBIC),
    "soup_flat_200k": lambda: F.triangle_soup(200_000, F.SOUP_BOX_FLAT),
    "dragon_standin": F.dragon_standin,
    "one_fat_cell": lambda: np.concatenate([F.cube()] * 40),          # 480 references in every cell: the heap-sort path
}


@pytest.mark.parametrize("name", list(MESHES))
def test_device_grid_is_the_oracles(name):
    tris = MESHES[name]()
    og = O.AccGrid.build_from_mesh(O.Mesh.from_triangles(tris))
    dg = A.AccGrid.build_from_mesh(A.Mesh.new(tris), device=0)
    oi, di = og.info(), dg.info()
    assert oi["resolution"] == di["resolution"]
    assert np.array_equal(oi["cell_size"].view(np.uint64), di["cell_size"].view(np.uint64))
    assert (oi["cell_count"], oi["reference_count"], oi["triangle_count"]) == (di["cell_count"], di["reference_count"], di["triangle_count"])
    ostart, orefs = og.csr()
    dstart, drefs = dg.cells()
    assert np.array_equal(ostart, dstart)
    assert np.array_equal(orefs, drefs)


def test_device_grid_errors_are_the_host_builds():
    flat = F.make_triangles(np.array([[0.0, 0, 0]]), np.array([[1.0, 0, 0]]), np.array([[0.0, 1, 0]]))
    deep = F.bumpy_sphere(20, 40, 1.0, 0.05, (0.5, 0.45, 1.6))
    for tris, status in ((flat, A.RM_ERR_DEGENERATE_BOUNDS), (deep, A.RM_ERR_GRID_INDEX_OOB)):
        with pytest.raises(A.RaymondError) as host:
            A.AccGrid.build_from_mesh(A.Mesh.new(tris))
        with pytest.raises(A.RaymondError) as dev:
            A.AccGrid.build_from_mesh(A.Mesh.new(tris), device=0)
        assert host.value.status == dev.value.status == status
        if status == A.RM_ERR_GRID_INDEX_OOB:
            assert str(host.value) == str(dev.value)          # names the same first offending triangle
    with pytest.raises(A.RaymondError) as e:
        A.AccGrid.build_from_mesh(A.Mesh.new(F.cube()), device=99)
    assert e.value.status == A.RM_ERR_CUDA


def test_scene_on_a_device_built_grid_traces_the_same():
    tris = F.translate(F.bumpy_sphere(60, 120, 1.0, 0.1, (1.0, 0.7, 0.45)), (0.1, -0.2, 3.0))
    s = A.Scene()
    s.push_grid(A.AccGrid.build_from_mesh(A.Mesh.new(tris), device=0), A.Material.Metal((1, 1, 0.1), 0.15))
    osc = O.Scene()
    osc.add_grid(O.AccGrid.build_from_mesh(O.Mesh.from_triangles(tris)), F.DRAGON_MATERIAL)
    rays = np.concatenate([O.primary_rays(F.camera(160, 120)), F.random_rays(20000)])
    assert_hits_equal(s.intersect(rays), osc.intersect(rays), "device-built grid")
