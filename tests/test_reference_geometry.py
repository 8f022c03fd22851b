"""The oracle's geometry (and, through the committed vectors, the CUDA path) against the
REFERENCE ITSELF: src/surface.cpp, cell.cpp, lattice.cpp, universe.cpp, geometry.cpp,
geometry_aux.cpp compiled unmodified from /root/reference into
oracle/_ref/libref_geom.so (oracle/Makefile), fed a geometry.xml text of each
BASELINE.json configuration.  Cell, instance, material, coordinate depth, crossing
surface token, level and lattice translation must be bit-exact; so must the distance."""
import os

import numpy as np
import pytest

import ref_geometry
import reference_cases as rc
from oracle.binding import Oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not os.path.exists(ref_geometry.LIB),
                               reason="oracle/_ref/libref_geom.so not built (no /root/reference here)")


def _compare(model, found, answers, dist, xyz, uvw, engine):
    cid, inst, mat, nco = found
    d, s, lv, lt = dist
    c2, i2, m2, n2 = engine.find_cell(xyz, uvw)
    ids = np.asarray(model["cell_id"])
    assert np.array_equal(np.where(c2 >= 0, ids[np.clip(c2, 0, None)], -1), cid)
    assert np.array_equal(i2, inst) and np.array_equal(m2, mat) and np.array_equal(n2, nco)
    d2, s2, lv2, lt2 = engine.distance_to_boundary(xyz, uvw)
    assert np.array_equal(s2, s) and np.array_equal(lv2, lv) and np.array_equal(lt2, lt)
    return d2, d


@needs_ref
@pytest.mark.parametrize("name", rc.NAMES)
def test_oracle_geometry_equals_reference(name):
    m, model, box = rc.build(name)
    ref = ref_geometry.ReferenceGeometry(m)
    assert ref.n_surfaces == int(model["n_surfaces"]) and ref.n_cells == int(model["n_cells"])
    assert ref.n_universes == int(model["n_universes"]) and ref.n_lattices == int(model["n_lattices"])
    xyz, uvw = rc.probe_points(box, 20000)
    wx, wu = rc.walk(ref, xyz[:3000], uvw[:3000])
    xyz, uvw = np.concatenate([xyz, wx]), np.concatenate([uvw, wu])
    found, dist = ref.find_cell(xyz, uvw), ref.distance_to_boundary(xyz, uvw)
    assert np.mean(found[0] >= 0) > 0.4   # the rest probes "no cell here" (outside / past a boundary)
    d2, d = _compare(model, found, None, dist, xyz, uvw, Oracle(model))
    assert np.array_equal(d2, d)               # same arithmetic, same libm: identical doubles


@needs_ref
@pytest.mark.parametrize("name", ["assembly_ce", "c5g7_core", "pwr_core_ce", "rotated_fills", "surface_zoo"])
def test_tracking_through_the_geometry_equals_reference(name):
    """straight-line tracks driven by the reference's own crossing logic (cross_lattice, the
    neighbour-list search with its exhaustive fallback): after every crossing the oracle is in
    the same cell, material and depth, and flies the identical distance"""
    m, model, box = rc.build(name)
    ref = ref_geometry.ReferenceGeometry(m)
    xyz, uvw = rc.probe_points(box, 4000, seed=41)
    steps = 30
    cr, mr, nr, dr = ref.track(xyz, uvw, steps)
    co, mo, no, do = Oracle(model).geom_track(xyz, uvw, steps)
    assert np.count_nonzero(cr >= 0) > 2000
    assert np.array_equal(co, cr) and np.array_equal(mo, mr) and np.array_equal(no, nr)
    assert np.array_equal(do, dr)


@needs_ref
def test_reflective_boundary_equals_reference():
    """ReflectiveBC::handle_particle of the reference on every surface kind (planes, general
    plane, cylinders, spheres, cones made reflective for the purpose): identical new directions"""
    import ctypes as C
    import copy
    from openmc_b200 import examples, mgxs
    m, lib_, box, of = examples.surface_zoo()
    m = copy.deepcopy(m)
    for s in m.surfaces:
        s.boundary = "reflective"
    model = mgxs.build_mg_model(m, lib_)
    ref = ref_geometry.ReferenceGeometry(m)
    o = Oracle(model); o.load()
    g = np.random.default_rng(8)
    n = 9000
    xyz = np.ascontiguousarray(g.uniform(-8, 8, size=(n, 3)))
    uvw = g.normal(size=(n, 3)); uvw /= np.linalg.norm(uvw, axis=1)[:, None]
    uvw = np.ascontiguousarray(uvw)
    tok = ((np.arange(n) % len(m.surfaces)) + 1).astype(np.int32) * g.choice([-1, 1], n).astype(np.int32)
    want = ref.reflect(tok, xyz, uvw)
    got = np.zeros_like(uvw)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert o.lib.orc_reflect(C.c_int64(n), p(np.ascontiguousarray(tok)), p(xyz), p(uvw), p(got)) == 0
    assert np.array_equal(got, want)
    assert not np.array_equal(want, uvw)


@needs_ref
def test_periodic_boundary_equals_reference():
    """TranslationalPeriodicBC::handle_particle of the reference (pairs given by id, the two
    unpaired surfaces paired up, general planes with a non-unit normal): the flattener's partner table and
    translation vectors move a particle to the identical position and surface token"""
    m, model, box = rc.build("periodic_box")
    ref = ref_geometry.ReferenceGeometry(m)
    g = np.random.default_rng(2)
    n = 5000
    lo, hi = np.array(box[:3]), np.array(box[3:])
    xyz = lo + g.uniform(size=(n, 3)) * (hi - lo)
    uvw = g.normal(size=(n, 3)); uvw /= np.linalg.norm(uvw, axis=1)[:, None]
    tok = g.choice([2, -2, 3, -3, 4, -4, 5, -5, 6, 1], size=n).astype(np.int32)
    new_xyz, new_tok = ref.periodic(tok, xyz, uvw)
    part = np.asarray(model["surf_periodic"])
    sh = np.asarray(model["surf_periodic_shift"]).reshape(-1, 3)
    s = np.abs(tok) - 1
    assert np.array_equal(new_tok, np.where(part[s] >= 0, np.sign(tok) * (part[s] + 1), 0))
    assert np.array_equal(new_xyz, np.where((part[s] >= 0)[:, None], xyz + sh[s], xyz))
    assert set(np.unique(new_tok)) == {-5, -4, -3, -2, 0, 2, 3, 4, 5}


@pytest.mark.parametrize("name", rc.NAMES)
def test_committed_reference_vectors_are_reproduced_by_the_oracle(name):
    """the vectors the GPU tier uses (tests/golden/refgeom_*.npz, written by
    tests/golden/make_reference_geometry.py from the reference library)"""
    m, model, box = rc.build(name)
    v = np.load(os.path.join(GOLDEN, f"refgeom_{name}.npz"))
    found = (v["cell_id"], v["instance"], v["material"], v["n_coord"])
    dist = (v["distance"], v["surface"], v["level"], v["lattice_trans"])
    d2, d = _compare(model, found, None, dist, v["xyz"], v["uvw"], Oracle(model))
    assert np.array_equal(d2, d)


@pytest.mark.parametrize("name", rc.NAMES)
def test_engine_sources_equal_reference_vectors(emu_lib, name):
    """the DEVICE functions (openmc_b200/csrc/device.h, host-compiled for the tests) against
    the reference vectors -- same libm on the host: distances identical too"""
    from openmc_b200 import lib
    m, model, box = rc.build(name)
    v = np.load(os.path.join(GOLDEN, f"refgeom_{name}.npz"))
    eng = lib.Engine(emu_lib)
    eng.init(model, 0, 65536)
    found = (v["cell_id"], v["instance"], v["material"], v["n_coord"])
    dist = (v["distance"], v["surface"], v["level"], v["lattice_trans"])
    d2, d = _compare(model, found, None, dist, v["xyz"], v["uvw"], eng)
    assert np.array_equal(d2, d)
    eng.finalize()


@pytest.mark.gpu
@pytest.mark.parametrize("name", rc.NAMES)
def test_gpu_geometry_equals_reference_vectors(gpu_engine, name):
    """CUDA path vs vectors computed by the reference's own geometry code: indices
    bit-exact, distances to 1e-12 (device sqrt / division are IEEE, -fmad=false)"""
    m, model, box = rc.build(name)
    v = np.load(os.path.join(GOLDEN, f"refgeom_{name}.npz"))
    gpu_engine.init(model, 0, 65536)
    found = (v["cell_id"], v["instance"], v["material"], v["n_coord"])
    dist = (v["distance"], v["surface"], v["level"], v["lattice_trans"])
    d2, d = _compare(model, found, None, dist, v["xyz"], v["uvw"], gpu_engine)
    np.testing.assert_allclose(d2, d, rtol=1e-12)
    gpu_engine.finalize()
