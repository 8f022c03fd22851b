"""Synthetic inputs for the BASELINE.json configs and for the reference's
self-contained multigroup regression tests (the known-answer tests that pin the
path without a nuclear-data download).

Each builder returns ``(Model, library, source_box, only_fissionable)``; the numbers are the
literal inputs of the cited reference files.
"""
from __future__ import annotations

import numpy as np

from .mgxs import MGXSLibrary, XSdata
from .model import Filter, Mesh, Model, Tally


# ----------------------------------------------------------------------------
# tests/regression_tests/mg_*/test.py : 2-group slab library
# ----------------------------------------------------------------------------
def slab_2g_library(order=1, names=("mat_1",)):
    """create_library() of tests/regression_tests/mg_legendre/test.py:12-41
    (order=1) and mg_max_order/test.py:12-41 (order=2)."""
    lib = MGXSLibrary([0.0, 0.625, 20.0e6])
    nu = [2.50, 2.50]
    fiss = np.array([0.002817, 0.097])
    capture = [0.008708, 0.02518]
    absorption = np.add(capture, fiss)
    if order == 1:
        scatter = np.array([[[0.31980, 0.06694], [0.004555, -0.0003972]],
                            [[0.00000, 0.00000], [0.424100, 0.05439000]]])
    else:
        scatter = np.array([[[0.31980, 0.06694, 0.003], [0.004555, -0.0003972, 0.00002]],
                            [[0.00000, 0.00000, 0.000], [0.424100, 0.05439000, 0.0025]]])
    total = [0.33588, 0.54628]
    chi = [1., 0.]
    for name in names:
        x = XSdata(name, 2)
        x.order = order
        x.set_nu_fission(np.multiply(nu, fiss))
        x.set_absorption(absorption)
        x.set_scatter_matrix(scatter)
        x.set_total(total)
        x.set_chi(chi)
        lib.add_xsdata(x)
    return lib


def mg_basic():
    """tests/regression_tests/mg_basic/test.py:12-95: six slab regions holding the
    same 2-group data as Legendre, 33-point tabular, 33-bin histogram, fission
    matrix, angle-dependent (2 x 2) and a microscopic combination."""
    lib = slab_2g_library(1, names=("mat_1", "mat_6"))
    mat_1 = lib.xsdatas["mat_1"]
    mat_2 = mat_1.convert_scatter_format("tabular", 33)
    mat_2.name = "mat_2"
    lib.add_xsdata(mat_2)
    mat_3 = mat_1.convert_scatter_format("histogram", 33)
    mat_3.name = "mat_3"
    lib.add_xsdata(mat_3)
    nu = [2.50, 2.50]
    fiss = np.array([0.002817, 0.097])
    mat_4 = XSdata("mat_4", 2)
    mat_4.order = 1
    mat_4.set_nu_fission(np.outer(np.multiply(nu, fiss), [1., 0.]))
    mat_4.set_absorption(mat_1.data[294.0]["absorption"])
    mat_4.set_scatter_matrix(mat_1.data[294.0]["scatter_matrix"])
    mat_4.set_total(mat_1.data[294.0]["total"])
    lib.add_xsdata(mat_4)
    mat_5 = mat_1.convert_representation("angle", 2, 2)
    mat_5.name = "mat_5"
    lib.add_xsdata(mat_5)
    names = ["base leg", "base tab", "base hist", "base matrix", "base ang", "micro"]
    m, box, of = slab_mg(num_regions=6, mat_names=names)
    micro = m.materials[-1]
    micro.macroscopic = None
    micro.units = "sum"
    micro.density = None
    micro.components = [("mat_1", 0.5), ("mat_6", 0.5)]
    return m, lib, box, of


def slab_mg(num_regions=1, mat_names=None):
    """openmc.examples.slab_mg, openmc/examples.py:557-656: 1-D slab, reflective
    at x=0, vacuum at x=929.45, uniform box source over the first region."""
    if mat_names is None:
        mat_names = ["mat_" + str(i + 1) for i in range(num_regions)]
    m = Model()
    mats = []
    for i, name in enumerate(mat_names):
        mats.append(m.material(id=i + 1, name=name, macroscopic="mat_" + str(i + 1),
                               units="macro", density=1.0))
    rad_outer = 929.45
    rads = np.linspace(0., rad_outer, len(mats) + 1, endpoint=True)[1:]
    surfs = [m.surface("x-plane", [0.0], "reflective")]
    for r, rad in enumerate(rads):
        surfs.append(m.surface("x-plane", [float(rad)],
                               "vacuum" if r == len(rads) - 1 else "transmission"))
    for c in range(len(surfs) - 1):
        m.cell(region=f"{surfs[c].id} -{surfs[c + 1].id}", fill=mats[c], universe=0)
    m.settings["legendre_to_tabular"] = False      # tabular_legendre = {'enable': False}
    INF = 1000.
    box = [0., -INF, -INF, float(rads[0]), INF, INF]
    return m, box, False


# ----------------------------------------------------------------------------
# C5G7 7-group data used by examples/pincell_multigroup and mg_temperature_multi
# ----------------------------------------------------------------------------
C5G7_EDGES = [1e-5, 0.0635, 10.0, 1.0e2, 1.0e3, 0.5e6, 1.0e6, 20.0e6]
_UO2_SCATTER = [
    [0.1275370, 0.0423780, 0.0000094, 0.0000000, 0.0000000, 0.0000000, 0.0000000],
    [0.0000000, 0.3244560, 0.0016314, 0.0000000, 0.0000000, 0.0000000, 0.0000000],
    [0.0000000, 0.0000000, 0.4509400, 0.0026792, 0.0000000, 0.0000000, 0.0000000],
    [0.0000000, 0.0000000, 0.0000000, 0.4525650, 0.0055664, 0.0000000, 0.0000000],
    [0.0000000, 0.0000000, 0.0000000, 0.0001253, 0.2714010, 0.0102550, 0.0000000],
    [0.0000000, 0.0000000, 0.0000000, 0.0000000, 0.0012968, 0.2658020, 0.0168090],
    [0.0000000, 0.0000000, 0.0000000, 0.0000000, 0.0000000, 0.0085458, 0.2730800]]
_H2O_SCATTER = [
    [0.0444777, 0.1134000, 0.0007235, 0.0000037, 0.0000001, 0.0000000, 0.0000000],
    [0.0000000, 0.2823340, 0.1299400, 0.0006234, 0.0000480, 0.0000074, 0.0000010],
    [0.0000000, 0.0000000, 0.3452560, 0.2245700, 0.0169990, 0.0026443, 0.0005034],
    [0.0000000, 0.0000000, 0.0000000, 0.0910284, 0.4155100, 0.0637320, 0.0121390],
    [0.0000000, 0.0000000, 0.0000000, 0.0000714, 0.1391380, 0.5118200, 0.0612290],
    [0.0000000, 0.0000000, 0.0000000, 0.0000000, 0.0022157, 0.6999130, 0.5373200],
    [0.0000000, 0.0000000, 0.0000000, 0.0000000, 0.0000000, 0.1324400, 2.4807000]]
_UO2_TOTAL = [0.1779492, 0.3298048, 0.4803882, 0.5543674, 0.3118013, 0.3951678, 0.5644058]
_UO2_ABS = [8.0248E-03, 3.7174E-03, 2.6769E-02, 9.6236E-02, 3.0020E-02, 1.1126E-01, 2.8278E-01]
_UO2_FIS = [7.21206E-03, 8.19301E-04, 6.45320E-03, 1.85648E-02, 1.78084E-02, 8.30348E-02, 2.16004E-01]
_UO2_NUFIS = [2.005998E-02, 2.027303E-03, 1.570599E-02, 4.518301E-02, 4.334208E-02, 2.020901E-01, 5.257105E-01]
_UO2_CHI = [5.8791E-01, 4.1176E-01, 3.3906E-04, 1.1761E-07, 0.0000E+00, 0.0000E+00, 0.0000E+00]
_H2O_TOTAL = [0.15920605, 0.412969593, 0.59030986, 0.58435, 0.718, 1.2544497, 2.650379]
_H2O_ABS = [6.0105E-04, 1.5793E-05, 3.3716E-04, 1.9406E-03, 5.7416E-03, 1.5001E-02, 3.7239E-02]


def _roll(sm):
    # np.rollaxis(np.array([sm]), 0, 3): [1][G][G'] -> [G][G'][1]
    return np.rollaxis(np.array([sm]), 0, 3)


def c5g7_library(two_temperatures=False):
    """tests/regression_tests/mg_temperature_multi/test.py:10-70 (two_temperatures)
    and examples/pincell_multigroup/build_xml.py:14-69 (single temperature)."""
    lib = MGXSLibrary(C5G7_EDGES)
    temps = [294.0, 600.0] if two_temperatures else [294.0]
    uo2 = XSdata("UO2", 7, temperatures=temps)
    uo2.order = 0
    for T in temps:
        uo2.set_total(_UO2_TOTAL, temperature=T)
        uo2.set_absorption(_UO2_ABS, temperature=T)
        uo2.set_scatter_matrix(_roll(_UO2_SCATTER), temperature=T)
        uo2.set_fission(_UO2_FIS, temperature=T)
        uo2.set_nu_fission(_UO2_NUFIS, temperature=T)
        chi = list(_UO2_CHI)
        if T == 600.0:
            chi[0], chi[1] = chi[1], chi[0]           # "Altered C5G7 data (permuted Chi)"
        uo2.set_chi(chi, temperature=T)
    h2o = XSdata("LWTR", 7)
    h2o.order = 0
    h2o.set_total(_H2O_TOTAL)
    h2o.set_absorption(_H2O_ABS)
    h2o.set_scatter_matrix(_roll(_H2O_SCATTER))
    lib.add_xsdata(uo2)
    lib.add_xsdata(h2o)
    return lib


def _pincell_geometry(m: Model, fuel_mat, water_mat, pitch=1.26, r_fuel=0.54, r_inner=None,
                      inner_T=None, outer_T=None):
    """fuel pin in a reflective square box (RectangularPrism): surfaces in the
    order openmc.model.RectangularPrism creates them (min_x, max_x, min_y, max_y)."""
    ids = {}
    if r_inner is not None:
        ids["ir"] = m.surface("z-cylinder", [0.0, 0.0, r_inner]).id
    ids["or"] = m.surface("z-cylinder", [0.0, 0.0, r_fuel]).id
    ids["x0"] = m.surface("x-plane", [-pitch / 2], "reflective").id
    ids["x1"] = m.surface("x-plane", [pitch / 2], "reflective").id
    ids["y0"] = m.surface("y-plane", [-pitch / 2], "reflective").id
    ids["y1"] = m.surface("y-plane", [pitch / 2], "reflective").id
    box = f"{ids['x0']} -{ids['x1']} {ids['y0']} -{ids['y1']}"
    if r_inner is not None:
        m.cell(region=f"-{ids['ir']}", fill=fuel_mat, temperature=inner_T)
        m.cell(region=f"{ids['ir']} -{ids['or']}", fill=fuel_mat, temperature=outer_T)
    else:
        m.cell(region=f"-{ids['or']}", fill=fuel_mat)
    m.cell(region=f"{ids['or']} {box}", fill=water_mat)
    return ids


def mg_temperature_multi():
    """tests/regression_tests/mg_temperature_multi/test.py:80-160."""
    m = Model()
    uo2 = m.material(name="UO2 fuel", macroscopic="UO2", units="macro", density=1.0)
    water = m.material(name="Water", macroscopic="LWTR", units="macro", density=1.0)
    pitch = 1.26
    _pincell_geometry(m, uo2, water, pitch, 0.54, r_inner=0.25, inner_T=600.0, outer_T=294.0)
    # tallies: the second `filters =` assignment wins -> one cell filter each
    m.filters = [Filter(1, "cell", [m.cells[0].id]), Filter(2, "cell", [m.cells[1].id])]
    m.tallies = [Tally(1, [1], ["flux"]), Tally(2, [2], ["flux"])]
    box = [-pitch / 2, -pitch / 2, -1, pitch / 2, pitch / 2, 1]
    return m, c5g7_library(True), box, True


def surface_zoo():
    """Every surface kind the device path supports in one MG model (sphere, x / y / z cones,
    x / y cylinders, a general plane, axis planes) with union / complement regions -- not one
    of BASELINE.json's configurations, a geometry test case: the reference's own surface code
    is asked about the same points (tests/test_reference_geometry.py)."""
    m = Model()
    uo2 = m.material(name="UO2 fuel", macroscopic="UO2", units="macro", density=1.0)
    water = m.material(name="Water", macroscopic="LWTR", units="macro", density=1.0)
    m.surface("sphere", [0.0, 0.0, 0.0, 10.0], "vacuum")            # 1
    m.surface("x-cylinder", [0.0, 0.0, 2.0])                         # 2
    m.surface("y-cylinder", [1.0, -1.0, 1.5])                        # 3
    m.surface("plane", [1.0, 1.0, 0.5, 1.0])                         # 4
    m.surface("z-cone", [0.0, 0.0, -3.0, 0.25])                      # 5
    m.surface("x-cone", [2.0, 1.0, 0.0, 0.5])                        # 6
    m.surface("y-cone", [0.0, -2.0, 1.0, 1.0])                       # 7
    m.surface("z-plane", [4.0])                                      # 8
    m.surface("sphere", [3.0, 3.0, 3.0, 2.0])                        # 9
    m.cell(region="-9", fill=uo2)
    m.cell(region="-1 9 -2", fill=water)
    m.cell(region="-1 9 2 -3", fill=uo2)
    m.cell(region="-1 9 2 3 -5 -8", fill=water)
    m.cell(region="-1 9 2 3 (5 | 8) (-6 | -7)", fill=uo2)
    m.cell(region="-1 9 2 3 (5 | 8) 6 7 4", fill=water)
    m.cell(region="-1 9 2 3 (5 | 8) 6 7 ~(4)", fill=uo2)
    box = [-8.0, -8.0, -8.0, 8.0, 8.0, 8.0]
    return m, c5g7_library(False), box, False


def rotational_wedge():
    """a 90-degree wedge of a fuel pin in water with ROTATIONAL periodic boundaries on its two flat
    sides (RotationalPeriodicBC about z, src/boundary_condition.cpp:161-287), reflective outside"""
    m = Model()
    fuel = m.material(name="UO2 fuel", macroscopic="UO2", units="macro", density=1.0)
    water = m.material(name="Water", macroscopic="LWTR", units="macro", density=1.0)
    cyl = m.surface("z-cylinder", [0.0, 0.0, 0.45]).id
    out = m.surface("z-cylinder", [0.0, 0.0, 1.1], "reflective").id
    # the planes' normals point INTO the wedge x > 0, y > 0
    px = m.surface("plane", [1.0, 0.0, 0.0, 0.0], "periodic", id=None).id      # x = 0, normal +x
    py = m.surface("plane", [0.0, 1.0, 0.0, 0.0], "periodic").id               # y = 0, normal +y
    for sfc in m.surfaces:
        if sfc.id == px:
            sfc.periodic_surface_id = py
        elif sfc.id == py:
            sfc.periodic_surface_id = px
    z0 = m.surface("z-plane", [-1.0], "reflective").id
    z1 = m.surface("z-plane", [1.0], "reflective").id
    wedge = f"{px} {py} {z0} -{z1}"
    m.cell(region=f"-{cyl} {wedge}", fill=fuel)
    m.cell(region=f"{cyl} -{out} {wedge}", fill=water)
    return m, c5g7_library(False), [0.0, 0.0, -1.0, 0.7, 0.7, 1.0], True


def quadric_pin():
    """a fuel region bounded by general quadrics (src/surface.cpp:855-949): a tilted elliptic
    cylinder (xy cross term), cut by a paraboloid-like sheet, in a reflective box of water"""
    m = Model()
    fuel = m.material(name="UO2 fuel", macroscopic="UO2", units="macro", density=1.0)
    water = m.material(name="Water", macroscopic="LWTR", units="macro", density=1.0)
    # x^2/0.45^2 + y^2/0.3^2 + 0.8 xy - 1 = 0
    q1 = m.surface("quadric", [1 / 0.45 ** 2, 1 / 0.3 ** 2, 0.0, 0.8, 0.0, 0.0, 0.0, 0.0, 0.0, -1.0]).id
    # z = 0.6 (x^2 + y^2) + 0.1 x - 0.4   ->   0.6 x^2 + 0.6 y^2 + 0.1 x - z - 0.4 = 0
    q2 = m.surface("quadric", [0.6, 0.6, 0.0, 0.0, 0.0, 0.0, 0.1, 0.0, -1.0, -0.4]).id
    x0 = m.surface("x-plane", [-0.63], "reflective").id
    x1 = m.surface("x-plane", [0.63], "reflective").id
    y0 = m.surface("y-plane", [-0.63], "reflective").id
    y1 = m.surface("y-plane", [0.63], "reflective").id
    z0 = m.surface("z-plane", [-1.0], "reflective").id
    z1 = m.surface("z-plane", [1.0], "reflective").id
    box = f"{x0} -{x1} {y0} -{y1} {z0} -{z1}"
    m.cell(region=f"-{q1} -{q2} {z0} -{z1}", fill=fuel)
    m.cell(region=f"({q1} | {q2}) {box}", fill=water)
    return m, c5g7_library(False), [-0.63, -0.63, -1.0, 0.63, 0.63, 1.0], True


def periodic_box():
    """Fuel pin in a box with translational periodic boundaries in x (paired by id), in y
    (general planes with a non-unit normal; the two unpaired periodic surfaces pair up,
    src/surface.cpp:1279-1313); reflective in z -- test case for TranslationalPeriodicBC
    (src/boundary_condition.cpp:66-155)."""
    m = Model()
    uo2 = m.material(name="UO2 fuel", macroscopic="UO2", units="macro", density=1.0)
    water = m.material(name="Water", macroscopic="LWTR", units="macro", density=1.0)
    m.surface("z-cylinder", [0.1, 0.0, 0.45])                                        # 1
    m.surface("x-plane", [-0.7], "periodic", periodic_surface_id=3)                  # 2
    m.surface("x-plane", [0.9], "periodic", periodic_surface_id=2)                   # 3
    m.surface("plane", [0.0, 2.0, 0.0, -1.1], "periodic")                            # 4 (unpaired): y = -0.55
    m.surface("plane", [0.0, 2.0, 0.0, 1.0], "periodic")                             # 5 (unpaired): y = 0.5
    m.surface("z-plane", [-1.0], "reflective")                                       # 6
    m.surface("z-plane", [1.0], "reflective")                                        # 7
    m.cell(region="-1 6 -7", fill=uo2)
    m.cell(region="1 2 -3 4 -5 6 -7", fill=water)
    box = [-0.7, -0.55, -1.0, 0.9, 0.5, 1.0]
    return m, c5g7_library(False), box, False


def rotated_fills():
    """A 3 x 3 lattice of pin universes placed, translated and ROTATED, inside a box next to a
    rotated and translated single pin universe -- geometry test case for rotated fills
    (Cell::rotation_, src/cell.cpp:48-90; src/geometry.cpp:200-222, 324-326)."""
    m = Model()
    uo2 = m.material(name="UO2 fuel", macroscopic="UO2", units="macro", density=1.0)
    water = m.material(name="Water", macroscopic="LWTR", units="macro", density=1.0)
    cyl = m.surface("z-cylinder", [0.1, -0.05, 0.4])                 # 1 (off-centre: rotation visible)
    px = m.surface("x-plane", [0.0])                                 # 2 half fuel / half water pin
    m.cell(region="-1 -2", fill=uo2, universe=1)
    m.cell(region="-1 2", fill=water, universe=1)
    m.cell(region="1", fill=water, universe=1)
    m.cell(region="", fill=water, universe=2)                        # all space: the lattice's outer universe
    lat = m.lattice(pitch=(1.2, 1.2), lower_left=(-1.8, -1.8), universes=[[1, 1, 1], [1, 1, 1], [1, 1, 1]],
                    outer=2)
    b = [m.surface("x-plane", [-4.0], "reflective"), m.surface("x-plane", [4.0], "reflective"),
         m.surface("y-plane", [-4.0], "reflective"), m.surface("y-plane", [4.0], "reflective"),
         m.surface("z-plane", [-3.0], "reflective"), m.surface("z-plane", [3.0], "reflective")]
    s1 = m.surface("sphere", [-1.5, 0.5, 0.2, 2.0])                  # holds the lattice
    s2 = m.surface("sphere", [2.2, -1.0, -0.3, 1.2])                 # holds the single pin
    m.cell(region=f"-{s1.id}", fill=lat, universe=0, translation=(-1.5, 0.5, 0.2), rotation=(20.0, -35.0, 50.0))
    m.cell(region=f"-{s2.id}", fill=1, universe=0, translation=(2.2, -1.0, -0.3), rotation=(0.0, 0.0, 77.0))
    m.cell(region=f"{s1.id} {s2.id} {b[0].id} -{b[1].id} {b[2].id} -{b[3].id} {b[4].id} -{b[5].id}", fill=water,
           universe=0)
    box = [-4.0, -4.0, -3.0, 4.0, 4.0, 3.0]
    return m, c5g7_library(False), box, False


def pincell_multigroup(mesh_dim=100, energy_bins=500):
    """Config C1: examples/pincell_multigroup/build_xml.py:14-149 (7-group C5G7
    UO2/LWTR pin, reflective box, 100x100 mesh {flux, fission, nu-fission} and a
    500-bin log energy flux spectrum)."""
    m = Model()
    uo2 = m.material(name="UO2 fuel", macroscopic="UO2", units="macro", density=1.0)
    water = m.material(name="Water", macroscopic="LWTR", units="macro", density=1.0)
    pitch = 1.26
    _pincell_geometry(m, uo2, water, pitch, 0.54)
    m.meshes = [Mesh(1, [mesh_dim, mesh_dim], [-pitch / 2, -pitch / 2], [pitch / 2, pitch / 2])]
    edges = np.logspace(np.log10(1e-5), np.log10(20.0e6), energy_bins + 1)
    m.filters = [Filter(1, "mesh", 1), Filter(2, "energy", list(edges))]
    m.tallies = [Tally(1, [1], ["flux", "fission", "nu-fission"]), Tally(2, [2], ["flux"])]
    box = [-pitch / 2, -pitch / 2, -1, pitch / 2, pitch / 2, 1]
    return m, c5g7_library(False), box, True


# ----------------------------------------------------------------------------
# Config C2: UO2 pincell, continuous energy, synthetic library
# ----------------------------------------------------------------------------
def pincell_ce(n_grid=100000, mesh_dim=100, energy_bins=500, tallies=True, library=None, temperatures=None):
    """examples/pincell/build_xml.py:11-120 geometry/settings (fuel, gap, clad,
    borated water in a reflective 1.25984 cm box; 100x100 mesh {flux, fission,
    nu-fission} + 500-bin flux spectrum) with the 10-nuclide synthetic library of
    openmc_b200.nuclide (no S(a,b)).  `temperatures` = (library temperatures [K], cell temperatures
    [K] of fuel, gap, clad, water): a library with several tables per nuclide and cells that pick
    among them (nearest temperature, the reference's default method)."""
    from . import nuclide as nd
    m = Model()
    uo2 = m.material(name="UO2 fuel at 2.4% wt enrichment", units="g/cm3", density=10.29769,
                     components=[("U235", 0.024 * 238.0 / 235.0 / (1 + 0.024 * (238.0 / 235.0 - 1))),
                                 ("U238", 1.0 - 0.024 * 238.0 / 235.0 / (1 + 0.024 * (238.0 / 235.0 - 1))),
                                 ("O16", 2.0)])
    helium = m.material(name="Helium for gap", units="g/cm3", density=0.001598,
                        components=[("He4", 1.0)])
    zirc = m.material(name="Zircaloy 4", units="g/cm3", density=6.55,
                      components=[("Sn120", -0.014), ("Fe56", -0.00265), ("Zr90", -0.60),
                                  ("Zr91", -0.38335)])
    water = m.material(name="Borated water", units="g/cm3", density=0.740582,
                       components=[("B10", 8.0e-6), ("H1", 5.0e-2), ("O16", 2.4e-2)])
    pitch = 1.25984
    s_fo = m.surface("z-cylinder", [0.0, 0.0, 0.39218]).id
    s_ci = m.surface("z-cylinder", [0.0, 0.0, 0.40005]).id
    s_co = m.surface("z-cylinder", [0.0, 0.0, 0.45720]).id
    x0 = m.surface("x-plane", [-pitch / 2], "reflective").id
    x1 = m.surface("x-plane", [pitch / 2], "reflective").id
    y0 = m.surface("y-plane", [-pitch / 2], "reflective").id
    y1 = m.surface("y-plane", [pitch / 2], "reflective").id
    cT = temperatures[1] if temperatures else (None, None, None, None)
    m.cell(region=f"-{s_fo}", fill=uo2, temperature=cT[0])
    m.cell(region=f"{s_fo} -{s_ci}", fill=helium, temperature=cT[1])
    m.cell(region=f"{s_ci} -{s_co}", fill=zirc, temperature=cT[2])
    m.cell(region=f"{s_co} {x0} -{x1} {y0} -{y1}", fill=water, temperature=cT[3])
    if tallies:
        m.meshes = [Mesh(1, [mesh_dim, mesh_dim], [-pitch / 2, -pitch / 2], [pitch / 2, pitch / 2])]
        edges = np.logspace(np.log10(1e-5), np.log10(20.0e6), energy_bins + 1)
        m.filters = [Filter(1, "mesh", 1), Filter(2, "energy", list(edges))]
        m.tallies = [Tally(1, [1], ["flux", "fission", "nu-fission"]), Tally(2, [2], ["flux"])]
    if library is None:
        library = nd.synthetic_library(nd.PINCELL_NUCLIDES, n_grid,
                                       temperatures=temperatures[0] if temperatures else None)
    box = [-pitch / 2, -pitch / 2, -1, pitch / 2, pitch / 2, 1]
    return m, library, box, True


# ----------------------------------------------------------------------------
# config C3: 17 x 17 PWR assembly, continuous energy (examples/assembly/assembly.py:20-122)
# ----------------------------------------------------------------------------
ASSEMBLY_NUCLIDES = [
    ("U234", 92, 234, True, False, False), ("U235", 92, 235, True, True, True),
    ("U238", 92, 238, True, True, False), ("O16", 8, 16, False, False, False),
    ("Zr90", 40, 90, False, True, False), ("Zr91", 40, 91, False, False, False),
    ("Zr92", 40, 92, False, False, False), ("Zr94", 40, 94, False, False, False),
    ("Zr96", 40, 96, False, False, False), ("H1", 1, 1, False, False, False),
    ("B10", 5, 10, False, False, False), ("B11", 5, 11, False, False, False),
]
# guide-tube positions (row, column) of examples/assembly/assembly.py:86-94
GUIDE_TUBES = [(2, 5), (2, 8), (2, 11), (3, 3), (3, 13), (5, 2), (5, 5), (5, 8), (5, 11), (5, 14),
               (8, 2), (8, 5), (8, 8), (8, 11), (8, 14), (11, 2), (11, 5), (11, 8), (11, 11), (11, 14),
               (13, 3), (13, 13), (14, 5), (14, 8), (14, 11)]


def _pin_universes(m: Model, fuel, clad, water, u_fuel, u_gt, r_fo=0.39218, r_co=0.45720):
    """fuel-pin and guide-tube universes (assembly.py:49-73); cells of a universe
    are listed inside-out like the reference builds them"""
    fo = m.surface("z-cylinder", [0.0, 0.0, r_fo]).id
    co = m.surface("z-cylinder", [0.0, 0.0, r_co]).id
    m.cell(region=f"-{fo}", fill=fuel, universe=u_fuel)
    m.cell(region=f"{fo} -{co}", fill=clad, universe=u_fuel)
    m.cell(region=f"{co}", fill=water, universe=u_fuel)
    m.cell(region=f"-{fo}", fill=water, universe=u_gt)
    m.cell(region=f"{fo} -{co}", fill=clad, universe=u_gt)
    m.cell(region=f"{co}", fill=water, universe=u_gt)


def assembly_ce(n_grid=100000, mesh_dim=17, tallies=True, library=None):
    """examples/assembly/assembly.py geometry (17 x 17 lattice, 25 guide tubes,
    reflective box, pitch 21.42 cm) with synthetic nuclides of the same names,
    no S(a,b); pin-wise mesh {flux, fission} tally (BASELINE.json config C3)."""
    from . import nuclide as nd
    m = Model()
    fuel = m.material(name="Fuel", units="g/cm3", density=10.29769,
                      components=[("U234", 4.4843e-6), ("U235", 5.5815e-4), ("U238", 2.2408e-2),
                                  ("O16", 4.5829e-2)])
    clad = m.material(name="Cladding", units="g/cm3", density=6.55,
                      components=[("Zr90", 2.1827e-2), ("Zr91", 4.7600e-3), ("Zr92", 7.2758e-3),
                                  ("Zr94", 7.3734e-3), ("Zr96", 1.1879e-3)])
    water = m.material(name="Hot borated water", units="g/cm3", density=0.740582,
                       components=[("H1", 4.9457e-2), ("O16", 2.4672e-2), ("B10", 8.0042e-6),
                                   ("B11", 3.2218e-5)])
    _pin_universes(m, fuel, clad, water, u_fuel=1, u_gt=2)
    pitch = 21.42
    univ = np.full((17, 17), 1)
    for r, c in GUIDE_TUBES:
        univ[r, c] = 2
    lat = m.lattice(pitch=(pitch / 17, pitch / 17), lower_left=(-pitch / 2, -pitch / 2),
                    universes=univ.tolist())
    x0 = m.surface("x-plane", [-pitch / 2], "reflective").id
    x1 = m.surface("x-plane", [pitch / 2], "reflective").id
    y0 = m.surface("y-plane", [-pitch / 2], "reflective").id
    y1 = m.surface("y-plane", [pitch / 2], "reflective").id
    m.cell(region=f"{x0} -{x1} {y0} -{y1}", fill=lat, universe=0)
    if tallies:
        m.meshes = [Mesh(1, [mesh_dim, mesh_dim], [-pitch / 2, -pitch / 2], [pitch / 2, pitch / 2])]
        m.filters = [Filter(1, "mesh", 1)]
        m.tallies = [Tally(1, [1], ["flux", "fission"])]
    if library is None:
        library = nd.synthetic_library(ASSEMBLY_NUCLIDES, n_grid)
    box = [-pitch / 2, -pitch / 2, -1, pitch / 2, pitch / 2, 1]
    return m, library, box, True


# ----------------------------------------------------------------------------
# config C4: C5G7-style 2-D core, multigroup, lattices of lattices
# ----------------------------------------------------------------------------
def c5g7_core_library():
    """UO2 / water: the literal C5G7 data above.  MOX (three enrichments), guide
    tube and fission chamber sets are GENERATED from them by fixed scalings (the
    benchmark's own tables are not in the reference tree): MOX raises thermal
    absorption / fission and hardens chi slightly; the guide tube is water with
    10 % less scattering; the fission chamber is a guide tube with a trace of
    fission.  Synthetic, stated as such (BASELINE.json: 'generated by script')."""
    lib = c5g7_library()
    uo2 = lib.xsdatas["UO2"].data[294.0]
    h2o = lib.xsdatas["LWTR"].data[294.0]

    def fuel_like(name, k_abs, k_fis):
        x = XSdata(name, 7)
        x.order = 0
        sc = np.array(uo2["scatter_matrix"])
        ab = np.array(_UO2_ABS) * np.array([1.0, 1.0, 1.1, 1.2, k_abs, k_abs, k_abs])
        fi = np.array(_UO2_FIS) * np.array([1.0, 1.0, 1.05, 1.1, k_fis, k_fis, k_fis])
        nu = np.array(_UO2_NUFIS) / np.array(_UO2_FIS) * 1.17        # nu of Pu-bearing fuel
        x.set_absorption(ab)
        x.set_scatter_matrix(sc)
        x.set_total(ab + sc.sum(axis=(1, 2)))
        x.set_fission(fi)
        x.set_nu_fission(fi * nu)
        x.set_chi(_UO2_CHI)
        return x

    for name, ka, kf in (("MOX43", 2.0, 1.6), ("MOX70", 2.7, 2.2), ("MOX87", 3.1, 2.6)):
        lib.add_xsdata(fuel_like(name, ka, kf))
    gt = XSdata("GT", 7)
    gt.order = 0
    sc = np.array(h2o["scatter_matrix"]) * 0.9
    gt.set_absorption(np.array(_H2O_ABS) * 0.9)
    gt.set_scatter_matrix(sc)
    gt.set_total(np.array(_H2O_ABS) * 0.9 + sc.sum(axis=(1, 2)))
    lib.add_xsdata(gt)
    fc = XSdata("FC", 7)
    fc.order = 0
    fis = np.array(_UO2_FIS) * 1.0e-6
    fc.set_absorption(np.array(_H2O_ABS) * 0.9 + fis)
    fc.set_scatter_matrix(sc)
    fc.set_total(np.array(_H2O_ABS) * 0.9 + fis + sc.sum(axis=(1, 2)))
    fc.set_fission(fis)
    fc.set_nu_fission(fis * 2.5)
    fc.set_chi(_UO2_CHI)
    lib.add_xsdata(fc)
    return lib


def c5g7_core(mesh_dim=510, tallies=True):
    """Quarter core: 2 x 2 fuel assemblies (UO2 / MOX checkerboard) + reflector,
    each assembly a 17 x 17 pin lattice inside a 3 x 3 core lattice (3 coordinate
    levels); reflective on the west / north faces, vacuum on the east / south ones;
    fine mesh {flux, fission} tally (BASELINE.json config C4)."""
    m = Model()
    names = ["UO2", "MOX43", "MOX70", "MOX87", "GT", "FC", "LWTR"]
    mats = {n: m.material(name=n, macroscopic=n, units="macro", density=1.0) for n in names}
    r_pin = 0.54
    cyl = m.surface("z-cylinder", [0.0, 0.0, r_pin]).id
    pin_u = {}
    for k, n in enumerate(["UO2", "MOX43", "MOX70", "MOX87", "GT", "FC"]):
        u = 10 + k
        pin_u[n] = u
        m.cell(region=f"-{cyl}", fill=mats[n], universe=u)
        m.cell(region=f"{cyl}", fill=mats["LWTR"], universe=u)
    u_water = 20
    m.cell(region="", fill=mats["LWTR"], universe=u_water)     # infinite water cell
    pitch = 1.26
    w = 17 * pitch

    def assembly(kind):
        a = np.full((17, 17), pin_u["UO2"])
        if kind == "MOX":      # enrichment zoning: 4.3 % rim, 7.0 % next ring, 8.7 % interior
            a[:, :] = pin_u["MOX87"]
            a[1:-1, 1:-1][[0, -1], :] = pin_u["MOX70"]
            a[1:-1, 1:-1][:, [0, -1]] = pin_u["MOX70"]
            a[[0, -1], :] = pin_u["MOX43"]
            a[:, [0, -1]] = pin_u["MOX43"]
        for r, c in GUIDE_TUBES:
            a[r, c] = pin_u["GT"]
        a[8, 8] = pin_u["FC"]
        return a.tolist()

    u_asm = {}
    for k, kind in enumerate(("UO2", "MOX")):
        lat = m.lattice(pitch=(pitch, pitch), lower_left=(-w / 2, -w / 2), universes=assembly(kind))
        u = 30 + k
        u_asm[kind] = u
        m.cell(region="", fill=lat, universe=u)
    core = m.lattice(pitch=(w, w), lower_left=(0.0, -3 * w),
                     universes=[[u_asm["UO2"], u_asm["MOX"], u_water],
                                [u_asm["MOX"], u_asm["UO2"], u_water],
                                [u_water, u_water, u_water]])
    x0 = m.surface("x-plane", [0.0], "reflective").id
    x1 = m.surface("x-plane", [3 * w], "vacuum").id
    y0 = m.surface("y-plane", [-3 * w], "vacuum").id
    y1 = m.surface("y-plane", [0.0], "reflective").id
    m.cell(region=f"{x0} -{x1} {y0} -{y1}", fill=core, universe=0)
    m.settings["legendre_to_tabular"] = True
    if tallies:
        m.meshes = [Mesh(1, [mesh_dim, mesh_dim], [0.0, -3 * w], [3 * w, 0.0])]
        m.filters = [Filter(1, "mesh", 1)]
        m.tallies = [Tally(1, [1], ["flux", "fission"])]
    box = [0.0, -2 * w, -1, 2 * w, 0.0, 1]
    return m, c5g7_core_library(), box, True


# ----------------------------------------------------------------------------
# hexagonal lattices (src/lattice.cpp:458-1009): a VVER-like bundle, multigroup
# ----------------------------------------------------------------------------
def hex_bundle(orientation="y", n_axial=None, n_rings=3, tallies=True):
    """n_rings rings of pins on a hexagonal pitch (centre pin MOX, then UO2 with guide tubes at
    three corners of the outer ring) in a vacuum cylinder; positions of the lattice outside the
    cylinder's rings fall to the `outer` water universe.  n_axial: a three-dimensional lattice
    (the axial levels alternate the centre pin)."""
    from .model import HexLattice  # noqa: F401
    m = Model()
    names = ["UO2", "MOX70", "GT", "LWTR"]
    mats = {n: m.material(name=n, macroscopic=n, units="macro", density=1.0) for n in names}
    cyl = m.surface("z-cylinder", [0.0, 0.0, 0.45]).id
    pin_u = {}
    for k, n in enumerate(["UO2", "MOX70", "GT"]):
        pin_u[n] = 10 + k
        m.cell(region=f"-{cyl}", fill=mats[n], universe=10 + k)
        m.cell(region=f"{cyl}", fill=mats["LWTR"], universe=10 + k)
    u_water = 20
    m.cell(region="", fill=mats["LWTR"], universe=u_water)
    pitch = 1.275
    n = 2 * n_rings - 1
    nz = n_axial or 1
    arr = np.full((nz, n, n), -1, dtype=np.int64)
    for z in range(nz):
        for i1 in range(n):
            for i0 in range(n):
                a, b = i0 - (n_rings - 1), i1 - (n_rings - 1)
                if not (n_rings - 2 < i0 + i1 < 3 * n_rings - 2):
                    continue
                ring = max(abs(a), abs(b), abs(a + b))
                u = pin_u["UO2"]
                if ring == 0:
                    u = pin_u["MOX70"] if z % 2 == 0 else pin_u["GT"]
                elif ring == n_rings - 1 and (a, b) in ((n_rings - 1, 0), (0, -(n_rings - 1)), (-(n_rings - 1), n_rings - 1)):
                    u = pin_u["GT"]
                arr[z, i1, i0] = u
    height = 4.0
    lat = m.hex_lattice(n_rings=n_rings, pitch=(pitch,) if n_axial is None else (pitch, height / nz),
                        center=(0.1, -0.05) if n_axial is None else (0.1, -0.05, 0.0),
                        universes=arr, orientation=orientation, n_axial=n_axial, outer=u_water)
    r_out = pitch * (n_rings - 0.3)
    rc = m.surface("z-cylinder", [0.1, -0.05, r_out], "vacuum").id
    z0 = m.surface("z-plane", [-height / 2], "reflective").id
    z1 = m.surface("z-plane", [height / 2], "reflective").id
    m.cell(region=f"-{rc} {z0} -{z1}", fill=lat, universe=0)
    m.settings["legendre_to_tabular"] = True
    if tallies:
        half = r_out
        m.meshes = [Mesh(1, [8, 8], [0.1 - half, -0.05 - half], [0.1 + half, -0.05 + half])]
        fuel_cell = m.cells[0].id     # the fuel of the UO2 pin universe: one instance per UO2 position
        m.filters = [Filter(1, "mesh", 1), Filter(2, "distribcell", [fuel_cell])]
        m.tallies = [Tally(1, [1], ["flux", "fission"]), Tally(2, [2], ["flux", "nu-fission"])]
    s = 0.6 * r_out
    box = [0.1 - s, -0.05 - s, -height / 2, 0.1 + s, -0.05 + s, height / 2]
    return m, c5g7_core_library(), box, True


# ----------------------------------------------------------------------------
# config C5: Hoogenboom-Martin full core (openmc/examples.py:94-439), continuous
# energy, fuel expanded to a depleted-fuel-like set of synthetic nuclides
# ----------------------------------------------------------------------------
_STRUCT_ZA = {"U234": (92, 234), "U235": (92, 235), "U238": (92, 238), "Xe135": (54, 135), "O16": (8, 16),
              "Zr90": (40, 90), "Zr91": (40, 91), "Zr92": (40, 92), "Zr94": (40, 94), "Zr96": (40, 96),
              "H1": (1, 1), "B10": (5, 10), "B11": (5, 11), "Fe54": (26, 54), "Fe56": (26, 56),
              "Fe57": (26, 57), "Fe58": (26, 58), "Ni58": (28, 58), "Ni60": (28, 60), "Mn55": (25, 55),
              "Cr52": (24, 52), "C0": (6, 12), "Cu63": (29, 63)}


def pwr_core_nuclides(n_fuel_nuclides=300):
    """(name, Z, A, fissionable, inelastic, delayed) for the structural nuclides of
    pwr_core plus synthetic fission products / actinides FPnnn filling the fuel up
    to n_fuel_nuclides"""
    out = []
    for name, (Z, A) in _STRUCT_ZA.items():
        fis = name in ("U234", "U235", "U238")
        out.append((name, Z, A, fis, name in ("U235", "U238", "Fe56", "Zr90"), name == "U235"))
    g = np.random.default_rng(5)
    for k in range(max(n_fuel_nuclides - 5, 0)):
        Z = int(g.integers(30, 96))
        A = int(round(Z * 2.45 + g.integers(-3, 4)))
        out.append((f"FP{k:03d}", Z, A, bool(Z >= 90 and g.uniform() < 0.3), False, False))
    return out


def pwr_core_ce(n_grid=100000, n_fuel_nuclides=300, tallies=False, library=None):
    """OECD/NEA Monte Carlo performance benchmark geometry, verbatim from
    openmc/examples.py:255-425 (241 assemblies of 17 x 17 pins in a 21 x 21 core
    lattice, lower / upper halves, axial reflector regions, vacuum outside), no
    S(a,b).  The fuel's five nuclides are joined by synthetic ones with log-uniform
    random densities (seed 5) -- the xs-lookup stress case (BASELINE.json C5)."""
    from . import nuclide as nd
    m = Model()
    nucs = pwr_core_nuclides(n_fuel_nuclides)
    g = np.random.default_rng(5)
    extra = [(n[0], float(np.exp(g.uniform(np.log(1e-9), np.log(1e-5))))) for n in nucs[len(_STRUCT_ZA):]]
    fuel = m.material(id=1, name="UOX fuel", units="g/cm3", density=10.062,
                      components=[("U234", 4.9476e-6), ("U235", 4.8218e-4), ("U238", 2.1504e-2),
                                  ("Xe135", 1.0801e-8), ("O16", 4.5737e-2)] + extra)
    clad = m.material(id=2, name="Zircaloy", units="g/cm3", density=5.77,
                      components=[("Zr90", 0.5145), ("Zr91", 0.1122), ("Zr92", 0.1715),
                                  ("Zr94", 0.1738), ("Zr96", 0.0280)])
    water_c = [("H1", 2.0), ("O16", 1.0), ("B10", 6.490e-4), ("B11", 2.689e-3)]
    cold = m.material(id=3, name="Cold borated water", units="atom/b-cm", density=0.07416, components=water_c)
    hot = m.material(id=4, name="Hot borated water", units="atom/b-cm", density=0.06614, components=water_c)
    rpv = m.material(id=5, name="Reactor pressure vessel steel", units="g/cm3", density=7.9, components=[
        ("Fe54", -0.05437098), ("Fe56", -0.88500663), ("Fe57", -0.0208008), ("Fe58", -0.00282159),
        ("Ni58", -0.0067198), ("Ni60", -0.0026776), ("Mn55", -0.01), ("Cr52", -0.002092475),
        ("C0", -0.0025), ("Cu63", -0.0013696)])

    def steel_water(id, name, density, h1, o16, b10, b11, fe54, fe56, fe57, fe58, ni58, mn55, cr52):
        return m.material(id=id, name=name, units="g/cm3", density=density, components=[
            ("H1", -h1), ("O16", -o16), ("B10", -b10), ("B11", -b11), ("Fe54", -fe54), ("Fe56", -fe56),
            ("Fe57", -fe57), ("Fe58", -fe58), ("Ni58", -ni58), ("Mn55", -mn55), ("Cr52", -cr52)])
    lower_rad = steel_water(6, "Lower radial reflector", 4.32, 0.0095661, 0.0759107, 3.08409e-5, 1.40499e-4,
                            0.035620772088, 0.579805982228, 0.01362750048, 0.001848545204,
                            0.055298376566, 0.0182870, 0.145407678031)
    upper_rad = steel_water(7, "Upper radial reflector / Top plate region", 4.28, 0.0086117, 0.0683369,
                            2.77638e-5, 1.26481e-4, 0.035953677186, 0.585224740891, 0.01375486056,
                            0.001865821363, 0.055815129186, 0.0184579, 0.146766614995)
    bot_plate = steel_water(8, "Bottom plate region", 7.184, 0.0011505, 0.0091296, 3.70915e-6, 1.68974e-5,
                            0.03855611055, 0.627585036425, 0.014750478, 0.002000875025, 0.059855207342,
                            0.0197940, 0.157390026871)
    bot_nozzle = steel_water(9, "Bottom nozzle region", 2.53, 0.0245014, 0.1944274, 7.89917e-5, 3.59854e-4,
                             0.030411411144, 0.495012237964, 0.01163454624, 0.001578204652,
                             0.047211231662, 0.0156126, 0.124142524198)
    top_nozzle = steel_water(10, "Top nozzle region", 1.746, 0.0358870, 0.2847761, 1.15699e-4, 5.27075e-4,
                             0.02644016154, 0.43037146399, 0.0101152584, 0.00137211607, 0.04104621835,
                             0.0135739, 0.107931450781)

    def zr_water(id, name, density, h1, o16, b10, b11, z90, z91, z92, z94, z96):
        return m.material(id=id, name=name, units="g/cm3", density=density, components=[
            ("H1", -h1), ("O16", -o16), ("B10", -b10), ("B11", -b11), ("Zr90", -z90), ("Zr91", -z91),
            ("Zr92", -z92), ("Zr94", -z94), ("Zr96", -z96)])
    top_fa = zr_water(11, "Top of fuel assemblies", 3.044, 0.0162913, 0.1292776, 5.25228e-5, 2.39272e-4,
                      0.43313403903, 0.09549277374, 0.14759527104, 0.15280552077, 0.02511169542)
    bot_fa = zr_water(12, "Bottom of fuel assemblies", 1.762, 0.0292856, 0.2323919, 9.44159e-5, 4.30120e-4,
                      0.3741373658, 0.0824858164, 0.1274914944, 0.1319920622, 0.0216912612)

    for sid, r in ((1, 0.41), (2, 0.475), (3, 0.56), (4, 0.62), (5, 187.6), (6, 209.0), (7, 229.0)):
        m.surface("z-cylinder", [0.0, 0.0, r], id=sid)
    m.surface("z-cylinder", [0.0, 0.0, 249.0], "vacuum", id=8)
    for sid, z, bc in ((31, -229.0, "vacuum"), (32, -199.0, "transmission"), (33, -193.0, "transmission"),
                       (34, -183.0, "transmission"), (35, 0.0, "transmission"), (36, 183.0, "transmission"),
                       (37, 203.0, "transmission"), (38, 215.0, "transmission"), (39, 223.0, "vacuum")):
        m.surface("z-plane", [z], bc, id=sid)

    def pin(u, base, inner, outer, s_in, s_out):
        m.cell(id=base, region=f"-{s_in}", fill=inner, universe=u)
        m.cell(id=base + 1, region=f"{s_in} -{s_out}", fill=clad, universe=u)
        m.cell(id=base + 2, region=f"{s_out}", fill=outer, universe=u)
    pin(1, 21, fuel, cold, 1, 2)
    pin(2, 24, cold, cold, 3, 4)
    pin(3, 27, fuel, hot, 1, 2)
    pin(4, 30, hot, hot, 3, 4)
    tube_x = [5, 8, 11, 3, 13, 2, 5, 8, 11, 14, 2, 5, 8, 11, 14, 2, 5, 8, 11, 14, 3, 13, 5, 8, 11]
    tube_y = [2, 2, 2, 3, 3, 5, 5, 5, 5, 5, 8, 8, 8, 8, 8, 11, 11, 11, 11, 11, 13, 13, 14, 14, 14]

    def assembly(u_fuel, u_tube, lid):
        a = np.full((17, 17), u_fuel)
        a[tube_x, tube_y] = u_tube
        return m.lattice(id=lid, pitch=(1.26, 1.26), lower_left=(-10.71, -10.71), universes=a.tolist())
    l100 = assembly(1, 2, 100)
    l101 = assembly(3, 4, 101)
    m.cell(id=50, region="34 -35", fill=cold, universe=5)
    m.cell(id=70, region="35 -36", fill=hot, universe=7)
    m.cell(id=60, region="34 -35", fill=l100, universe=6)
    m.cell(id=80, region="35 -36", fill=l101, universe=8)
    widths = [0, 0, 7, 11, 13, 15, 15, 17, 17, 17, 17, 17, 17, 17, 15, 15, 13, 11, 7, 0, 0]

    def core(u_water, u_fa, lid):
        rows = [[u_water] * ((21 - w) // 2) + [u_fa] * w + [u_water] * (21 - w - (21 - w) // 2) for w in widths]
        return m.lattice(id=lid, pitch=(21.42, 21.42), lower_left=(-224.91, -224.91), universes=rows)
    l200 = core(5, 6, 200)
    l201 = core(7, 8, 201)
    m.cell(id=1, region="-6 34 -35", fill=l200, universe=0)
    m.cell(id=2, region="-6 35 -36", fill=l201, universe=0)
    m.cell(id=3, region="-7 31 -32", fill=bot_plate, universe=0)
    m.cell(id=4, region="-5 32 -33", fill=bot_nozzle, universe=0)
    m.cell(id=5, region="-5 33 -34", fill=bot_fa, universe=0)
    m.cell(id=6, region="-5 36 -37", fill=top_fa, universe=0)
    m.cell(id=7, region="-5 37 -38", fill=top_nozzle, universe=0)
    m.cell(id=8, region="-7 38 -39", fill=upper_rad, universe=0)
    m.cell(id=9, region="6 -7 32 -38", fill=bot_nozzle, universe=0)
    m.cell(id=10, region="7 -8 31 -39", fill=rpv, universe=0)
    m.cell(id=11, region="5 -6 32 -34", fill=lower_rad, universe=0)
    m.cell(id=12, region="5 -6 36 -38", fill=upper_rad, universe=0)
    if tallies:
        m.meshes = [Mesh(1, [21 * 17, 21 * 17], [-224.91, -224.91], [224.91, 224.91])]
        m.filters = [Filter(1, "mesh", 1)]
        m.tallies = [Tally(1, [1], ["flux", "fission"])]
    if library is None:
        library = nd.synthetic_library(nucs, n_grid)
    box = [-160.0, -160.0, -183.0, 160.0, 160.0, 183.0]
    return m, library, box, False
