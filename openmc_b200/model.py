"""Host-side model description and flattener.

In the real integration libopenmc's own loaders (``read_model_xml``,
``finalize_geometry``, ``finalize_cross_sections``) build ``model::`` / ``data::``
and a flattener copies them into the POD arrays of ``omcgpu_model``
(INTEGRATION.md).  The XML/HDF5 front end is out of scope here, so
this module provides the small amount of host logic that produces the *same
flattened arrays* from a Python description: surfaces, CSG regions, cells,
universes, rectangular lattices, materials and tallies.  The derived quantities
follow the reference's load-time code:

* region tokenisation / De Morgan / simple-cell stripping   src/cell.cpp:556-682
* lattice universe ordering (rows listed top to bottom)     src/lattice.cpp:214-231
* cell sqrt(kT) defaults                                    src/geometry_aux.cpp (assign_temperatures)
* density normalisation                                     src/material.cpp:417-467
* tally strides                                             src/tallies/tally.cpp:503-516
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _abi as abi

K_BOLTZMANN = 8.617333262e-5       # include/openmc/constants.h:102
N_AVOGADRO = 0.602214076
MASS_NEUTRON = 1.00866491595
TEMPERATURE_DEFAULT = 293.6         # src/settings.cpp:141

_SURF_KINDS = {
    "x-plane": (abi.OMCGPU_SURF_XPLANE, 1), "y-plane": (abi.OMCGPU_SURF_YPLANE, 1),
    "z-plane": (abi.OMCGPU_SURF_ZPLANE, 1), "plane": (abi.OMCGPU_SURF_PLANE, 4),
    "x-cylinder": (abi.OMCGPU_SURF_XCYLINDER, 3), "y-cylinder": (abi.OMCGPU_SURF_YCYLINDER, 3),
    "z-cylinder": (abi.OMCGPU_SURF_ZCYLINDER, 3), "sphere": (abi.OMCGPU_SURF_SPHERE, 4),
    "x-cone": (abi.OMCGPU_SURF_XCONE, 4), "y-cone": (abi.OMCGPU_SURF_YCONE, 4),
    "z-cone": (abi.OMCGPU_SURF_ZCONE, 4), "quadric": (abi.OMCGPU_SURF_QUADRIC, 10),   # 10 coefficients in quad_coef
}
_BCS = {"transmission": abi.OMCGPU_BC_NONE, "vacuum": abi.OMCGPU_BC_VACUUM,
        "reflective": abi.OMCGPU_BC_REFLECTIVE, "periodic": abi.OMCGPU_BC_PERIODIC,
        "white": abi.OMCGPU_BC_WHITE}
_SCORES = {"flux": abi.OMCGPU_SCORE_FLUX, "total": abi.OMCGPU_SCORE_TOTAL,
           "absorption": abi.OMCGPU_SCORE_ABSORPTION, "fission": abi.OMCGPU_SCORE_FISSION,
           "nu-fission": abi.OMCGPU_SCORE_NU_FISSION, "scatter": abi.OMCGPU_SCORE_SCATTER}

OP_LP, OP_RP = abi.OMCGPU_OP_LEFT_PAREN, abi.OMCGPU_OP_RIGHT_PAREN
OP_COMP, OP_INT, OP_UNION = (abi.OMCGPU_OP_COMPLEMENT, abi.OMCGPU_OP_INTERSECTION,
                             abi.OMCGPU_OP_UNION)


@dataclass
class Surface:
    id: int
    kind: str
    coeffs: Sequence[float]
    boundary: str = "transmission"
    periodic_surface_id: Optional[int] = None   # partner of a 'periodic' plane


@dataclass
class Material:
    id: int
    name: str = ""
    # [(nuclide or macroscopic name, atom fraction (>0) or weight fraction (<0))]
    components: list = field(default_factory=list)
    density: Optional[float] = None
    units: str = "atom/b-cm"           # 'macro', 'sum', 'g/cm3', 'atom/b-cm'
    macroscopic: Optional[str] = None  # <macroscopic name=.../>
    temperature: Optional[float] = None


@dataclass
class Cell:
    id: int
    region: str = ""                   # OpenMC region syntax with surface ids
    fill: object = None                # Material | Universe id | RectLattice | None (void)
    universe: int = 0
    temperature: Optional[float] = None
    translation: Sequence[float] = (0.0, 0.0, 0.0)
    rotation: Optional[Sequence[float]] = None   # (phi, theta, psi) in degrees, fill cells only
    density: object = None             # g/cm3 of the cell (one value or one per instance): Cell::density_mult_
    fill_type: str = "material"        # 'material' | 'universe' | 'lattice'


@dataclass
class RectLattice:
    id: int
    pitch: Sequence[float]
    lower_left: Sequence[float]
    universes: object                  # [ny][nx] or [nz][ny][nx], rows listed top (max y) first
    outer: Optional[int] = None


@dataclass
class HexLattice:
    """HexLattice of src/lattice.cpp:458-1009.  `universes` is the reference's own storage: the
    skewed "square" array [n_axial][2 n_rings - 1][2 n_rings - 1] indexed [z][i_alpha][i_x]
    (orientation y; [z][i_y][i_alpha] for x), flat index (2n-1)^2 z + (2n-1) i1 + i0
    (get_flat_index :975-981), with -1 at the positions outside the hexagon
    (are_valid_indices :727-735)."""
    id: int
    n_rings: int
    pitch: Sequence[float]             # (radial,) or (radial, axial)
    center: Sequence[float]            # (x, y) or (x, y, z)
    universes: object
    orientation: str = "y"
    n_axial: Optional[int] = None      # None: two-dimensional
    outer: Optional[int] = None

    def valid(self, i0, i1):
        n = self.n_rings
        return 0 <= i0 < 2 * n - 1 and 0 <= i1 < 2 * n - 1 and n - 2 < i0 + i1 < 3 * n - 2

    def input_order(self):
        """(i1, i0) array positions in the order the <universes> text lists one axial level:
        the index walks of fill_lattice_y / fill_lattice_x (src/lattice.cpp:548-685)"""
        n = self.n_rings
        seq = []
        if self.orientation == "y":
            i_x, i_a = 1, n - 1
            for k in range(n - 1):                      # upper triangle
                i_x -= 1
                for j in range(k + 1):
                    seq.append((i_a + n - 1, i_x + n - 1)); i_x += 2; i_a -= 1
                i_x -= 2 * (k + 1); i_a += k + 1
            for k in range(2 * n - 1):                  # middle
                if k % 2 == 0:
                    i_x -= 1
                else:
                    i_x += 1; i_a -= 1
                for j in range(n - (k % 2)):
                    seq.append((i_a + n - 1, i_x + n - 1)); i_x += 2; i_a -= 1
                i_x -= 2 * (n - (k % 2)); i_a += n - (k % 2)
            for k in range(n - 1):                      # lower triangle
                i_x += 1; i_a -= 1
                for j in range(n - k - 1):
                    seq.append((i_a + n - 1, i_x + n - 1)); i_x += 2; i_a -= 1
                i_x -= 2 * (n - k - 1); i_a += n - k - 1
        else:
            i_a, i_y = -(n - 1), n - 1
            for k in range(n - 1):
                for j in range(k + n):
                    seq.append((i_y + n - 1, i_a + n - 1)); i_a += 1
                i_a = -(n - 1); i_y -= 1
            for k in range(n):
                i_a = -(n - 1) + k
                for j in range(2 * n - k - 1):
                    seq.append((i_y + n - 1, i_a + n - 1)); i_a += 1
                i_y -= 1
        return seq


@dataclass
class Mesh:
    id: int
    dimension: Sequence[int]
    lower_left: Sequence[float]
    upper_right: Sequence[float]


@dataclass
class Filter:
    id: int
    type: str                          # 'mesh' 'energy' 'energyout' 'cell' 'material' 'universe' 'distribcell'
    bins: object                       # mesh id | energy edges | list of ids | [cell id]


@dataclass
class Tally:
    id: int
    filters: list                      # filter ids
    scores: list                       # names
    estimator: str = "tracklength"     # 'tracklength' | 'collision' | 'analog' (Tally::estimator_)


# ----------------------------------------------------------------------------
# region parsing
# ----------------------------------------------------------------------------
def _tokenize(spec: str):
    toks = []
    i = 0
    while i < len(spec):
        ch = spec[i]
        if ch == "(":
            toks.append(OP_LP); i += 1
        elif ch == ")":
            toks.append(OP_RP); i += 1
        elif ch == "|":
            toks.append(OP_UNION); i += 1
        elif ch == "~":
            toks.append(OP_COMP); i += 1
        elif ch in "+-" or ch.isdigit():
            j = i + 1
            while j < len(spec) and spec[j].isdigit():
                j += 1
            toks.append(int(spec[i:j])); i = j
        elif ch.isspace():
            i += 1
        else:
            raise ValueError(f"Region specification contains invalid character, {ch!r}")
    # insert implicit intersections (src/cell.cpp:600-612)
    i = 0
    while i < len(toks) - 1:
        left = toks[i] < OP_UNION or toks[i] == OP_RP
        right = toks[i + 1] < OP_UNION or toks[i + 1] in (OP_LP, OP_COMP)
        if left and right:
            toks.insert(i + 1, OP_INT)
        i += 1
    return toks


def _parse_region(toks):
    """Recursive descent over the infix tokens -> nested tuple tree.

    ('s', id) | ('and', [..]) | ('or', [..]); complements are pushed down with
    De Morgan's laws (src/cell.cpp:614-640, apply_demorgan :684-698)."""
    pos = 0

    def negate(node):
        if node[0] == "s":
            return ("s", -node[1])
        return ("or" if node[0] == "and" else "and", [negate(c) for c in node[1]])

    def primary():
        nonlocal pos
        t = toks[pos]
        if t == OP_COMP:
            pos += 1
            return negate(primary())
        if t == OP_LP:
            pos += 1
            node = union()
            assert toks[pos] == OP_RP, "unbalanced parentheses"
            pos += 1
            return node
        pos += 1
        return ("s", t)

    def inter():
        nonlocal pos
        items = [primary()]
        while pos < len(toks) and toks[pos] == OP_INT:
            pos += 1
            items.append(primary())
        return items[0] if len(items) == 1 else ("and", items)

    def union():
        nonlocal pos
        items = [inter()]
        while pos < len(toks) and toks[pos] == OP_UNION:
            pos += 1
            items.append(inter())
        return items[0] if len(items) == 1 else ("or", items)

    tree = union()
    assert pos == len(toks), "trailing tokens in region"
    return tree


def _emit(node, parent=None):
    """Tree -> infix tokens, parenthesised so that intersections bind tighter than
    unions (the effect of Region::add_precedence, src/cell.cpp:769-800)."""
    if node[0] == "s":
        return [node[1]]
    op = OP_INT if node[0] == "and" else OP_UNION
    out = []
    for k, c in enumerate(node[1]):
        if k:
            out.append(op)
        sub = _emit(c, node[0])
        if c[0] != "s" and c[0] != node[0]:
            sub = [OP_LP] + sub + [OP_RP]
        out += sub
    return out


def region_tokens(spec: str, surface_index: dict):
    """Returns (tokens, simple) with surface ids replaced by +-(index+1)."""
    if not spec or not spec.strip():
        return [], True
    tree = _parse_region(_tokenize(spec))

    def flatten(node):
        if node[0] == "s":
            return node
        kids = []
        for c in node[1]:
            c = flatten(c)
            if c[0] == node[0]:
                kids += c[1]
            else:
                kids.append(c)
        return (node[0], kids)

    tree = flatten(tree)
    toks = _emit(tree)
    simple = OP_UNION not in toks
    if simple:
        toks = [t for t in toks if t < OP_UNION]
    out = []
    for t in toks:
        if t < OP_UNION:
            idx = surface_index[abs(t)] + 1
            out.append(idx if t > 0 else -idx)
        else:
            out.append(t)
    return out, simple


def rotation_matrix(angles):
    """Cell::set_rotation, src/cell.cpp:60-85: the (inverse) rotation matrix for the angles
    (phi, theta, psi) in degrees, term for term (math.sin / cos are the C library's)."""
    import math
    phi = -angles[0] * math.pi / 180.0
    theta = -angles[1] * math.pi / 180.0
    psi = -angles[2] * math.pi / 180.0
    c, s_ = math.cos, math.sin
    return [c(theta) * c(psi),
            -c(phi) * s_(psi) + s_(phi) * s_(theta) * c(psi),
            s_(phi) * s_(psi) + c(phi) * s_(theta) * c(psi),
            c(theta) * s_(psi),
            c(phi) * c(psi) + s_(phi) * s_(theta) * s_(psi),
            -s_(phi) * c(psi) + c(phi) * s_(theta) * s_(psi),
            -s_(theta),
            s_(phi) * c(theta),
            c(phi) * c(theta)]


# ----------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------
class Model:
    """Geometry + materials + tallies; ``flatten`` produces omcgpu_model fields."""

    def __init__(self):
        self.surfaces: list[Surface] = []
        self.cells: list[Cell] = []
        self.lattices: list[RectLattice] = []
        self.materials: list[Material] = []
        self.meshes: list[Mesh] = []
        self.filters: list[Filter] = []
        self.tallies: list[Tally] = []
        self.settings = dict(run_CE=0, survival_biasing=0, weight_cutoff=0.25, weight_survive=1.0,
                             energy_cutoff=0.0, max_particle_events=1000000,
                             create_delayed_neutrons=1, free_gas_threshold=400.0,
                             n_log_bins=8000, legendre_to_tabular=True,
                             legendre_to_tabular_points=None, max_order=None,
                             temperature_tolerance=10.0)

    # -- convenience constructors ------------------------------------------------
    def surface(self, kind, coeffs, boundary="transmission", id=None, periodic_surface_id=None):
        s = Surface(id if id is not None else len(self.surfaces) + 1, kind, list(coeffs), boundary,
                    periodic_surface_id)
        self.surfaces.append(s)
        return s

    def material(self, **kw):
        kw.setdefault("id", len(self.materials) + 1)
        m = Material(**kw)
        self.materials.append(m)
        return m

    def cell(self, **kw):
        kw.setdefault("id", len(self.cells) + 1)
        c = Cell(**kw)
        if isinstance(c.fill, (RectLattice, HexLattice)):
            c.fill_type = "lattice"
        elif isinstance(c.fill, int):
            c.fill_type = "universe"
        else:
            c.fill_type = "material"
        self.cells.append(c)
        return c

    def lattice(self, **kw):
        kw.setdefault("id", len(self.lattices) + 1)
        l = RectLattice(**kw)
        self.lattices.append(l)
        return l

    def hex_lattice(self, **kw):
        kw.setdefault("id", len(self.lattices) + 1)
        l = HexLattice(**kw)
        self.lattices.append(l)
        return l

    def _periodic_pairs(self, s_index):
        """TranslationalPeriodicBC (src/boundary_condition.cpp:66-127): the translation between
        two parallel periodic planes, u (d2 - d1) with u the first plane's normal at the origin
        and d the signed distance of a plane from the origin along u.  Pairs come from
        periodic_surface_id, or the two unpaired periodic surfaces (src/surface.cpp:1279-1313)."""
        ns = len(self.surfaces)
        partner = np.full(max(ns, 1), -1, np.int32)
        shift = np.zeros((max(ns, 1), 3))
        axis = np.full(max(ns, 1), -1, np.int32)
        per = [s for s in self.surfaces if s.boundary == "periodic"]
        pairs = set()
        loose = []
        for s in per:
            if s.periodic_surface_id is not None:
                pairs.add((min(s.id, s.periodic_surface_id), max(s.id, s.periodic_surface_id)))
            else:
                loose.append(s.id)
        if len(loose) == 2:
            pairs.add((min(loose), max(loose)))
        elif loose:
            raise ValueError("periodic surfaces without a partner")

        def plane(s):
            if s.kind == "x-plane":
                return (1.0, 0.0, 0.0, s.coeffs[0])
            if s.kind == "y-plane":
                return (0.0, 1.0, 0.0, s.coeffs[0])
            if s.kind == "z-plane":
                return (0.0, 0.0, 1.0, s.coeffs[0])
            if s.kind == "plane":
                return tuple(s.coeffs)
            raise ValueError("translational periodic BCs need planes")

        def signed_distance(pl, u):
            # evaluate(origin) = -D; distance(origin, +-u): d = -f / projection, src/surface.cpp:207-220, 354-366
            a, b, c, d0 = pl
            e = -d0
            if abs(e) <= 1e-12:
                return 0.0
            sign = -1.0 if e > 1e-12 else 1.0       # direction -u (e > 0) or +u (e < 0)
            proj = sign * (a * u[0] + b * u[1] + c * u[2])
            dist = -e / proj
            if dist < 0.0:
                dist = float("inf")
            return -dist if e > 1e-12 else dist

        for lo, hi in sorted(pairs):
            i, j = s_index[lo], s_index[hi]
            p1, p2 = plane(self.surfaces[i]), plane(self.surfaces[j])
            n1 = np.array(p1[:3]) / math.sqrt(sum(c * c for c in p1[:3]))
            n2 = np.array(p2[:3]) / math.sqrt(sum(c * c for c in p2[:3]))
            if not abs(1.0 - float(n1 @ n2)) < 1e-14:
                # RotationalPeriodicBC (src/surface.cpp:1336-1373, src/boundary_condition.cpp:161-240): the
                # common zero component of the two normals is the axis; angle = (atan2 of the second normal
                # + pi) - atan2 of the first, in the (independent, dependent) plane of that axis
                zero = next((k for k in range(3) if abs(n1[k]) <= 1e-14 and abs(n2[k]) <= 1e-14), None)
                if zero is None or abs(p1[3]) > 1e-12 or abs(p2[3]) > 1e-12:
                    raise ValueError("rotational periodic BCs need two planes through the origin with a common zero normal component")
                a1, a2 = (1, 2) if zero == 0 else ((0, 2) if zero == 1 else (0, 1))
                m1, m2 = p1[:3], p2[:3]          # normal({0,0,0}) of the planes, not normalised
                angle = (math.atan2(m2[a2], m2[a1]) + math.pi) - math.atan2(m1[a2], m1[a1])
                partner[i], partner[j] = j, i
                axis[i] = axis[j] = zero
                shift[i] = [math.cos(angle), math.sin(angle), 0.0]
                shift[j] = [math.cos(-angle), math.sin(-angle), 0.0]
                continue
            u = p1[:3]
            d1, d2 = signed_distance(p1, u), signed_distance(p2, u)
            t = [u[k] * (d2 - d1) for k in range(3)]
            partner[i], partner[j] = j, i
            shift[i] = t
            shift[j] = [-x for x in t]
        return partner, shift, axis

    # -- flattening --------------------------------------------------------------
    def flatten_geometry(self) -> dict:
        v = {}
        s_index = {s.id: i for i, s in enumerate(self.surfaces)}
        ns = len(self.surfaces)
        coef = np.zeros((max(ns, 1), abi.OMCGPU_SURF_NCOEF))
        stype = np.zeros(max(ns, 1), np.int32)
        sbc = np.zeros(max(ns, 1), np.int32)
        quads = []
        for i, s in enumerate(self.surfaces):
            kind, nco = _SURF_KINDS[s.kind]
            if len(s.coeffs) != nco:
                raise ValueError(f"surface {s.id}: {s.kind} needs {nco} coefficients")
            stype[i] = kind
            if kind == abi.OMCGPU_SURF_QUADRIC:       # row of the side table
                coef[i, 0] = len(quads)
                quads.append([float(x) for x in s.coeffs])
            else:
                coef[i, :nco] = s.coeffs
            sbc[i] = _BCS[s.boundary]
        partner, shift, paxis = self._periodic_pairs({s.id: i for i, s in enumerate(self.surfaces)})
        v.update(n_surfaces=ns, surf_type=stype, surf_bc=sbc, surf_coef=coef, n_quadrics=len(quads),
                 quad_coef=np.array(quads or [[0.0]]).reshape(-1), surf_periodic=partner,
                 surf_periodic_shift=shift, surf_periodic_axis=paxis,
                 surf_id=np.array([s.id for s in self.surfaces] or [0], np.int32))

        # universes: ids in order of first appearance (cells' universe attribute,
        # fills and lattice entries), src/geometry_aux.cpp
        univ_ids = []

        def add_univ(u):
            if u not in univ_ids:
                univ_ids.append(u)
        for c in self.cells:
            add_univ(c.universe)
        for c in self.cells:
            if c.fill_type == "universe":
                add_univ(c.fill)
        for l in self.lattices:
            for u in np.asarray(l.universes).reshape(-1):
                if int(u) >= 0:                      # (hex lattices: -1 outside the hexagon)
                    add_univ(int(u))
            if l.outer is not None:
                add_univ(l.outer)
        u_index = {u: i for i, u in enumerate(univ_ids)}
        l_index = {l.id: i for i, l in enumerate(self.lattices)}
        m_index = {m.id: i for i, m in enumerate(self.materials)}
        self._u_index, self._m_index, self._l_index = u_index, m_index, l_index
        self._c_index = {c.id: i for i, c in enumerate(self.cells)}

        nc = len(self.cells)
        ctype = np.zeros(nc, np.int32); cuniv = np.zeros(nc, np.int32)
        cfill = np.full(nc, -1, np.int32); csimple = np.zeros(nc, np.int32)
        rofs = [0]; rtok = []
        mofs = [0]; cmat = []
        kofs = [0]; ckt = []
        trans = np.zeros((nc, 3))
        rotated = np.zeros(nc, np.int32); rot = np.zeros((nc, 9))
        for i, c in enumerate(self.cells):
            toks, simple = region_tokens(c.region, s_index)
            rtok += toks; rofs.append(len(rtok)); csimple[i] = int(simple)
            cuniv[i] = u_index[c.universe]
            trans[i] = c.translation
            if c.rotation is not None:
                if c.fill_type == "material":
                    raise ValueError(f"cell {c.id}: cannot rotate a material cell")
                rotated[i] = 1
                rot[i] = rotation_matrix(c.rotation)
            if c.fill_type == "material":
                ctype[i] = abi.OMCGPU_FILL_MATERIAL
                cmat.append(-1 if c.fill is None else m_index[c.fill.id])
                T = c.temperature
                if T is None and c.fill is not None and c.fill.temperature is not None:
                    T = c.fill.temperature
                if T is None:
                    T = TEMPERATURE_DEFAULT
                ckt.append(math.sqrt(K_BOLTZMANN * T))
            elif c.fill_type == "universe":
                ctype[i] = abi.OMCGPU_FILL_UNIVERSE
                cfill[i] = u_index[c.fill]
            else:
                ctype[i] = abi.OMCGPU_FILL_LATTICE
                cfill[i] = l_index[c.fill.id]
            mofs.append(len(cmat)); kofs.append(len(ckt))
        v.update(n_cells=nc, cell_id=np.array([c.id for c in self.cells], np.int32),
                 cell_type=ctype, cell_universe=cuniv, cell_fill=cfill, cell_simple=csimple,
                 cell_region_ofs=np.array(rofs, np.int32), region_tokens=np.array(rtok or [0], np.int32),
                 cell_mat_ofs=np.array(mofs, np.int32), cell_material=np.array(cmat or [0], np.int32),
                 cell_sqrtkT_ofs=np.array(kofs, np.int32), cell_sqrtkT=np.array(ckt or [0.0]),
                 # finalize_cell_densities (src/geometry_aux.cpp:197-211): {1.0} unless the cell gives
                 # densities; those are divided by the material's density by the CE builder
                 cell_density_ofs=np.arange(nc + 1, dtype=np.int32), cell_density_mult=np.ones(max(nc, 1)),
                 cell_translation=trans, cell_rotated=rotated, cell_rotation=rot,
                 cell_distribcell_index=np.full(nc, -1, np.int32),
                 cell_offset_ofs=np.zeros(nc + 1, np.int32), cell_offset=np.zeros(1, np.int32))

        nu = len(univ_ids)
        uofs = [0]; ucells = []
        for u in univ_ids:
            ucells += [i for i, c in enumerate(self.cells) if c.universe == u]
            uofs.append(len(ucells))
        v.update(n_universes=nu, univ_cell_ofs=np.array(uofs, np.int32),
                 univ_cells=np.array(ucells, np.int32))

        nl = len(self.lattices)
        lshape = np.ones((max(nl, 1), 3), np.int32); l3d = np.zeros(max(nl, 1), np.int32)
        louter = np.full(max(nl, 1), -1, np.int32)
        lpitch = np.zeros((max(nl, 1), 3)); lll = np.zeros((max(nl, 1), 3))
        lofs = [0]; lun = []
        ltype = np.zeros(max(nl, 1), np.int32)
        for i, l in enumerate(self.lattices):
            if isinstance(l, HexLattice):
                n = 2 * l.n_rings - 1
                arr = np.asarray(l.universes, dtype=np.int64).reshape(-1, n, n)
                ltype[i] = abi.OMCGPU_LATTICE_HEX_Y if l.orientation == "y" else abi.OMCGPU_LATTICE_HEX_X
                l3d[i] = int(l.n_axial is not None)
                lshape[i] = (n, n, arr.shape[0])
                lpitch[i, 0] = l.pitch[0]
                if l3d[i]:
                    lpitch[i, 1] = l.pitch[1]
                lll[i, :len(l.center)] = l.center       # the lattice centre
                if l.outer is not None:
                    louter[i] = u_index[l.outer]
                for z in range(arr.shape[0]):
                    for i1 in range(n):
                        for i0 in range(n):
                            if l.valid(i0, i1) != (arr[z, i1, i0] >= 0):
                                raise ValueError(f"hex lattice {l.id}: universes must be >= 0 exactly inside the hexagon")
                lun += [u_index[int(u)] if u >= 0 else -1 for u in arr.reshape(-1)]
                lofs.append(len(lun))
                continue
            arr = np.asarray(l.universes)
            if arr.ndim == 2:
                arr = arr[None]
            else:
                l3d[i] = 1
            nz, ny, nx = arr.shape
            lshape[i] = (nx, ny, nz)
            nd = 3 if l3d[i] else 2
            lpitch[i, :nd] = l.pitch[:nd]
            lll[i, :nd] = l.lower_left[:nd]
            if l.outer is not None:
                louter[i] = u_index[l.outer]
            # rows are listed top-to-bottom: flip y (src/lattice.cpp:214-231)
            flipped = arr[:, ::-1, :]
            lun += [u_index[int(u)] for u in flipped.reshape(-1)]
            lofs.append(len(lun))
        v.update(n_lattices=nl, lat_shape=lshape, lat_is3d=l3d, lat_type=ltype, lat_outer=louter, lat_pitch=lpitch,
                 lat_lower_left=lll, lat_univ_ofs=np.array(lofs, np.int32),
                 lat_universes=np.array(lun or [0], np.int32),
                 lat_offset_ofs=np.zeros(max(nl, 1) + 1, np.int32),
                 lat_offsets=np.zeros(1, np.int32))

        # root universe = the one no cell / lattice is filled with; depth
        filled = set()
        for c in self.cells:
            if c.fill_type == "universe":
                filled.add(c.fill)
        for l in self.lattices:
            filled |= {int(u) for u in np.asarray(l.universes).reshape(-1) if int(u) >= 0}
            if l.outer is not None:
                filled.add(l.outer)
        roots = [u for u in univ_ids if u not in filled]
        if len(roots) != 1:
            raise ValueError(f"expected exactly one root universe, got {roots}")
        v["root_universe"] = u_index[roots[0]]

        def depth(u):
            d = 1
            for c in self.cells:
                if c.universe != u:
                    continue
                if c.fill_type == "universe":
                    d = max(d, 1 + depth(c.fill))
                elif c.fill_type == "lattice":
                    subs = {int(x) for x in np.asarray(c.fill.universes).reshape(-1) if int(x) >= 0}
                    if c.fill.outer is not None:
                        subs.add(c.fill.outer)
                    d = max(d, 1 + max(depth(s) for s in subs))
            return d
        v["n_coord_levels"] = depth(roots[0])
        if v["n_coord_levels"] > abi.OMCGPU_MAX_COORD:
            raise ValueError("geometry nests deeper than OMCGPU_MAX_COORD")
        self._prepare_distribcell(v)
        return v

    def _prepare_distribcell(self, v):
        """prepare_distribcell + count_universe_instances, src/geometry_aux.cpp:327-500,
        src/lattice.cpp:106-129: distribcell maps for the cells named by DistribcellFilters
        (and, with settings['material_cell_offsets'], for every material cell), cell / lattice
        offset tables, instances per cell."""
        nc, nu, nl = v["n_cells"], v["n_universes"], v["n_lattices"]
        ctype, cfill = v["cell_type"], v["cell_fill"]
        uofs, ucells = v["univ_cell_ofs"], v["univ_cells"]
        lofs, lun = v["lat_univ_ofs"], v["lat_universes"]

        def cells_of(u):
            return ucells[uofs[u]:uofs[u + 1]]

        def count(search, target, memo):       # count_universe_instances :462-500
            if search == target:
                return 1
            if search in memo:
                return memo[search]
            n = 0
            for c in cells_of(search):
                if ctype[c] == abi.OMCGPU_FILL_UNIVERSE:
                    n += count(int(cfill[c]), target, memo)
                elif ctype[c] == abi.OMCGPU_FILL_LATTICE:
                    l = int(cfill[c])
                    for u in lun[lofs[l]:lofs[l + 1]]:
                        if u >= 0:                   # LatticeIter skips the empty positions of a hex lattice
                            n += count(int(u), target, memo)
            memo[search] = n
            return n

        root = v["root_universe"]
        univ_instances = [count(root, u, {}) for u in range(nu)]
        self._n_instances = [univ_instances[v["cell_universe"][c]] for c in range(nc)]

        targets = {self._c_index[f.bins[0]] for f in self.filters if f.type == "distribcell"}
        if self.settings.get("material_cell_offsets"):
            targets |= {c for c in range(nc) if ctype[c] == abi.OMCGPU_FILL_MATERIAL}
        dindex = np.full(nc, -1, np.int32)
        target_univs = []
        for u in range(nu):
            for c in cells_of(u):
                if int(c) in targets:
                    if u not in target_univs:
                        target_univs.append(u)
                    dindex[c] = target_univs.index(u)
        n_maps = len(target_univs)
        v["cell_distribcell_index"] = dindex
        if n_maps == 0:
            return
        # offset tables
        cofs = [0]
        for c in range(nc):
            cofs.append(cofs[-1] + (n_maps if ctype[c] != abi.OMCGPU_FILL_MATERIAL else 0))
        coff = np.full(max(cofs[-1], 1), -1, np.int32)
        lat_size = [int(lofs[l + 1] - lofs[l]) for l in range(nl)]
        loofs = [0]
        for l in range(nl):
            loofs.append(loofs[-1] + n_maps * lat_size[l])
        loff = np.full(max(loofs[-1], 1), -1, np.int32)
        for mp, tu in enumerate(target_univs):
            memo = {}
            for u in range(nu):
                offset = 0
                for c in cells_of(u):
                    if ctype[c] == abi.OMCGPU_FILL_UNIVERSE:
                        coff[cofs[c] + mp] = offset
                        offset += count(int(cfill[c]), tu, memo)
                    elif ctype[c] == abi.OMCGPU_FILL_LATTICE:
                        coff[cofs[c] + mp] = offset
                        l = int(cfill[c])
                        o = 0                   # Lattice::fill_offset_table
                        for k, uu in enumerate(lun[lofs[l]:lofs[l + 1]]):
                            if uu < 0:
                                continue
                            loff[loofs[l] + mp * lat_size[l] + k] = o
                            o += count(int(uu), tu, memo)
                        offset += o
        v["cell_offset_ofs"] = np.array(cofs, np.int32)
        v["cell_offset"] = coff
        v["lat_offset_ofs"] = np.array(loofs + [loofs[-1]] * (max(nl, 1) + 1 - len(loofs)), np.int32)
        v["lat_offsets"] = loff

    def flatten_tallies(self, n_groups=0, group_edges=None) -> dict:
        v = {}
        if not hasattr(self, "_n_instances"):
            self.flatten_geometry()
        nm = len(self.meshes)
        mnd = np.zeros(max(nm, 1), np.int32); msh = np.ones((max(nm, 1), 3), np.int32)
        mll = np.zeros((max(nm, 1), 3)); mw = np.ones((max(nm, 1), 3))
        mesh_index = {}
        for i, m in enumerate(self.meshes):
            nd = len(m.dimension)
            mnd[i] = nd; msh[i, :nd] = m.dimension; mll[i, :nd] = m.lower_left
            # width = (upper_right - lower_left) / shape  (src/mesh.cpp RegularMesh ctor)
            for k in range(nd):
                mw[i, k] = (m.upper_right[k] - m.lower_left[k]) / m.dimension[k]
            mesh_index[m.id] = i
        v.update(n_meshes=nm, mesh_ndim=mnd, mesh_shape=msh, mesh_lower_left=mll, mesh_width=mw)

        nf = len(self.filters)
        ftype = np.zeros(max(nf, 1), np.int32); fnb = np.zeros(max(nf, 1), np.int32)
        fmesh = np.full(max(nf, 1), -1, np.int32); fmg = np.zeros(max(nf, 1), np.int32)
        fofs = [0]; fi = []; fd = []
        f_index = {}
        for i, f in enumerate(self.filters):
            f_index[f.id] = i
            if f.type == "mesh":
                ftype[i] = abi.OMCGPU_FILTER_MESH
                fmesh[i] = mesh_index[f.bins]
                fnb[i] = int(np.prod(self.meshes[fmesh[i]].dimension))
            elif f.type in ("energy", "energyout"):
                ftype[i] = abi.OMCGPU_FILTER_ENERGY if f.type == "energy" else abi.OMCGPU_FILTER_ENERGYOUT
                edges = [float(e) for e in f.bins]
                fnb[i] = len(edges) - 1
                fd += edges; fi += [0] * len(edges)
                # EnergyFilter::set_bins: matches the transport group structure?
                if group_edges is not None and len(edges) == n_groups + 1:
                    fmg[i] = int(all(e == g for e, g in zip(edges, sorted(group_edges))))
            elif f.type == "distribcell":
                # DistribcellFilter::set_cell: one bin per instance of the cell
                ftype[i] = abi.OMCGPU_FILTER_DISTRIBCELL
                (cid,) = f.bins
                fnb[i] = int(self._n_instances[self._c_index[cid]])
                fi += [self._c_index[cid]]; fd += [0.0]
            else:
                kind = {"cell": (abi.OMCGPU_FILTER_CELL, self._c_index, len(self.cells)),
                        "material": (abi.OMCGPU_FILTER_MATERIAL, self._m_index, len(self.materials)),
                        "universe": (abi.OMCGPU_FILTER_UNIVERSE, self._u_index, len(self._u_index))}[f.type]
                ftype[i] = kind[0]
                mp = [-1] * kind[2]
                for b, ident in enumerate(f.bins):
                    mp[kind[1][ident]] = b
                fnb[i] = len(f.bins)
                fi += mp; fd += [0.0] * len(mp)
            fofs.append(len(fi))
        v.update(n_filters=nf, filter_type=ftype, filter_n_bins=fnb, filter_mesh=fmesh,
                 filter_matches_groups=fmg, filter_data_ofs=np.array(fofs, np.int32),
                 filter_idata=np.array(fi or [0], np.int32), filter_ddata=np.array(fd or [0.0]))

        nt = len(self.tallies)
        tnf = np.zeros(max(nt, 1), np.int32); tf = np.zeros((max(nt, 1), abi.OMCGPU_MAX_FILTERS), np.int32)
        tst = np.zeros((max(nt, 1), abi.OMCGPU_MAX_FILTERS), np.int32)
        tns = np.zeros(max(nt, 1), np.int32); tsc = np.zeros((max(nt, 1), abi.OMCGPU_MAX_SCORES), np.int32)
        tofs = [0]
        test = np.full(max(nt, 1), abi.OMCGPU_ESTIMATOR_TRACKLENGTH, np.int32)
        for i, t in enumerate(self.tallies):
            test[i] = {"tracklength": abi.OMCGPU_ESTIMATOR_TRACKLENGTH, "collision": abi.OMCGPU_ESTIMATOR_COLLISION,
                       "analog": abi.OMCGPU_ESTIMATOR_ANALOG}[t.estimator]
            tnf[i] = len(t.filters)
            stride = 1
            for k in range(len(t.filters) - 1, -1, -1):   # src/tallies/tally.cpp:503-516
                fidx = f_index[t.filters[k]]
                tf[i, k] = fidx
                tst[i, k] = stride
                stride *= int(fnb[fidx])
            tns[i] = len(t.scores)
            tsc[i, :len(t.scores)] = [_SCORES[s] for s in t.scores]
            tofs.append(tofs[-1] + stride * len(t.scores))
        v.update(n_tallies=nt, tally_estimator=test, tally_n_filters=tnf, tally_filters=tf, tally_strides=tst,
                 tally_n_scores=tns, tally_scores=tsc, tally_result_ofs=np.array(tofs, np.int64))
        return v

    def tally_shape(self, i):
        t = self.tallies[i]
        fmap = {f.id: f for f in self.filters}
        nb = 1
        for fid in t.filters:
            f = fmap[fid]
            if f.type == "mesh":
                m = [m for m in self.meshes if m.id == f.bins][0]
                nb *= int(np.prod(m.dimension))
            elif f.type in ("energy", "energyout"):
                nb *= len(f.bins) - 1
            elif f.type == "distribcell":
                nb *= int(self._n_instances[self._c_index[f.bins[0]]])
            else:
                nb *= len(f.bins)
        return nb, len(t.scores)

    def settings_fields(self) -> dict:
        s = self.settings
        return dict(run_CE=s["run_CE"], survival_biasing=int(s["survival_biasing"]),
                    max_particle_events=s["max_particle_events"],
                    create_delayed_neutrons=int(s["create_delayed_neutrons"]),
                    weight_cutoff=s["weight_cutoff"], weight_survive=s["weight_survive"],
                    energy_cutoff=s["energy_cutoff"], free_gas_threshold=s["free_gas_threshold"])


def normalize_density(mat: Material, awr: Sequence[float]):
    """Material::normalize_density, src/material.cpp:417-467 -> atom densities."""
    dens = [float(d) for _, d in mat.components]
    if mat.units == "macro":
        density = 1.0 if mat.density is None else float(mat.density)
        dens = [density for _ in dens]
    elif mat.units == "sum":
        density = 0.0
        for d in dens:
            density += d
    elif mat.units in ("g/cc", "g/cm3"):
        density = -float(mat.density)
    elif mat.units == "kg/m3":
        density = -1.0e-3 * float(mat.density)
    elif mat.units == "atom/b-cm":
        density = float(mat.density)
    elif mat.units in ("atom/cc", "atom/cm3"):
        density = 1.0e-24 * float(mat.density)
    else:
        raise ValueError(f"Unknown units {mat.units!r}")
    percent_in_atom = dens[0] >= 0.0
    density_in_atom = density >= 0.0
    if not percent_in_atom:
        dens = [-d / a for d, a in zip(dens, awr)]
    total = 0.0
    for d in dens:
        total += d
    dens = [d / total for d in dens]
    if not density_in_atom:
        sum_percent = 0.0
        for d, a in zip(dens, awr):
            sum_percent += d * a
        sum_percent = 1.0 / sum_percent
        density = -density * N_AVOGADRO / MASS_NEUTRON * sum_percent
    return [d * density for d in dens]
