// omcgpu_host.cpp -- the libopenmc side of the drop-in (C++, compiled INTO libopenmc next to
// its own sources; see omcgpu_host.h and INTEGRATION.md).
//
//   flatten_model()          copies model:: / data:: / settings:: -- what read_model_xml,
//                            finalize_geometry and finalize_cross_sections have built -- into the
//                            POD arrays of omcgpu_model (include/omcgpu.h)
//   simulation_init()        flatten + omcgpu_init: stands where init_event_queues() stands,
//                            /root/reference/src/simulation.cpp:99-105
//   transport_generation()   one generation on the device in place of transport_event_based()
//                            / transport_history_based(), src/simulation.cpp:255-259; leaves the
//                            host what event_death() (src/particle.cpp:475-512), the tally scoring
//                            and sort_fission_bank() + synchronize_bank() leave: global_tally_*,
//                            results_(:,:,VALUE), n_lost_particles, the NEXT simulation::source_bank
//   bank_synchronized()      tells finalize_generation() (src/simulation.cpp:546-552) that the bank
//                            sort and the comb were done on the device
//   simulation_finalize()    omcgpu_finalize, next to free_event_queues(), src/finalize.cpp:57-59
//
// The device library is loaded with dlopen (OMCGPU_LIB, default libomcgpu.so next to this
// library): libopenmc keeps building and running without CUDA, and `OPENMC_GPU=1` selects
// the path at run time.  Model features the device path refuses end in fatal_error -- there
// is no CPU fallback behind the switch.
//
// Several of the reference's classes keep the members the flattener needs private or
// protected (Region::expression_, Mgxs::xs, the filters' bins, ...).  An in-tree
// integration would add `friend` declarations; this translation unit, which must compile
// against the UNMODIFIED headers, opens them with the two defines below (layout unchanged).
#include <algorithm>
#include <array>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include <dlfcn.h>

#include "xtensor/xarray.hpp"
#include "xtensor/xtensor.hpp"
#include <fmt/format.h>

#define private public
#define protected public
#include "openmc/timer.h" // Timer::elapsed_: the device's kernel times are added to the event timers
#include "openmc/bank.h"
#include "openmc/boundary_condition.h"
#include "openmc/cell.h"
#include "openmc/distribution_energy.h"
#include "openmc/eigenvalue.h"
#include "openmc/endf.h"
#include "openmc/geometry.h"
#include "openmc/lattice.h"
#include "openmc/material.h"
#include "openmc/mesh.h"
#include "openmc/mgxs.h"
#include "openmc/mgxs_interface.h"
#include "openmc/nuclide.h"
#include "openmc/reaction.h"
#include "openmc/reaction_product.h"
#include "openmc/scattdata.h"
#include "openmc/secondary_uncorrelated.h"
#include "openmc/surface.h"
#include "openmc/tallies/filter.h"
#include "openmc/tallies/filter_cell.h"
#include "openmc/tallies/filter_distribcell.h"
#include "openmc/tallies/filter_energy.h"
#include "openmc/tallies/filter_material.h"
#include "openmc/tallies/filter_mesh.h"
#include "openmc/tallies/filter_universe.h"
#include "openmc/tallies/tally.h"
#include "openmc/universe.h"
#include "openmc/xsdata.h"
#undef private
#undef protected

#include "openmc/capi.h"
#include "openmc/constants.h"
#include "openmc/error.h"
#include "openmc/message_passing.h"
#include "openmc/settings.h"
#include "openmc/simulation.h"
#include "openmc/timer.h"

#include "omcgpu.h"
#include "omcgpu_host.h"

namespace openmc {
namespace gpu {

static_assert(sizeof(SourceSite) == sizeof(omcgpu_site), "SourceSite and omcgpu_site must be bit-compatible");

//==============================================================================
// the flattened model and its storage
//==============================================================================
namespace {

struct Flat {
  omcgpu_model m {};
  std::deque<std::vector<int32_t>> iv;
  std::deque<std::vector<int64_t>> lv;
  std::deque<std::vector<double>> dv;
  std::map<std::string, int64_t> length; // element counts, for omcgpu_host_model_array

  const int32_t* keep(const char* name, std::vector<int32_t>&& v)
  {
    length[name] = v.size();
    if (v.empty())
      v.push_back(0);
    iv.push_back(std::move(v));
    return iv.back().data();
  }
  const int64_t* keep(const char* name, std::vector<int64_t>&& v)
  {
    length[name] = v.size();
    if (v.empty())
      v.push_back(0);
    lv.push_back(std::move(v));
    return lv.back().data();
  }
  const double* keep(const char* name, std::vector<double>&& v)
  {
    length[name] = v.size();
    if (v.empty())
      v.push_back(0.0);
    dv.push_back(std::move(v));
    return dv.back().data();
  }
};
#define KEEP(field, vec) F.m.field = F.keep(#field, std::move(vec))

std::unique_ptr<Flat> g_flat;
std::string g_error;

[[noreturn]] void unsupported(const std::string& what)
{
  throw std::runtime_error("OPENMC_GPU: " + what + " is not supported by the device path");
}

using IV = std::vector<int32_t>;
using LV = std::vector<int64_t>;
using DV = std::vector<double>;

//------------------------------------------------------------------------------
void flatten_settings(Flat& F)
{
  auto& m = F.m;
  m.run_CE = settings::run_CE ? 1 : 0;
  m.n_coord_levels = model::n_coord_levels;
  m.root_universe = model::root_universe;
  m.survival_biasing = settings::survival_biasing ? 1 : 0;
  m.max_particle_events = settings::max_particle_events;
  m.create_delayed_neutrons = settings::create_delayed_neutrons ? 1 : 0;
  m.weight_cutoff = settings::weight_cutoff;
  m.weight_survive = settings::weight_survive;
  m.energy_cutoff = settings::energy_cutoff[0];
  m.free_gas_threshold = settings::free_gas_threshold;
  if (settings::photon_transport)
    unsupported("photon transport");
  if (settings::run_mode != RunMode::EIGENVALUE)
    unsupported("a run mode other than eigenvalue");
  if (settings::ufs_on || settings::weight_windows_on || settings::ifp_on)
    unsupported("uniform fission site weighting / weight windows / IFP");
  // what the reference does per particle on the HOST while it transports it, and the device path does
  // not do: refused by name rather than silently left out
  if (settings::surf_source_write || settings::surf_source_read)
    unsupported("surface source files (surf_source_write / surf_source_read)");
  if (settings::collision_track)
    unsupported("the collision track file");
  if (settings::write_all_tracks || !settings::track_identifiers.empty() || settings::trace_batch > 0)
    unsupported("particle tracks / particle tracing");
  if (settings::check_overlaps)
    unsupported("the cell overlap check");
  if (settings::res_scat_on)
    unsupported("resonance upscattering (res_scat)");
  if (settings::time_cutoff[0] < INFTY)
    unsupported("a neutron time cutoff");
  if (model::n_coord_levels > OMCGPU_MAX_COORD)
    unsupported("geometry nested deeper than OMCGPU_MAX_COORD levels");
}

//------------------------------------------------------------------------------
// model::surfaces (src/surface.cpp) and their boundary conditions
void flatten_surfaces(Flat& F)
{
  const int ns = model::surfaces.size();
  IV type(ns), bc(ns), partner(ns, -1), id(ns), paxis(ns, -1);
  DV coef((size_t)ns * OMCGPU_SURF_NCOEF, 0.0), shift((size_t)ns * 3, 0.0), quads;
  for (int i = 0; i < ns; ++i) {
    const Surface* s = model::surfaces[i].get();
    double* c = coef.data() + (size_t)i * OMCGPU_SURF_NCOEF;
    id[i] = s->id_;
    if (auto p = dynamic_cast<const SurfaceXPlane*>(s)) {
      type[i] = OMCGPU_SURF_XPLANE; c[0] = p->x0_;
    } else if (auto p = dynamic_cast<const SurfaceYPlane*>(s)) {
      type[i] = OMCGPU_SURF_YPLANE; c[0] = p->y0_;
    } else if (auto p = dynamic_cast<const SurfaceZPlane*>(s)) {
      type[i] = OMCGPU_SURF_ZPLANE; c[0] = p->z0_;
    } else if (auto p = dynamic_cast<const SurfacePlane*>(s)) {
      type[i] = OMCGPU_SURF_PLANE; c[0] = p->A_; c[1] = p->B_; c[2] = p->C_; c[3] = p->D_;
    } else if (auto p = dynamic_cast<const SurfaceXCylinder*>(s)) {
      type[i] = OMCGPU_SURF_XCYLINDER; c[0] = p->y0_; c[1] = p->z0_; c[2] = p->radius_;
    } else if (auto p = dynamic_cast<const SurfaceYCylinder*>(s)) {
      type[i] = OMCGPU_SURF_YCYLINDER; c[0] = p->x0_; c[1] = p->z0_; c[2] = p->radius_;
    } else if (auto p = dynamic_cast<const SurfaceZCylinder*>(s)) {
      type[i] = OMCGPU_SURF_ZCYLINDER; c[0] = p->x0_; c[1] = p->y0_; c[2] = p->radius_;
    } else if (auto p = dynamic_cast<const SurfaceSphere*>(s)) {
      type[i] = OMCGPU_SURF_SPHERE; c[0] = p->x0_; c[1] = p->y0_; c[2] = p->z0_; c[3] = p->radius_;
    } else if (auto p = dynamic_cast<const SurfaceXCone*>(s)) {
      type[i] = OMCGPU_SURF_XCONE; c[0] = p->x0_; c[1] = p->y0_; c[2] = p->z0_; c[3] = p->radius_sq_;
    } else if (auto p = dynamic_cast<const SurfaceYCone*>(s)) {
      type[i] = OMCGPU_SURF_YCONE; c[0] = p->x0_; c[1] = p->y0_; c[2] = p->z0_; c[3] = p->radius_sq_;
    } else if (auto p = dynamic_cast<const SurfaceZCone*>(s)) {
      type[i] = OMCGPU_SURF_ZCONE; c[0] = p->x0_; c[1] = p->y0_; c[2] = p->z0_; c[3] = p->radius_sq_;
    } else if (auto p = dynamic_cast<const SurfaceQuadric*>(s)) {
      type[i] = OMCGPU_SURF_QUADRIC;
      c[0] = quads.size() / 10; // row of the side table
      const double q[10] = {p->A_, p->B_, p->C_, p->D_, p->E_, p->F_, p->G_, p->H_, p->J_, p->K_};
      quads.insert(quads.end(), q, q + 10);
    } else {
      unsupported(fmt::format("surface {} (torus / DAGMC)", s->id_));
    }
    const BoundaryCondition* b = s->bc_.get();
    if (!b) {
      bc[i] = OMCGPU_BC_NONE;
    } else if (b->has_albedo()) {
      unsupported(fmt::format("the albedo of surface {}", s->id_));
    } else if (dynamic_cast<const VacuumBC*>(b)) {
      bc[i] = OMCGPU_BC_VACUUM;
    } else if (dynamic_cast<const ReflectiveBC*>(b)) {
      bc[i] = OMCGPU_BC_REFLECTIVE;
    } else if (dynamic_cast<const WhiteBC*>(b)) {
      bc[i] = OMCGPU_BC_WHITE;
    } else if (auto p = dynamic_cast<const TranslationalPeriodicBC*>(b)) {
      // handle_particle (src/boundary_condition.cpp:101-127): + translation_ when the first
      // surface of the pair is struck, - translation_ on the second
      bc[i] = OMCGPU_BC_PERIODIC;
      const bool first = p->i_surf_ == i;
      partner[i] = first ? p->j_surf_ : p->i_surf_;
      for (int k = 0; k < 3; ++k)
        shift[3 * i + k] = first ? p->translation_[k] : -p->translation_[k];
    } else if (auto p = dynamic_cast<const RotationalPeriodicBC*>(b)) {
      // handle_particle (src/boundary_condition.cpp:242-287): rotation by +angle_ when the first surface of
      // the pair is struck, -angle_ on the second; cos / sin taken here with the host's libm, as there
      bc[i] = OMCGPU_BC_PERIODIC;
      const bool first = p->i_surf_ == i;
      partner[i] = first ? p->j_surf_ : p->i_surf_;
      const double theta = first ? p->angle_ : -p->angle_;
      shift[3 * i + 0] = std::cos(theta);
      shift[3 * i + 1] = std::sin(theta);
      paxis[i] = p->zero_axis_idx_;
    } else {
      unsupported(fmt::format("the boundary condition of surface {}", s->id_));
    }
  }
  F.m.n_surfaces = ns;
  KEEP(surf_type, type);
  KEEP(surf_bc, bc);
  KEEP(surf_periodic, partner);
  KEEP(surf_periodic_shift, shift);
  KEEP(surf_periodic_axis, paxis);
  KEEP(surf_id, id);
  KEEP(surf_coef, coef);
  F.m.n_quadrics = quads.size() / 10;
  KEEP(quad_coef, quads);
}

//------------------------------------------------------------------------------
// model::cells, model::universes, model::lattices
void flatten_cells(Flat& F)
{
  const int nc = model::cells.size();
  IV id(nc), type(nc), univ(nc), fill(nc, -1), simple(nc), rofs {0}, rtok, mofs {0}, cmat, kofs {0}, rotated(nc, 0),
    dindex(nc, -1), oofs {0}, off;
  DV ckt, cden, trans((size_t)nc * 3, 0.0), rot((size_t)nc * 9, 0.0);
  IV dofs {0};
  int n_maps = 0;
  for (int i = 0; i < nc; ++i)
    n_maps = std::max(n_maps, model::cells[i]->distribcell_index_ + 1);
  for (int i = 0; i < nc; ++i) {
    const Cell& c = *model::cells[i];
    auto csg = dynamic_cast<const CSGCell*>(&c);
    if (!csg)
      unsupported(fmt::format("cell {} (not a CSG cell)", c.id_));
    id[i] = c.id_;
    univ[i] = c.universe_;
    simple[i] = csg->region_.simple_ ? 1 : 0;
    for (int32_t t : csg->region_.expression_) {
      if (t >= OP_UNION) { // the five operator tokens count down from INT32_MAX (include/openmc/cell.h:31-35)
        rtok.push_back(t == OP_LEFT_PAREN ? OMCGPU_OP_LEFT_PAREN
                       : t == OP_RIGHT_PAREN ? OMCGPU_OP_RIGHT_PAREN
                       : t == OP_COMPLEMENT ? OMCGPU_OP_COMPLEMENT
                       : t == OP_INTERSECTION ? OMCGPU_OP_INTERSECTION
                                              : OMCGPU_OP_UNION);
      } else {
        rtok.push_back(t);
      }
    }
    rofs.push_back(rtok.size());
    for (int k = 0; k < 3; ++k)
      trans[3 * i + k] = c.translation_[k];
    if (!c.rotation_.empty()) {
      rotated[i] = 1;
      for (int k = 0; k < 9; ++k)
        rot[9 * i + k] = c.rotation_[k];
    }
    for (double d : c.density_mult_) // {1.0} unless the cell gives densities (finalize_cell_densities)
      cden.push_back(d);
    dofs.push_back(cden.size());
    if (c.type_ == Fill::MATERIAL) {
      type[i] = OMCGPU_FILL_MATERIAL;
      for (int32_t mt : c.material_)
        cmat.push_back(mt == MATERIAL_VOID ? -1 : mt);
      for (double s : c.sqrtkT_)
        ckt.push_back(s);
    } else if (c.type_ == Fill::UNIVERSE) {
      type[i] = OMCGPU_FILL_UNIVERSE;
      fill[i] = c.fill_;
    } else {
      type[i] = OMCGPU_FILL_LATTICE;
      fill[i] = c.fill_;
    }
    mofs.push_back(cmat.size());
    kofs.push_back(ckt.size());
    dindex[i] = c.distribcell_index_ >= 0 ? c.distribcell_index_ : -1;
    if (c.type_ != Fill::MATERIAL)
      for (int mp = 0; mp < n_maps; ++mp)
        off.push_back(mp < (int)c.offset_.size() ? c.offset_[mp] : -1);
    oofs.push_back(off.size());
  }
  F.m.n_cells = nc;
  KEEP(cell_id, id);
  KEEP(cell_type, type);
  KEEP(cell_universe, univ);
  KEEP(cell_fill, fill);
  KEEP(cell_simple, simple);
  KEEP(cell_region_ofs, rofs);
  KEEP(region_tokens, rtok);
  KEEP(cell_mat_ofs, mofs);
  KEEP(cell_material, cmat);
  KEEP(cell_sqrtkT_ofs, kofs);
  KEEP(cell_sqrtkT, ckt);
  KEEP(cell_density_ofs, dofs);
  KEEP(cell_density_mult, cden);
  KEEP(cell_translation, trans);
  KEEP(cell_rotated, rotated);
  KEEP(cell_rotation, rot);
  KEEP(cell_distribcell_index, dindex);
  if (n_maps == 0) {
    oofs.assign(nc + 1, 0);
    off.clear();
  }
  KEEP(cell_offset_ofs, oofs);
  KEEP(cell_offset, off);

  const int nu = model::universes.size();
  IV uofs {0}, ucells;
  for (int u = 0; u < nu; ++u) {
    if (model::universes[u]->geom_type() != GeometryType::CSG)
      unsupported("a DAGMC universe");
    for (int32_t c : model::universes[u]->cells_)
      ucells.push_back(c);
    uofs.push_back(ucells.size());
  }
  F.m.n_universes = nu;
  KEEP(univ_cell_ofs, uofs);
  KEEP(univ_cells, ucells);

  const int nl = model::lattices.size();
  const int nl1 = std::max(nl, 1);
  IV shape((size_t)nl1 * 3, 1), is3d(nl1, 0), outer(nl1, -1), lofs {0}, lun, loofs {0}, loff;
  DV pitch((size_t)nl1 * 3, 0.0), ll((size_t)nl1 * 3, 0.0);
  IV ltype(nl1, OMCGPU_LATTICE_RECT);
  for (int l = 0; l < nl; ++l) {
    const Lattice& base = *model::lattices[l];
    if (auto h = dynamic_cast<const HexLattice*>(model::lattices[l].get())) {
      // src/lattice.cpp:458-1009: the skewed square array as stored (C_NONE = -1 outside the hexagon)
      ltype[l] = h->orientation_ == HexLattice::Orientation::y ? OMCGPU_LATTICE_HEX_Y : OMCGPU_LATTICE_HEX_X;
      shape[3 * l + 0] = shape[3 * l + 1] = 2 * h->n_rings_ - 1;
      shape[3 * l + 2] = h->n_axial_;
      is3d[l] = h->is_3d_ ? 1 : 0;
      pitch[3 * l + 0] = h->pitch_[0];
      pitch[3 * l + 1] = h->is_3d_ ? h->pitch_[1] : 0.0;
      ll[3 * l + 0] = h->center_[0];
      ll[3 * l + 1] = h->center_[1];
      ll[3 * l + 2] = h->is_3d_ ? h->center_[2] : 0.0;
    } else if (auto r = dynamic_cast<const RectLattice*>(model::lattices[l].get())) {
      for (int k = 0; k < 3; ++k) {
        shape[3 * l + k] = r->n_cells_[k];
        pitch[3 * l + k] = r->pitch_[k];
        ll[3 * l + k] = r->lower_left_[k];
      }
      is3d[l] = r->is_3d_ ? 1 : 0;
      if (!r->is_3d_) { // the reader leaves the z entries of a 2-D lattice unset
        pitch[3 * l + 2] = 0.0;
        ll[3 * l + 2] = 0.0;
      }
    } else {
      unsupported(fmt::format("lattice {} (neither rectangular nor hexagonal)", base.id_));
    }
    outer[l] = base.outer_ == NO_OUTER_UNIVERSE ? -1 : base.outer_;
    for (int32_t u : base.universes_)
      lun.push_back(u);
    lofs.push_back(lun.size());
    if (n_maps > 0)
      for (size_t k = 0; k < (size_t)n_maps * base.universes_.size(); ++k)
        loff.push_back(k < base.offsets_.size() ? base.offsets_[k] : -1);
    loofs.push_back(loff.size());
  }
  while ((int)loofs.size() < nl1 + 1)
    loofs.push_back(loofs.back());
  F.m.n_lattices = nl;
  KEEP(lat_shape, shape);
  KEEP(lat_is3d, is3d);
  KEEP(lat_type, ltype);
  KEEP(lat_outer, outer);
  KEEP(lat_pitch, pitch);
  KEEP(lat_lower_left, ll);
  KEEP(lat_univ_ofs, lofs);
  KEEP(lat_universes, lun);
  KEEP(lat_offset_ofs, loofs);
  KEEP(lat_offsets, loff);
}

//------------------------------------------------------------------------------
// model::materials
void flatten_materials(Flat& F)
{
  const int nm = model::materials.size();
  IV id(nm), fis(nm), ofs {0}, nuc;
  DV dens;
  for (int i = 0; i < nm; ++i) {
    const Material& mat = *model::materials[i];
    id[i] = mat.id_;
    fis[i] = mat.fissionable_ ? 1 : 0;
    if (!mat.thermal_tables_.empty())
      unsupported(fmt::format("the S(a,b) table of material {}", mat.id_));
    for (size_t k = 0; k < mat.nuclide_.size(); ++k) {
      nuc.push_back(mat.nuclide_[k]);
      dens.push_back(mat.atom_density_(k));
    }
    ofs.push_back(nuc.size());
  }
  F.m.n_materials = nm;
  KEEP(mat_id, id);
  KEEP(mat_fissionable, fis);
  KEEP(mat_nuc_ofs, ofs);
  KEEP(mat_nuclide, nuc);
  KEEP(mat_atom_density, dens);
}

//------------------------------------------------------------------------------
// data::nuclides (continuous energy): grids, the derived xs_ table, the log hash, and the
// reactions / products / functions collision physics samples from
struct Functions {
  IV type, ofs {0};
  DV data;
  int add(const Function1D* f)
  {
    if (!f)
      return -1;
    if (auto p = dynamic_cast<const Polynomial*>(f)) {
      type.push_back(OMCGPU_FN_POLYNOMIAL);
      data.insert(data.end(), p->coef_.begin(), p->coef_.end());
    } else if (auto t = dynamic_cast<const Tabulated1D*>(f)) {
      if (t->n_regions_ > 1 || (t->n_regions_ == 1 && t->int_[0] != Interpolation::lin_lin))
        unsupported("a tabulated function that is not linear-linear in one region");
      type.push_back(OMCGPU_FN_TABULATED1D);
      data.insert(data.end(), t->x_.begin(), t->x_.end());
      data.insert(data.end(), t->y_.begin(), t->y_.end());
    } else {
      unsupported("a yield / parameter function that is neither Polynomial nor Tabulated1D");
    }
    ofs.push_back(data.size());
    return type.size() - 1;
  }
};

void flatten_nuclides(Flat& F)
{
  auto& m = F.m;
  const int nn = data::nuclides.size();
  m.n_log_bins = settings::n_log_bins;
  m.log_spacing = simulation::log_spacing;
  m.energy_min = data::energy_min[0];
  m.energy_max = data::energy_max[0];
  if (settings::temperature_multipole || settings::res_scat_on)
    unsupported("windowed multipole / resonance scattering");
  // A TABLE is one nuclide at one loaded temperature (kTs_[t], grid_[t], xs_[t], Reaction::xs_[t]);
  // every nuc_* / rx_* array has one entry per table, nuclide-major, ascending kT
  bool multi = false;
  for (int i = 0; i < nn; ++i)
    multi = multi || data::nuclides[i]->kTs_.size() != 1;
  if (multi && settings::temperature_method != TemperatureMethod::NEAREST)
    unsupported("temperature interpolation between the tables of a nuclide (only the nearest-temperature method)");
  DV awr, kT, energy, xs, rx_xs, rx_q, p_decay, p_par;
  IV tab_ofs {0}, fissionable, gidx, rx_ofs {0}, fis_rx, n_prec, tot_nu, rx_mt, rx_cm, rx_inel, rx_thr, rx_pofs {0}, p_em,
    p_yield, p_law, p_fn;
  LV gofs {0}, rx_xofs {0};
  Functions fn;
  for (int i = 0; i < nn; ++i) {
    const Nuclide& nuc = *data::nuclides[i];
    if (nuc.urr_present_ && settings::urr_ptables_on)
      unsupported(fmt::format("the unresolved-resonance probability tables of {}", nuc.name_));
    if (nuc.multipole_)
      unsupported("windowed multipole data");
    for (size_t it = 0; it < nuc.kTs_.size(); ++it) {
      awr.push_back(nuc.awr_);
      kT.push_back(nuc.kTs_[it]);
      fissionable.push_back(nuc.fissionable_ ? 1 : 0);
      const auto& grid = nuc.grid_[it];
      const auto& tab = nuc.xs_[it];
      const size_t n = grid.energy.size();
      energy.insert(energy.end(), grid.energy.begin(), grid.energy.end());
      for (size_t k = 0; k < n; ++k)
        for (int x = 0; x < 4; ++x) // total, absorption, fission, nu-fission (photon production dropped)
          xs.push_back(tab(k, x));
      gidx.insert(gidx.end(), grid.grid_index.begin(), grid.grid_index.end());
      if ((int)grid.grid_index.size() != settings::n_log_bins + 1)
        unsupported("a log-hash table whose size is not n_log_bins + 1");
      gofs.push_back(energy.size());
      n_prec.push_back(nuc.n_precursor_);
      tot_nu.push_back(fn.add(nuc.total_nu_.get()));
      fis_rx.push_back(-1);
      const int r0 = rx_mt.size();
      for (size_t r = 0; r < nuc.reactions_.size(); ++r) {
        const Reaction& rx = *nuc.reactions_[r];
        if (rx.redundant_)
          unsupported(fmt::format("the redundant reaction MT {} of {}", rx.mt_, nuc.name_));
        if (!nuc.fission_rx_.empty() && nuc.fission_rx_[0] == &rx)
          fis_rx.back() = r;
        rx_mt.push_back(rx.mt_);
        rx_cm.push_back(rx.scatter_in_cm_ ? 1 : 0);
        rx_inel.push_back(std::find(nuc.index_inelastic_scatter_.begin(), nuc.index_inelastic_scatter_.end(), (int)r) !=
                              nuc.index_inelastic_scatter_.end()
                            ? 1
                            : 0);
        rx_thr.push_back(rx.xs_[it].threshold);
        rx_xs.insert(rx_xs.end(), rx.xs_[it].value.begin(), rx.xs_[it].value.end());
        rx_xofs.push_back(rx_xs.size());
        rx_q.push_back(rx.q_value_);
        for (const auto& p : rx.products_) {
          if (p.particle_ != ParticleType::neutron)
            continue; // photon production: photon transport is refused above
          p_em.push_back(p.emission_mode_ == ReactionProduct::EmissionMode::prompt    ? 0
                         : p.emission_mode_ == ReactionProduct::EmissionMode::delayed ? 1
                                                                                      : 2);
          // decay_rate_ is read (and meaningful) for delayed emission only, src/reaction_product.cpp:43-52
          p_decay.push_back(p.emission_mode_ == ReactionProduct::EmissionMode::delayed ? p.decay_rate_ : 0.0);
          p_yield.push_back(fn.add(p.yield_.get()));
          if (p.distribution_.size() != 1)
            unsupported(fmt::format("a product of {} MT {} with several distributions", nuc.name_, rx.mt_));
          auto d = dynamic_cast<const UncorrelatedAngleEnergy*>(p.distribution_[0].get());
          if (!d)
            unsupported(fmt::format("the correlated angle-energy distribution of {} MT {}", nuc.name_, rx.mt_));
          if (!d->angle_.empty())
            unsupported(fmt::format("the anisotropic angular distribution of {} MT {}", nuc.name_, rx.mt_));
          const EnergyDistribution* e = d->energy_.get();
          int law = OMCGPU_ELAW_NONE, f0 = -1, f1 = -1;
          double par0 = 0.0, par1 = 0.0;
          if (!e) {
            if (rx.mt_ != ELASTIC)
              unsupported(fmt::format("{} MT {} without an energy distribution", nuc.name_, rx.mt_));
          } else if (auto l = dynamic_cast<const LevelInelastic*>(e)) {
            law = OMCGPU_ELAW_LEVEL; par0 = l->threshold_; par1 = l->mass_ratio_;
          } else if (auto l = dynamic_cast<const MaxwellEnergy*>(e)) {
            law = OMCGPU_ELAW_MAXWELL; par0 = l->u_; f0 = fn.add(&l->theta_);
          } else if (auto l = dynamic_cast<const Evaporation*>(e)) {
            law = OMCGPU_ELAW_EVAPORATION; par0 = l->u_; f0 = fn.add(&l->theta_);
          } else if (auto l = dynamic_cast<const WattEnergy*>(e)) {
            law = OMCGPU_ELAW_WATT; par0 = l->u_; f0 = fn.add(&l->a_); f1 = fn.add(&l->b_);
          } else {
            unsupported(fmt::format("the energy distribution of {} MT {} (continuous tabular / discrete photon)", nuc.name_, rx.mt_));
          }
          p_law.push_back(law);
          p_par.push_back(par0);
          p_par.push_back(par1);
          p_fn.push_back(f0);
          p_fn.push_back(f1);
        }
        rx_pofs.push_back(p_em.size());
      }
      if (rx_mt.size() == (size_t)r0 || rx_mt[r0] != ELASTIC)
        unsupported(fmt::format("{}: elastic scattering is not the first reaction", nuc.name_));
      rx_ofs.push_back(rx_mt.size());
    }
    tab_ofs.push_back(kT.size());
  }
  m.n_nuclides = kT.size();
  m.n_nuclide_temps = multi ? nn : 0;
  if (!multi)
    tab_ofs.assign(1, 0);
  KEEP(nuct_ofs, tab_ofs);
  KEEP(nuc_awr, awr);
  KEEP(nuc_kT, kT);
  KEEP(nuc_fissionable, fissionable);
  KEEP(nuc_grid_ofs, gofs);
  KEEP(nuc_energy, energy);
  KEEP(nuc_xs, xs);
  KEEP(nuc_grid_index, gidx);
  KEEP(nuc_rx_ofs, rx_ofs);
  KEEP(nuc_fission_rx, fis_rx);
  KEEP(nuc_n_precursor, n_prec);
  KEEP(nuc_total_nu_fn, tot_nu);
  m.n_reactions = rx_mt.size();
  KEEP(rx_mt, rx_mt);
  KEEP(rx_scatter_in_cm, rx_cm);
  KEEP(rx_inelastic, rx_inel);
  KEEP(rx_threshold, rx_thr);
  KEEP(rx_xs_ofs, rx_xofs);
  KEEP(rx_xs, rx_xs);
  KEEP(rx_q_value, rx_q);
  KEEP(rx_prod_ofs, rx_pofs);
  m.n_products = p_em.size();
  KEEP(prod_emission, p_em);
  KEEP(prod_decay_rate, p_decay);
  KEEP(prod_yield_fn, p_yield);
  KEEP(prod_elaw, p_law);
  KEEP(prod_par, p_par);
  KEEP(prod_fn, p_fn);
  m.n_functions = fn.type.size();
  KEEP(fn_type, fn.type);
  KEEP(fn_ofs, fn.ofs);
  KEEP(fn_data, fn.data);
}

//------------------------------------------------------------------------------
// data::mg: one macroscopic Mgxs per material (src/mgxs_interface.cpp:118-152)
void flatten_mg(Flat& F)
{
  auto& m = F.m;
  const int G = data::mg.num_energy_groups_;
  const int DG = data::mg.num_delayed_groups_;
  m.mg_n_groups = G;
  m.mg_n_delayed = DG;
  m.energy_min = data::mg.energy_bins_.back();
  m.energy_max = data::mg.energy_bins_.front();
  DV bin_avg(data::mg.energy_bin_avg_.begin(), data::mg.energy_bin_avg_.end());
  DV bins(data::mg.energy_bins_.begin(), data::mg.energy_bins_.end());
  KEEP(mg_energy_bin_avg, bin_avg);
  KEEP(mg_energy_bins, bins);
  const int nm = model::materials.size();
  IV n_temps(nm), n_pol(nm), n_azi(nm), iso(nm), fis(nm), blk_ofs {0}, ang_ofs {0};
  DV kts, total, absorption, nu_fission, prompt_nu_fission, fission, inverse_velocity, delayed_nu_fission, chi_prompt,
    chi_delayed, decay_rate;
  IV sc_format, sc_order, sc_gmin, sc_gmax, sc_gofs, sc_dofs, sc_mu_ofs;
  DV sc_scattxs, sc_energy, sc_mult, sc_maxval, sc_dist, sc_fmu, sc_mu, sc_dmu;
  for (int i = 0; i < nm; ++i) {
    const Mgxs& x = data::mg.macro_xs_[i];
    n_temps[i] = x.xs.size();
    n_pol[i] = x.n_pol;
    n_azi[i] = x.n_azi;
    iso[i] = x.is_isotropic ? 1 : 0;
    fis[i] = x.fissionable ? 1 : 0;
    const int n_ang = x.n_pol * x.n_azi;
    for (size_t t = 0; t < x.xs.size(); ++t) {
      const XsData& b = x.xs[t];
      kts.push_back(x.kTs(t));
      for (int a = 0; a < n_ang; ++a) {
        for (int g = 0; g < G; ++g) {
          total.push_back(b.total(a, g));
          absorption.push_back(b.absorption(a, g));
          inverse_velocity.push_back(b.inverse_velocity(a, g));
        }
        for (int d = 0; d < DG; ++d)
          decay_rate.push_back(b.decay_rate(a, d));
        if (x.fissionable) {
          for (int g = 0; g < G; ++g) {
            nu_fission.push_back(b.nu_fission(a, g));
            prompt_nu_fission.push_back(b.prompt_nu_fission(a, g));
            fission.push_back(b.fission(a, g));
          }
          for (int d = 0; d < DG; ++d)
            for (int g = 0; g < G; ++g)
              delayed_nu_fission.push_back(b.delayed_nu_fission(a, d, g));
          for (int gi = 0; gi < G; ++gi)
            for (int go = 0; go < G; ++go)
              chi_prompt.push_back(b.chi_prompt(a, gi, go));
          for (int d = 0; d < DG; ++d) // the reference keeps [gin][gout][dg]; the device reads [dg][gin][gout]
            for (int gi = 0; gi < G; ++gi)
              for (int go = 0; go < G; ++go)
                chi_delayed.push_back(b.chi_delayed(a, gi, go, d));
        } else {
          nu_fission.insert(nu_fission.end(), G, 0.0);
          prompt_nu_fission.insert(prompt_nu_fission.end(), G, 0.0);
          fission.insert(fission.end(), G, 0.0);
          delayed_nu_fission.insert(delayed_nu_fission.end(), (size_t)DG * G, 0.0);
          chi_prompt.insert(chi_prompt.end(), (size_t)G * G, 0.0);
          chi_delayed.insert(chi_delayed.end(), (size_t)DG * G * G, 0.0);
        }
        const ScattData* s = b.scatter[a].get();
        const ScattDataLegendre* leg = dynamic_cast<const ScattDataLegendre*>(s);
        const ScattDataHistogram* his = dynamic_cast<const ScattDataHistogram*>(s);
        const ScattDataTabular* tabu = dynamic_cast<const ScattDataTabular*>(s);
        const int order = s->dist[0][0].size();
        sc_format.push_back(leg ? OMCGPU_SCATT_LEGENDRE : his ? OMCGPU_SCATT_HISTOGRAM : OMCGPU_SCATT_TABULAR);
        sc_order.push_back(order);
        sc_mu_ofs.push_back(sc_mu.size());
        if (his) {
          sc_mu.insert(sc_mu.end(), his->mu.begin(), his->mu.end());
          sc_dmu.push_back(his->dmu);
        } else if (tabu) {
          sc_mu.insert(sc_mu.end(), tabu->mu.begin(), tabu->mu.end());
          sc_dmu.push_back(tabu->dmu);
        } else {
          sc_mu.push_back(0.0);
          sc_dmu.push_back(0.0);
        }
        for (int g = 0; g < G; ++g) {
          sc_gmin.push_back(s->gmin(g));
          sc_gmax.push_back(s->gmax(g));
          sc_scattxs.push_back(s->scattxs(g));
        }
        for (int gin = 0; gin < G; ++gin) {
          sc_gofs.push_back(sc_energy.size());
          sc_dofs.push_back(sc_dist.size());
          for (size_t k = 0; k < s->energy[gin].size(); ++k) {
            sc_energy.push_back(s->energy[gin][k]);
            sc_mult.push_back(s->mult[gin][k]);
            sc_maxval.push_back(leg ? leg->max_val[gin][k] : 0.0);
            sc_dist.insert(sc_dist.end(), s->dist[gin][k].begin(), s->dist[gin][k].end());
            if (his)
              sc_fmu.insert(sc_fmu.end(), his->fmu[gin][k].begin(), his->fmu[gin][k].end());
            else if (tabu)
              sc_fmu.insert(sc_fmu.end(), tabu->fmu[gin][k].begin(), tabu->fmu[gin][k].end());
            else
              sc_fmu.insert(sc_fmu.end(), order, 0.0);
          }
        }
      }
      ang_ofs.push_back(ang_ofs.back() + n_ang);
    }
    blk_ofs.push_back(kts.size());
  }
  KEEP(mgm_n_temps, n_temps);
  KEEP(mgm_n_pol, n_pol);
  KEEP(mgm_n_azi, n_azi);
  KEEP(mgm_isotropic, iso);
  KEEP(mgm_fissionable, fis);
  KEEP(mgm_blk_ofs, blk_ofs);
  KEEP(mgm_kTs, kts);
  KEEP(blk_ang_ofs, ang_ofs);
  KEEP(mg_total, total);
  KEEP(mg_absorption, absorption);
  KEEP(mg_nu_fission, nu_fission);
  KEEP(mg_prompt_nu_fission, prompt_nu_fission);
  KEEP(mg_fission, fission);
  KEEP(mg_inverse_velocity, inverse_velocity);
  KEEP(mg_delayed_nu_fission, delayed_nu_fission);
  KEEP(mg_chi_prompt, chi_prompt);
  KEEP(mg_chi_delayed, chi_delayed);
  KEEP(mg_decay_rate, decay_rate);
  KEEP(sc_format, sc_format);
  KEEP(sc_order, sc_order);
  KEEP(sc_gmin, sc_gmin);
  KEEP(sc_gmax, sc_gmax);
  KEEP(sc_gofs, sc_gofs);
  KEEP(sc_dofs, sc_dofs);
  KEEP(sc_scattxs, sc_scattxs);
  KEEP(sc_energy, sc_energy);
  KEEP(sc_mult, sc_mult);
  KEEP(sc_maxval, sc_maxval);
  KEEP(sc_dist, sc_dist);
  KEEP(sc_fmu, sc_fmu);
  KEEP(sc_mu_ofs, sc_mu_ofs);
  KEEP(sc_mu, sc_mu);
  KEEP(sc_dmu, sc_dmu);
}

//------------------------------------------------------------------------------
// model::meshes, model::tally_filters, model::tallies
void flatten_tallies(Flat& F)
{
  const int nmesh = model::meshes.size();
  const int nmesh1 = std::max(nmesh, 1);
  IV ndim(nmesh1, 0), shape((size_t)nmesh1 * 3, 1);
  DV ll((size_t)nmesh1 * 3, 0.0), width((size_t)nmesh1 * 3, 1.0);
  for (int i = 0; i < nmesh; ++i) {
    auto r = dynamic_cast<const RegularMesh*>(model::meshes[i].get());
    if (!r)
      continue; // refused below if a mesh filter names it
    ndim[i] = r->n_dimension_;
    for (int k = 0; k < r->n_dimension_; ++k) {
      shape[3 * i + k] = r->shape_[k];
      ll[3 * i + k] = r->lower_left_[k];
      width[3 * i + k] = r->width_[k];
    }
  }
  F.m.n_meshes = nmesh;
  KEEP(mesh_ndim, ndim);
  KEEP(mesh_shape, shape);
  KEEP(mesh_lower_left, ll);
  KEEP(mesh_width, width);

  const int nf = model::tally_filters.size();
  const int nf1 = std::max(nf, 1);
  IV ftype(nf1, 0), fnb(nf1, 0), fmesh(nf1, -1), fmg(nf1, 0), fofs {0}, fi;
  DV fd;
  std::vector<bool> usable(nf1, true);
  auto id_map = [&](const std::unordered_map<int32_t, int>& map, int n) {
    IV mp(n, -1);
    for (auto& kv : map)
      mp[kv.first] = kv.second;
    fi.insert(fi.end(), mp.begin(), mp.end());
    fd.insert(fd.end(), n, 0.0);
  };
  for (int i = 0; i < nf; ++i) {
    const Filter* f = model::tally_filters[i].get();
    fnb[i] = f->n_bins();
    if (auto p = dynamic_cast<const MeshFilter*>(f)) {
      ftype[i] = OMCGPU_FILTER_MESH;
      fmesh[i] = p->mesh_;
      usable[i] = f->type() == FilterType::MESH && !p->translated_ &&
                  dynamic_cast<const RegularMesh*>(model::meshes[p->mesh_].get()) != nullptr;
    } else if (auto p = dynamic_cast<const EnergyFilter*>(f)) {
      // EnergyoutFilter derives from EnergyFilter: same bins, matched on the post-collision energy
      ftype[i] = f->type() == FilterType::ENERGY_OUT ? OMCGPU_FILTER_ENERGYOUT : OMCGPU_FILTER_ENERGY;
      usable[i] = f->type() == FilterType::ENERGY || f->type() == FilterType::ENERGY_OUT;
      fmg[i] = p->matches_transport_groups_ ? 1 : 0;
      for (double e : p->bins_) {
        fd.push_back(e);
        fi.push_back(0);
      }
    } else if (auto p = dynamic_cast<const DistribcellFilter*>(f)) {
      ftype[i] = OMCGPU_FILTER_DISTRIBCELL;
      fi.push_back(p->cell_);
      fd.push_back(0.0);
    } else if (auto p = dynamic_cast<const CellFilter*>(f)) {
      ftype[i] = OMCGPU_FILTER_CELL;
      usable[i] = f->type() == FilterType::CELL;
      id_map(p->map_, model::cells.size());
    } else if (auto p = dynamic_cast<const MaterialFilter*>(f)) {
      ftype[i] = OMCGPU_FILTER_MATERIAL;
      usable[i] = f->type() == FilterType::MATERIAL;
      id_map(p->map_, model::materials.size());
    } else if (auto p = dynamic_cast<const UniverseFilter*>(f)) {
      ftype[i] = OMCGPU_FILTER_UNIVERSE;
      id_map(p->map_, model::universes.size());
    } else {
      usable[i] = false;
    }
    fofs.push_back(fi.size());
  }
  F.m.n_filters = nf;
  KEEP(filter_type, ftype);
  KEEP(filter_n_bins, fnb);
  KEEP(filter_mesh, fmesh);
  KEEP(filter_matches_groups, fmg);
  KEEP(filter_data_ofs, fofs);
  KEEP(filter_idata, fi);
  KEEP(filter_ddata, fd);

  const int nt = model::tallies.size();
  const int nt1 = std::max(nt, 1);
  IV est(nt1, OMCGPU_ESTIMATOR_TRACKLENGTH), tnf(nt1, 0), tf((size_t)nt1 * OMCGPU_MAX_FILTERS, 0),
    tst((size_t)nt1 * OMCGPU_MAX_FILTERS, 0), tns(nt1, 0), tsc((size_t)nt1 * OMCGPU_MAX_SCORES, 0);
  LV tofs {0};
  for (int t = 0; t < nt; ++t) {
    const Tally& ty = *model::tallies[t];
    const auto why = [&](const char* what) { unsupported(fmt::format("tally {}: {}", ty.id_, what)); };
    if (ty.nuclides_.size() != 1 || ty.nuclides_[0] != -1)
      why("nuclide bins");
    if (ty.deriv_ != C_NONE)
      why("a derivative");
    if ((int)ty.filters_.size() > OMCGPU_MAX_FILTERS || (int)ty.scores_.size() > OMCGPU_MAX_SCORES)
      why("more filters or scores than the device tally holds");
    est[t] = ty.estimator_ == TallyEstimator::TRACKLENGTH ? OMCGPU_ESTIMATOR_TRACKLENGTH
             : ty.estimator_ == TallyEstimator::COLLISION ? OMCGPU_ESTIMATOR_COLLISION
                                                          : OMCGPU_ESTIMATOR_ANALOG;
    tnf[t] = ty.filters_.size();
    for (size_t k = 0; k < ty.filters_.size(); ++k) {
      if (!usable[ty.filters_[k]])
        why("a filter type the device does not evaluate");
      tf[(size_t)t * OMCGPU_MAX_FILTERS + k] = ty.filters_[k];
      tst[(size_t)t * OMCGPU_MAX_FILTERS + k] = ty.strides_[k];
    }
    tns[t] = ty.scores_.size();
    for (size_t k = 0; k < ty.scores_.size(); ++k) {
      int s;
      switch (ty.scores_[k]) {
      case SCORE_FLUX: s = OMCGPU_SCORE_FLUX; break;
      case SCORE_TOTAL: s = OMCGPU_SCORE_TOTAL; break;
      case SCORE_ABSORPTION: s = OMCGPU_SCORE_ABSORPTION; break;
      case SCORE_FISSION: s = OMCGPU_SCORE_FISSION; break;
      case SCORE_NU_FISSION: s = OMCGPU_SCORE_NU_FISSION; break;
      case SCORE_SCATTER: s = OMCGPU_SCORE_SCATTER; break;
      default: why("a score other than flux / total / absorption / fission / nu-fission / scatter"); s = 0;
      }
      tsc[(size_t)t * OMCGPU_MAX_SCORES + k] = s;
    }
    tofs.push_back(tofs.back() + (int64_t)ty.n_filter_bins_ * ty.scores_.size());
  }
  F.m.n_tallies = nt;
  KEEP(tally_estimator, est);
  KEEP(tally_n_filters, tnf);
  KEEP(tally_filters, tf);
  KEEP(tally_strides, tst);
  KEEP(tally_n_scores, tns);
  KEEP(tally_scores, tsc);
  KEEP(tally_result_ofs, tofs);
}

} // namespace

const omcgpu_model* flatten_model()
{
  auto F = std::make_unique<Flat>();
  flatten_settings(*F);
  flatten_surfaces(*F);
  flatten_cells(*F);
  flatten_materials(*F);
  if (settings::run_CE) {
    flatten_nuclides(*F);
  } else {
    F->m.n_log_bins = 0;
    F->m.log_spacing = 1.0;
    flatten_mg(*F);
  }
  // tallies need their strides (set in openmc_simulation_init right after the hook's site)
  for (auto& t : model::tallies)
    t->set_strides();
  flatten_tallies(*F);
  g_flat = std::move(F);
  return &g_flat->m;
}

//==============================================================================
// the device library, loaded at run time
//==============================================================================
namespace {

struct Api {
  void* handle = nullptr;
  int (*init)(const omcgpu_model*, int, int64_t) = nullptr;
  int (*finalize)() = nullptr;
  const char* (*last_error)() = nullptr;
  int (*next_generation)(const omcgpu_gen_in*, omcgpu_site*, int64_t, omcgpu_gen_out*) = nullptr;
  int (*tally_results)(int32_t, double*, int64_t) = nullptr;
  int (*host_register)(void*, int64_t) = nullptr;
  int (*host_unregister)(void*) = nullptr;
  int64_t (*lost_particles)(double*, int64_t) = nullptr;
  int (*needs_extended_build)(const omcgpu_model*) = nullptr;
  int (*count_fission_sites)(int32_t, const int32_t*, const double*, const double*, double*, int64_t, int32_t*) = nullptr;
  int (*update)(const omcgpu_model*, uint32_t) = nullptr;
} api;
bool g_ready = false;
std::vector<double> g_values;

template<class Fn>
void bind(Fn& fn, const char* name)
{
  fn = reinterpret_cast<Fn>(dlsym(api.handle, name));
  if (!fn)
    fatal_error(fmt::format("OPENMC_GPU: {} lacks the symbol {}", "libomcgpu", name));
}

bool g_extended = false; // the model needs libomcgpu_ext.so (omcgpu_needs_extended_build)

void load_library()
{
  if (api.handle)
    return;
  std::string path;
  if (const char* env = std::getenv("OMCGPU_LIB")) {
    path = env;
  } else {
    Dl_info info;
    if (dladdr(reinterpret_cast<void*>(&load_library), &info) && info.dli_fname) {
      path = info.dli_fname;
      path = path.substr(0, path.find_last_of('/') + 1) + (g_extended ? "libomcgpu_ext.so" : "libomcgpu.so");
    } else {
      path = g_extended ? "libomcgpu_ext.so" : "libomcgpu.so";
    }
  }
  api.handle = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!api.handle)
    fatal_error(fmt::format("OPENMC_GPU: cannot load the device library {}: {}", path, dlerror()));
  bind(api.init, "omcgpu_init");
  bind(api.finalize, "omcgpu_finalize");
  bind(api.last_error, "omcgpu_last_error");
  bind(api.next_generation, "omcgpu_next_generation");
  bind(api.tally_results, "omcgpu_tally_results");
  bind(api.host_register, "omcgpu_host_register");
  bind(api.host_unregister, "omcgpu_host_unregister");
  bind(api.lost_particles, "omcgpu_lost_particles");
  bind(api.needs_extended_build, "omcgpu_needs_extended_build");
  bind(api.count_fission_sites, "omcgpu_count_fission_sites");
  bind(api.update, "omcgpu_update");
}

void check(int rc, const char* what)
{
  if (rc != 0)
    fatal_error(fmt::format("OPENMC_GPU: {} failed ({}): {}", what, rc, api.last_error ? api.last_error() : "?"));
}

} // namespace

bool enabled()
{
  static const bool on = [] {
    const char* env = std::getenv("OPENMC_GPU");
    return env && env[0] && std::strcmp(env, "0") != 0;
  }();
  return on;
}

void simulation_init()
{
  load_library();
  const omcgpu_model* m;
  try {
    m = flatten_model();
  } catch (const std::runtime_error& e) {
    fatal_error(e.what());
  }
  // which build of the device library (include/omcgpu.h, omcgpu_needs_extended_build): the default
  // one answers; a model with hexagonal lattices or secondary neutrons gets libomcgpu_ext.so
  if (!std::getenv("OMCGPU_LIB") && api.needs_extended_build(m)) {
    dlclose(api.handle);
    api = Api {};
    g_extended = true;
    load_library();
  }
  int device = 0;
  if (const char* env = std::getenv("OPENMC_GPU_DEVICE"))
    device = std::atoi(env);
  check(api.init(m, device, std::max<int64_t>(simulation::work_per_rank, 1)), "omcgpu_init");
  // the banks were allocated by allocate_banks() just above the hook: page-lock them so that the
  // per-generation copies are DMA transfers straight from / into them
  api.host_register(simulation::source_bank.data(), simulation::source_bank.size() * sizeof(SourceSite));
  g_signature = structure_signature();
  g_ready = true;
}

// openmc_cell_set_temperature / openmc_material_set_density(ies) / openmc_cell_set_density
// (include/openmc/capi.h) change the HOST model between batches -- the whole point of openmc.lib in
// coupled calculations.  Before every generation the mutable arrays are rebuilt from model:: exactly
// as the flattener built them (a few thousand doubles) and compared with what the device has:
// changed temperatures / atom densities go over with omcgpu_update, anything that would change an
// array's shape or a density multiplier (folded into the device materials at omcgpu_init) stops the
// run by name.  No silent stale data.
// what omcgpu_init turned into device tables and omcgpu_update cannot change: cell fills / materials,
// the tallies with their filters, scores and estimators.  openmc.lib can change all of them between
// batches (openmc_cell_set_fill, openmc_extend_tallies, openmc_tally_set_scores ...); the device
// would go on with the old ones, so a changed signature stops the run by name.
static uint64_t structure_signature()
{
  uint64_t h = 1469598103934665603ull; // FNV-1a
  auto mix = [&](int64_t v) {
    for (int k = 0; k < 8; ++k) {
      h ^= (uint64_t)(v >> (8 * k)) & 0xff;
      h *= 1099511628211ull;
    }
  };
  mix(model::cells.size());
  for (const auto& cp : model::cells) {
    mix((int64_t)cp->type_);
    mix(cp->fill_);
    mix(cp->material_.size());
    for (int32_t m : cp->material_)
      mix(m);
  }
  mix(model::tally_filters.size());
  for (const auto& f : model::tally_filters)
    mix(f->n_bins());
  mix(model::tallies.size());
  for (const auto& tp : model::tallies) {
    mix((int64_t)tp->estimator_);
    mix(tp->n_filter_bins_);
    for (auto f : tp->filters())
      mix(f);
    for (int sc : tp->scores_)
      mix(sc);
    for (int nu : tp->nuclides_)
      mix(nu);
  }
  return h;
}
static uint64_t g_signature = 0;

static void refresh_mutable_properties()
{
  if (structure_signature() != g_signature)
    fatal_error("OPENMC_GPU: cell fills / materials or the tallies (filters, scores, estimators) changed after "
                "initialisation: not supported by the device path (it holds the tables omcgpu_init built)");
  Flat& F = *g_flat;
  std::vector<double> ckt, dens, cden;
  for (const auto& cp : model::cells) {
    const Cell& c = *cp;
    if (c.type_ == Fill::MATERIAL)
      for (double x : c.sqrtkT_)
        ckt.push_back(x);
    for (double d : c.density_mult_)
      cden.push_back(d);
  }
  for (const auto& mp : model::materials)
    for (size_t k = 0; k < mp->nuclide_.size(); ++k)
      dens.push_back(mp->atom_density_(k));
  auto differs = [&](const char* name, const double* have, const std::vector<double>& now) {
    if ((int64_t)now.size() != F.length[name])
      fatal_error(fmt::format("OPENMC_GPU: the number of entries of {} changed after initialisation ({} -> {}): "
                              "not supported by the device path", name, F.length[name], now.size()));
    return std::memcmp(have, now.data(), now.size() * sizeof(double)) != 0;
  };
  if (differs("cell_density_mult", F.m.cell_density_mult, cden))
    fatal_error("OPENMC_GPU: cell densities changed after initialisation: not supported by the device path "
                "(they are folded into the device materials at omcgpu_init)");
  uint32_t what = 0;
  if (differs("cell_sqrtkT", F.m.cell_sqrtkT, ckt)) {
    std::memcpy(const_cast<double*>(F.m.cell_sqrtkT), ckt.data(), ckt.size() * sizeof(double));
    what |= OMCGPU_UPDATE_CELL_TEMPERATURES;
  }
  if (differs("mat_atom_density", F.m.mat_atom_density, dens)) {
    std::memcpy(const_cast<double*>(F.m.mat_atom_density), dens.data(), dens.size() * sizeof(double));
    what |= OMCGPU_UPDATE_MATERIAL_DENSITIES;
  }
  if (what)
    check(api.update(&F.m, what), "omcgpu_update");
}

// Particle::mark_as_lost (src/particle.cpp:762-785) for the particles the device lost in this
// generation: the warning per particle (the device keeps the state of the first
// OMCGPU_MAX_LOST_RECORDS), the count, and the reference's abort rule.  No particle restart files:
// the history was not run on the host.
static void report_lost_particles(const omcgpu_gen_out& out)
{
  double rec[OMCGPU_MAX_LOST_RECORDS][OMCGPU_LOST_RECORD];
  const int64_t n = api.lost_particles(&rec[0][0], OMCGPU_MAX_LOST_RECORDS);
  for (int64_t k = 0; k < n; ++k)
    warning(fmt::format("Particle {} was lost on the device (no cell found) at r = ({}, {}, {}), "
                        "u = ({}, {}, {}), E = {}",
      (int64_t)rec[k][0], rec[k][1], rec[k][2], rec[k][3], rec[k][4], rec[k][5], rec[k][6], rec[k][7]));
  if (out.n_lost > n)
    warning(fmt::format("{} more particles were lost in this generation", out.n_lost - n));
  simulation::n_lost_particles += out.n_lost;
  const auto n_run = simulation::current_batch * settings::gen_per_batch * simulation::work_per_rank;
  if (simulation::n_lost_particles >= settings::max_lost_particles &&
      simulation::n_lost_particles >= settings::rel_max_lost_particles * n_run)
    fatal_error("Maximum number of lost particles has been reached.");
}

// the device's per-kernel times go where print_runtime reads them (src/output.cpp:439-446, shown
// for event-based runs): the crossing / collision kernels carry the lookup and the flight of the
// following event (sweep mode 2), so "XS lookups" is the stand-alone lookup kernel only and
// "Advancing" stays zero; tally scoring on the device is booked under "Particle death", the event
// that hands the history's scores over
static void add_seconds(Timer& t, double ms)
{
  t.elapsed_ += 1e-3 * ms;
}

void transport_generation()
{
  if (!g_ready)
    fatal_error("OPENMC_GPU: transport_generation before simulation_init");
  refresh_mutable_properties();
  omcgpu_gen_in in {};
  in.n_particles = simulation::work_per_rank;
  in.id_offset = simulation::work_index[mpi::rank];
  // Particle::from_source / initialize_history, src/simulation.cpp:613-616
  in.particle_seed_base = (simulation::total_gen + overall_generation() - 1) * settings::n_particles;
  in.master_seed = openmc_get_seed();
  in.stride = openmc_get_stride();
  in.keff = simulation::keff;
  in.tallies_active = model::active_tallies.empty() ? 0 : 1;
  in.source_bank_clean = 0; // the host bank is authoritative (ownership)
  omcgpu_gen_out out {};
  // event loop, fission-bank sort (src/bank.cpp:76-134) and the sampling half of synchronize_bank
  // (src/eigenvalue.cpp:152-178, same seed: total_gen + overall_generation()) on the device; the NEXT
  // source bank comes back into simulation::source_bank
  check(api.next_generation(&in, reinterpret_cast<omcgpu_site*>(simulation::source_bank.data()),
          simulation::total_gen + overall_generation(), &out),
    "omcgpu_next_generation");

  // what event_death() leaves (src/particle.cpp:475-512)
  global_tally_collision += out.global_tally[OMCGPU_K_COLLISION];
  global_tally_absorption += out.global_tally[OMCGPU_K_ABSORPTION];
  global_tally_tracklength += out.global_tally[OMCGPU_K_TRACKLENGTH];
  global_tally_leakage += out.global_tally[OMCGPU_LEAKAGE];
  simulation::total_weight += out.total_weight;
  if (out.n_lost > 0)
    report_lost_particles(out);
  add_seconds(simulation::time_event_calculate_xs, out.ms_xs);
  add_seconds(simulation::time_event_surface_crossing, out.ms_surface);
  add_seconds(simulation::time_event_collision, out.ms_collide);
  add_seconds(simulation::time_event_death, out.ms_score);
  if (out.n_bank_full > 0)
    warning("The shared fission bank is full. Additional fission sites created in this generation will not be banked.");
  if (out.n_fission == 0)
    fatal_error("No fission sites banked on MPI rank " + std::to_string(mpi::rank)); // src/eigenvalue.cpp:121-124

  // tally VALUE arrays -> results_(:, :, VALUE), the hand-off before Tally::accumulate
  if (in.tallies_active) {
    for (size_t t = 0; t < model::tallies.size(); ++t) {
      Tally& ty = *model::tallies[t];
      const int64_t n = (int64_t)ty.n_filter_bins_ * ty.scores_.size();
      g_values.assign(n, 0.0);
      check(api.tally_results(t, g_values.data(), n), "omcgpu_tally_results");
      if (!ty.active_)
        continue;
      const size_t ns = ty.scores_.size();
      for (int64_t k = 0; k < n; ++k)
        ty.results_(k / ns, k % ns, TallyResult::VALUE) += g_values[k];
    }
  }
}

bool bank_synchronized()
{
  return enabled() && g_ready;
}

// shannon_entropy, src/eigenvalue.cpp:579-607: the reference counts the sites of
// simulation::fission_bank on the entropy mesh (Mesh::count_sites, src/mesh.cpp:941-997); in the
// drop-in that bank never leaves the device, so the device counts (unit weights: the sums are exact
// whatever the order) and the rest is the reference's own arithmetic
void shannon_entropy()
{
  const RegularMesh* mesh = simulation::entropy_mesh;
  const int nd = mesh->n_dimension_;
  int32_t shape[3] = {1, 1, 1};
  double ll[3] = {0, 0, 0}, w[3] = {1, 1, 1};
  int64_t n_bins = 1;
  for (int k = 0; k < nd; ++k) {
    shape[k] = mesh->shape_[k];
    ll[k] = mesh->lower_left_[k];
    w[k] = mesh->width_[k];
    n_bins *= shape[k];
  }
  std::vector<double> p(n_bins, 0.0);
  int32_t sites_outside = 0;
  check(api.count_fission_sites(nd, shape, ll, w, p.data(), n_bins, &sites_outside), "omcgpu_count_fission_sites");
  if (sites_outside && mpi::master)
    warning("Fission source site(s) outside of entropy box.");
  if (mpi::master) {
    double total = 0.0; // xt::sum(p)
    for (double x : p)
      total += x;
    double H = 0.0;
    for (double x : p) {
      const double p_i = x / total;
      if (p_i > 0.0)
        H -= p_i * std::log2(p_i);
    }
    simulation::entropy.push_back(H);
  }
}

void simulation_finalize()
{
  if (!g_ready)
    return;
  api.host_unregister(simulation::source_bank.data());
  api.finalize();
  g_ready = false;
  g_flat.reset();
}

} // namespace gpu
} // namespace openmc

//==============================================================================
// C entry points (for openmc.lib-style callers and the flattener tests)
//==============================================================================
extern "C" const omcgpu_model* openmc_gpu_flatten_model(void)
{
  try {
    return openmc::gpu::flatten_model();
  } catch (const std::runtime_error& e) {
    openmc::gpu::g_error = e.what();
    openmc::set_errmsg(e.what());
    return nullptr;
  }
}

extern "C" int64_t openmc_gpu_model_array_length(const char* field)
{
  if (!openmc::gpu::g_flat)
    return -1;
  auto it = openmc::gpu::g_flat->length.find(field);
  return it == openmc::gpu::g_flat->length.end() ? -1 : it->second;
}
