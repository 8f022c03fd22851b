/*
 * omcgpu.h -- C ABI of the B200-native event-based transport engine.
 *
 * This is the drop-in boundary described in SURVEY.md section 8(b).  OpenMC has
 * no plugin interface for transport; the path sits behind ONE internal call
 * site:
 *
 *   openmc_next_batch -> transport_event_based()      src/simulation.cpp:255-259
 *   allocation in openmc_simulation_init               src/simulation.cpp:101-105
 *   release in free_memory                             src/finalize.cpp:57-59
 *
 * libopenmc (host, C++) keeps owning XML/HDF5 loading, batch statistics and
 * statepoints.  After openmc_simulation_init it flattens its read-only model
 * (model::surfaces/cells/universes/lattices/materials, data::nuclides or
 * data::mg, model::tallies/filters/meshes) into the POD arrays of
 * `omcgpu_model` and hands them to omcgpu_init().  Every generation it calls
 * omcgpu_transport_generation() instead of transport_event_based().
 *
 * Rules of this header: extern "C", plain pointers and sizes, fixed-width
 * integer types, no C++ or torch types.  All arrays are caller-owned host
 * memory unless a function name says `_device`.  Every function returns 0 on
 * success or a negative OMCGPU_E_* code; the message is available from
 * omcgpu_last_error() (same convention as openmc_err_msg, src/error.cpp:20-33).
 *
 * The struct field lines are kept in a rigid `type name;` form because the
 * Python binding (openmc_b200/_abi.py) generates its ctypes mirror by parsing
 * this file -- the header is the single source of truth.
 */
#ifndef OMCGPU_H
#define OMCGPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* error codes (negative like OPENMC_E_*, src/error.cpp:20-30)                */
/* ------------------------------------------------------------------------- */
#define OMCGPU_E_UNASSIGNED   (-1)
#define OMCGPU_E_ALLOCATE     (-2)
#define OMCGPU_E_OUT_OF_BOUNDS (-3)
#define OMCGPU_E_INVALID_ARGUMENT (-5)
#define OMCGPU_E_GEOMETRY     (-10)
#define OMCGPU_E_DATA         (-11)
#define OMCGPU_E_PHYSICS      (-12)
#define OMCGPU_E_UNSUPPORTED  (-20)  /* feature the device path refuses (no CPU fallback) */
#define OMCGPU_E_CUDA         (-30)
#define OMCGPU_E_NCCL         (-31)

/* ------------------------------------------------------------------------- */
/* enumerations (values are part of the ABI)                                   */
/* ------------------------------------------------------------------------- */

/* surface kinds: src/surface.cpp:1172-1385 (type strings of the XML reader) */
#define OMCGPU_SURF_XPLANE    0   /* coef: x0                  */
#define OMCGPU_SURF_YPLANE    1   /* coef: y0                  */
#define OMCGPU_SURF_ZPLANE    2   /* coef: z0                  */
#define OMCGPU_SURF_PLANE     3   /* coef: A B C D             */
#define OMCGPU_SURF_XCYLINDER 4   /* coef: y0 z0 r             */
#define OMCGPU_SURF_YCYLINDER 5   /* coef: x0 z0 r             */
#define OMCGPU_SURF_ZCYLINDER 6   /* coef: x0 y0 r             */
#define OMCGPU_SURF_SPHERE    7   /* coef: x0 y0 z0 r          */
#define OMCGPU_SURF_XCONE     8   /* coef: x0 y0 z0 r^2  (src/surface.cpp:688-742) */
#define OMCGPU_SURF_YCONE     9
#define OMCGPU_SURF_ZCONE     10
#define OMCGPU_SURF_QUADRIC   11  /* coef: row index into quad_coef = A B C D E F G H J K  (src/surface.cpp:855-949) */
#define OMCGPU_SURF_NCOEF     4   /* doubles stored per surface */

/* boundary conditions: src/boundary_condition.cpp:18-62 */
#define OMCGPU_BC_NONE        0   /* transmission */
#define OMCGPU_BC_VACUUM      1
#define OMCGPU_BC_REFLECTIVE  2
#define OMCGPU_BC_PERIODIC    3   /* translational between two parallel planes (src/boundary_condition.cpp:66-155), or rotational
                                   about x / y / z between two planes through the origin (:161-287): surf_periodic_axis */
#define OMCGPU_BC_WHITE       4   /* cosine-distributed diffuse reflection (WhiteBC :53-62, Surface::diffuse_reflect src/surface.cpp:142-161) */

/* cell fill: include/openmc/cell.h Fill enum */
#define OMCGPU_FILL_MATERIAL  0
#define OMCGPU_FILL_UNIVERSE  1
#define OMCGPU_FILL_LATTICE   2

/* region operator tokens: include/openmc/cell.h:31-35 */
#define OMCGPU_OP_LEFT_PAREN   2147483647
#define OMCGPU_OP_RIGHT_PAREN  2147483646
#define OMCGPU_OP_COMPLEMENT   2147483645
#define OMCGPU_OP_INTERSECTION 2147483644
#define OMCGPU_OP_UNION        2147483643

/* MG scattering representation: include/openmc/constants.h AngleDistributionType */
#define OMCGPU_SCATT_LEGENDRE  0
#define OMCGPU_SCATT_HISTOGRAM 1
#define OMCGPU_SCATT_TABULAR   2

/* CE secondary-energy laws supported on the device (src/distribution_energy.cpp) */
#define OMCGPU_ELAW_NONE        0   /* elastic: kinematics only */
#define OMCGPU_ELAW_LEVEL       1   /* LevelInelastic  :42-55   par: threshold, mass_ratio */
#define OMCGPU_ELAW_MAXWELL     2   /* MaxwellEnergy   :274-298 par: u ; fn0 = theta(E)    */
#define OMCGPU_ELAW_EVAPORATION 3   /* Evaporation     :301-327 par: u ; fn0 = theta(E)    */
#define OMCGPU_ELAW_WATT        4   /* WattEnergy      :333-361 par: u ; fn0 = a(E), fn1 = b(E) */

/* 1-D functions (src/endf.cpp:125-226) */
#define OMCGPU_FN_POLYNOMIAL   0   /* data = coefficients, increasing powers  */
#define OMCGPU_FN_TABULATED1D  1   /* data = x[n] then y[n]; lin-lin only     */

/* tally filters / scores / estimators (subset; include/openmc/constants.h:295-334) */
#define OMCGPU_FILTER_MESH      0
#define OMCGPU_FILTER_ENERGY    1
#define OMCGPU_FILTER_CELL      2
#define OMCGPU_FILTER_MATERIAL  3
#define OMCGPU_FILTER_UNIVERSE  4
#define OMCGPU_FILTER_DISTRIBCELL 5 /* filter_idata = { cell index }; bins = instances of that cell
                                     (src/tallies/filter_distribcell.cpp:40-63) */
#define OMCGPU_FILTER_ENERGYOUT 6 /* EnergyoutFilter (src/tallies/filter_energy.cpp:100-125): the energy / group AFTER the
                                     collision; analog tallies only; filter_ddata = bin edges as for ENERGY */
/* LatticeType + HexLattice::Orientation, include/openmc/lattice.h */
#define OMCGPU_LATTICE_RECT  0
#define OMCGPU_LATTICE_HEX_Y 1
#define OMCGPU_LATTICE_HEX_X 2
/* TallyEstimator, include/openmc/constants.h:299 */
#define OMCGPU_ESTIMATOR_ANALOG      0 /* scored after the collision from the collision records */
#define OMCGPU_ESTIMATOR_TRACKLENGTH 1
#define OMCGPU_ESTIMATOR_COLLISION   2
#define OMCGPU_SCORE_FLUX        0
#define OMCGPU_SCORE_TOTAL       1
#define OMCGPU_SCORE_ABSORPTION  2
#define OMCGPU_SCORE_FISSION     3
#define OMCGPU_SCORE_NU_FISSION  4
#define OMCGPU_SCORE_SCATTER     5
#define OMCGPU_MAX_FILTERS       3
#define OMCGPU_MAX_SCORES        6

/* global tallies, order of include/openmc/constants.h:334 (GlobalTally) */
#define OMCGPU_K_COLLISION   0
#define OMCGPU_K_ABSORPTION  1
#define OMCGPU_K_TRACKLENGTH 2
#define OMCGPU_LEAKAGE       3

#define OMCGPU_FINE_PER_LOG_BIN 16 /* fine-hash bins per bin of the reference's log grid */
#define OMCGPU_MAX_COORD  6      /* deepest universe nesting kept per particle */

/* ------------------------------------------------------------------------- */
/* SourceSite -- bit-compatible with include/openmc/particle_data.h:43-57 and */
/* the ctypes mirror openmc/lib/core.py:23-34 (104 bytes)                     */
/* ------------------------------------------------------------------------- */
typedef struct omcgpu_site {
  double r[3];
  double u[3];
  double E;              /* eV (CE) or group index stored as a double (MG) */
  double time;
  double wgt;
  int32_t delayed_group;
  int32_t surf_id;
  int32_t particle;      /* 0 = neutron */
  int32_t parent_nuclide;
  int64_t parent_id;
  int64_t progeny_id;
} omcgpu_site;

/* ------------------------------------------------------------------------- */
/* The flattened read-only model.                      */
/* Index conventions: everything is a 0-based *index* (not a user id);        */
/* `*_ofs` arrays have n+1 entries (CSR style).                               */
/* ------------------------------------------------------------------------- */
typedef struct omcgpu_model {
  /* ---- run constants: settings:: / simulation:: (src/settings.cpp:40-160) -- */
  int32_t run_CE;                 /* settings::run_CE                                  */
  int32_t n_coord_levels;         /* model::n_coord_levels                             */
  int32_t root_universe;          /* model::root_universe                              */
  int32_t survival_biasing;       /* settings::survival_biasing                        */
  int32_t max_particle_events;    /* settings::max_particle_events (1000000)           */
  int32_t create_delayed_neutrons;/* settings::create_delayed_neutrons                 */
  double weight_cutoff;           /* settings::weight_cutoff  (0.25)                   */
  double weight_survive;          /* settings::weight_survive (1.0)                    */
  double energy_cutoff;           /* settings::energy_cutoff[neutron]                  */
  double energy_min;              /* data::energy_min[neutron]                         */
  double energy_max;              /* data::energy_max[neutron]                         */
  double free_gas_threshold;      /* settings::free_gas_threshold (400)                */

  /* ---- surfaces: model::surfaces ------------------------------------------ */
  int32_t n_surfaces;
  const int32_t* surf_type;       /* [n_surfaces] OMCGPU_SURF_*                        */
  const int32_t* surf_bc;         /* [n_surfaces] OMCGPU_BC_*                          */
  const int32_t* surf_periodic;   /* [n_surfaces] partner surface of a periodic BC, else -1 */
  const double* surf_periodic_shift; /* [n_surfaces][3] translational: added to r when this periodic surface is struck;
                                        rotational: { cos(theta), sin(theta), 0 } of the rotation applied to r and u */
  const int32_t* surf_periodic_axis; /* [n_surfaces] -1 translational, else 0 / 1 / 2 = rotation about x / y / z */
  const int32_t* surf_id;         /* [n_surfaces] user ids (messages only)             */
  const double* surf_coef;        /* [n_surfaces][OMCGPU_SURF_NCOEF]                   */
  int32_t n_quadrics;             /* general quadrics among the surfaces               */
  const double* quad_coef;        /* [n_quadrics][10]: kept apart so that the common surfaces stay 32 bytes */

  /* ---- cells: model::cells (include/openmc/cell.h:339-375) ---------------- */
  int32_t n_cells;
  const int32_t* cell_id;         /* user ids                                          */
  const int32_t* cell_type;       /* OMCGPU_FILL_*                                     */
  const int32_t* cell_universe;   /* universe this cell belongs to                     */
  const int32_t* cell_fill;       /* universe / lattice index, -1 for material cells   */
  const int32_t* cell_simple;     /* Region::simple_                                   */
  const int32_t* cell_region_ofs; /* [n_cells+1] into region_tokens                    */
  const int32_t* region_tokens;   /* +-(surface index+1) or OMCGPU_OP_*                */
  const int32_t* cell_mat_ofs;    /* [n_cells+1] into cell_material (1 or n_instances) */
  const int32_t* cell_material;   /* material index, -1 = MATERIAL_VOID                */
  const int32_t* cell_sqrtkT_ofs; /* [n_cells+1] into cell_sqrtkT                      */
  const double* cell_sqrtkT;      /* sqrt(k_B T) per instance (or 1 entry)             */
  const int32_t* cell_density_ofs;/* [n_cells+1] into cell_density_mult                */
  const double* cell_density_mult;/* Cell::density_mult_ (src/cell.cpp:104-111): density of the instance / the
                                     material's density, per instance or 1 entry; {1.0} when the cell gives none */
  const double* cell_translation; /* [n_cells][3]                                      */
  const int32_t* cell_rotated;    /* [n_cells] fill cell has a rotation (Cell::rotation_ non-empty)   */
  const double* cell_rotation;    /* [n_cells][9] rotation matrix as Cell::set_rotation stores it (src/cell.cpp:48-90) */
  const int32_t* cell_distribcell_index; /* -1 = none                                  */
  const int32_t* cell_offset_ofs; /* [n_cells+1] into cell_offset                      */
  const int32_t* cell_offset;     /* Cell::offset_[map]                                */

  /* ---- universes: model::universes ---------------------------------------- */
  int32_t n_universes;
  const int32_t* univ_cell_ofs;   /* [n_universes+1]                                   */
  const int32_t* univ_cells;      /* Universe::cells_ in search order                  */
  /* derived by omcgpu_init (ignored on input): cells whose region names a surface
   * half-space -- the search list after a surface crossing, in place of the
   * reference's run-time neighbour lists (src/geometry.cpp:256-282).  Token slot
   * 2*s is surface s negative side (-(s+1)), 2*s+1 its positive side. */
  const int32_t* tok_cell_ofs;    /* [2*n_surfaces+1]                                  */
  const int32_t* tok_cells;       /* cell indices, ascending                           */

  /* ---- rectangular lattices: model::lattices (src/lattice.cpp:162-377) ---- */
  int32_t n_lattices;
  const int32_t* lat_shape;       /* [n_lattices][3]  n_cells_                         */
  const int32_t* lat_is3d;        /* [n_lattices]                                      */
  const int32_t* lat_type;        /* [n_lattices] OMCGPU_LATTICE_*; hex (src/lattice.cpp:458-1009): lat_shape =
                                     (2 n_rings - 1, 2 n_rings - 1, n_axial), lat_pitch = (radial, axial), lat_lower_left =
                                     centre, lat_universes = the skewed square array with -1 outside the hexagon */
  int32_t n_hex_lattices;         /* derived by omcgpu_init (ignored on input): lattices with lat_type != RECT */
  const int32_t* lat_outer;       /* [n_lattices] outer universe or -1                 */
  const double* lat_pitch;        /* [n_lattices][3]                                   */
  const double* lat_lower_left;   /* [n_lattices][3]                                   */
  const int32_t* lat_univ_ofs;    /* [n_lattices+1] into lat_universes                 */
  const int32_t* lat_universes;   /* x fastest, then y, then z (reader order)          */
  const int32_t* lat_offset_ofs;  /* [n_lattices+1] into lat_offsets                   */
  const int32_t* lat_offsets;     /* [n_maps][nz*ny*nx] distribcell offsets            */

  /* ---- materials: model::materials (include/openmc/material.h:185-203) ---- */
  int32_t n_materials;
  const int32_t* mat_id;
  const int32_t* mat_fissionable;
  const int32_t* mat_nuc_ofs;     /* [n_materials+1]                                   */
  const int32_t* mat_nuclide;     /* nuclide index                                     */
  const double* mat_atom_density; /* atom/b-cm                                         */
  const int32_t* mat_base;        /* derived by omcgpu_init (ignored on input): NULL, or for a model with several
                                     temperatures per nuclide the model material each device material was derived from */

  /* ---- continuous-energy nuclides: data::nuclides (src/nuclide.cpp) ------- */
  /* A "table" is one nuclide at one loaded temperature (Nuclide::kTs_[t], grid_[t], xs_[t],
   * Reaction::xs_[t]); every nuc_* / rx_* array below has one entry per TABLE.  With one
   * temperature per nuclide (n_nuclide_temps == 0) tables and nuclides coincide and
   * mat_nuclide holds table indices.  Otherwise mat_nuclide holds nuclide indices d and the
   * tables of nuclide d are nuct_ofs[d] .. nuct_ofs[d+1]-1 in ascending kT; a lookup takes
   * the one nearest to the cell's kT (TemperatureMethod::NEAREST, src/nuclide.cpp:682-691;
   * stochastic interpolation is refused).  omcgpu_init resolves that choice once per
   * (material, cell temperature) pair into derived materials, so the kernels pay nothing. */
  int32_t n_nuclide_temps;        /* 0, or the number of distinct nuclides               */
  const int32_t* nuct_ofs;        /* [n_nuclide_temps+1] into the tables                  */
  int32_t n_nuclides;             /* number of tables                                    */
  int32_t n_log_bins;             /* settings::n_log_bins (8000)                       */
  double log_spacing;             /* simulation::log_spacing                           */
  const double* hash_edge;        /* [n_log_bins+2] smallest E whose host-libm hash bin is k */
  /* derived by omcgpu_init (ignored on input): fine hash, OMCGPU_FINE_PER_LOG_BIN
   * equal-lethargy bins per reference bin; fine_window[bin][nuc] = {lo, hi}
   * brackets the number of grid points below any energy of the (widened) bin */
  int32_t n_fine_bins;
  int32_t reserved_ce;
  double fine_inv_spacing;        /* n_fine_bins / log(energy_max / energy_min)       */
  const int32_t* fine_window;     /* [n_fine_bins][n_nuclides][2]                      */
  const double* nuc_awr;          /* [n_nuclides]                                      */
  const double* nuc_kT;           /* [n_nuclides] kT of the table                      */
  const int32_t* nuc_fissionable; /* [n_nuclides]                                      */
  const int64_t* nuc_grid_ofs;    /* [n_nuclides+1] into nuc_energy / nuc_xs           */
  const double* nuc_energy;       /* concatenated energy grids                         */
  const double* nuc_xs;           /* [n_points][4] total, absorption, fission, nu-fission */
  const int32_t* nuc_grid_index;  /* [n_nuclides][n_log_bins+1]                        */
  /* reactions used by collision physics (src/physics.cpp:674-752) */
  const int32_t* nuc_rx_ofs;      /* [n_nuclides+1] into rx_*; reaction 0 = elastic    */
  const int32_t* nuc_fission_rx;  /* [n_nuclides] index into rx_* of first fission rx, -1 */
  const int32_t* nuc_n_precursor; /* [n_nuclides]                                      */
  const int32_t* nuc_total_nu_fn; /* [n_nuclides] function index of total_nu or -1     */
  int32_t n_reactions;
  const int32_t* rx_mt;           /* [n_reactions]                                     */
  const int32_t* rx_scatter_in_cm;/* [n_reactions]                                     */
  const int32_t* rx_inelastic;    /* [n_reactions] is in index_inelastic_scatter_      */
  const int32_t* rx_threshold;    /* [n_reactions] threshold grid index                */
  const int64_t* rx_xs_ofs;       /* [n_reactions+1] into rx_xs                        */
  const double* rx_xs;            /* concatenated reaction cross sections              */
  const double* rx_q_value;       /* [n_reactions]                                     */
  const int32_t* rx_prod_ofs;     /* [n_reactions+1] into prod_* (neutron products)    */
  int32_t n_products;
  const int32_t* prod_emission;   /* 0 prompt, 1 delayed, 2 total                      */
  const double* prod_decay_rate;  /* [n_products]                                      */
  const int32_t* prod_yield_fn;   /* function index                                    */
  const int32_t* prod_elaw;       /* OMCGPU_ELAW_*                                     */
  const double* prod_par;         /* [n_products][2] law parameters                    */
  const int32_t* prod_fn;         /* [n_products][2] law function indices              */
  int32_t n_functions;
  const int32_t* fn_type;         /* OMCGPU_FN_*                                       */
  const int32_t* fn_ofs;          /* [n_functions+1] into fn_data                      */
  const double* fn_data;

  /* ---- multigroup data: data::mg (src/mgxs*.cpp, xsdata.cpp, scattdata.cpp) */
  int32_t mg_n_groups;
  int32_t mg_n_delayed;
  const double* mg_energy_bin_avg;/* [G]                                               */
  const double* mg_energy_bins;   /* [G+1] descending (data::mg.energy_bins_)          */
  /* one macroscopic Mgxs per material: data::mg.macro_xs_[m] */
  const int32_t* mgm_n_temps;     /* [n_materials]                                     */
  const int32_t* mgm_n_pol;       /* [n_materials]                                     */
  const int32_t* mgm_n_azi;       /* [n_materials]                                     */
  const int32_t* mgm_isotropic;   /* [n_materials]                                     */
  const int32_t* mgm_fissionable; /* [n_materials]                                     */
  const int32_t* mgm_blk_ofs;     /* [n_materials+1] XsData blocks (one per temperature) */
  const double* mgm_kTs;          /* [n_blocks] kT of each block                       */
  /* per block b, angle a: flat index ba = blk_ang_ofs[b] + a */
  const int32_t* blk_ang_ofs;     /* [n_blocks+1]                                      */
  const double* mg_total;         /* [n_ba][G]                                         */
  const double* mg_absorption;    /* [n_ba][G]                                         */
  const double* mg_nu_fission;    /* [n_ba][G]                                         */
  const double* mg_prompt_nu_fission; /* [n_ba][G]                                     */
  const double* mg_fission;       /* [n_ba][G]                                         */
  const double* mg_inverse_velocity; /* [n_ba][G]                                      */
  const double* mg_delayed_nu_fission; /* [n_ba][DG][G]                                */
  const double* mg_chi_prompt;    /* [n_ba][G][G]                                      */
  const double* mg_chi_delayed;   /* [n_ba][DG][G][G]                                  */
  const double* mg_decay_rate;    /* [n_ba][DG]                                        */
  /* scattering: one ScattData per ba */
  const int32_t* sc_format;       /* [n_ba] OMCGPU_SCATT_*                             */
  const int32_t* sc_order;        /* [n_ba] entries per (gin,gout): coeffs / mu points / bins */
  const int32_t* sc_gmin;         /* [n_ba][G]                                         */
  const int32_t* sc_gmax;         /* [n_ba][G]                                         */
  const int32_t* sc_gofs;         /* [n_ba][G] offset of (gin, i_gout=0) in sc_energy/sc_mult */
  const int32_t* sc_dofs;         /* [n_ba][G] offset of (gin, i_gout=0) in sc_dist/sc_fmu; i_gout adds order */
  const double* sc_scattxs;       /* [n_ba][G]                                         */
  const double* sc_energy;        /* ragged [gin][i_gout] outgoing-group pdf           */
  const double* sc_mult;          /* ragged [gin][i_gout]                              */
  const double* sc_maxval;        /* ragged [gin][i_gout] (Legendre rejection bound)   */
  const double* sc_dist;          /* ragged [gin][i_gout][order] coeffs or CDF         */
  const double* sc_fmu;           /* ragged [gin][i_gout][order] pdf (tabular/histogram) */
  const int32_t* sc_mu_ofs;       /* [n_ba] into sc_mu                                 */
  const double* sc_mu;            /* mu grids (tabular: order pts, histogram: order+1) */
  const double* sc_dmu;           /* [n_ba]                                            */

  /* ---- tallies (tracklength and collision estimators; src/tallies/) --------- */
  int32_t n_meshes;
  const int32_t* mesh_ndim;       /* [n_meshes]                                        */
  const int32_t* mesh_shape;      /* [n_meshes][3]                                     */
  const double* mesh_lower_left;  /* [n_meshes][3]                                     */
  const double* mesh_width;       /* [n_meshes][3]                                     */
  int32_t n_filters;
  const int32_t* filter_type;     /* [n_filters] OMCGPU_FILTER_*                       */
  const int32_t* filter_n_bins;   /* [n_filters]                                       */
  const int32_t* filter_mesh;     /* [n_filters] mesh index (mesh filter)              */
  const int32_t* filter_matches_groups; /* [n_filters] EnergyFilter::matches_transport_groups_ */
  const int32_t* filter_data_ofs; /* [n_filters+1] into filter_idata / filter_ddata    */
  const int32_t* filter_idata;    /* cell/material/universe: map[index] -> bin or -1   */
  const double* filter_ddata;     /* energy: bin edges                                 */
  int32_t n_tallies;
  int32_t n_collision_tallies;    /* derived by omcgpu_init (ignored on input): tallies with the collision estimator */
  const int32_t* tally_estimator; /* [n_tallies] OMCGPU_ESTIMATOR_* (Tally::estimator_)       */
  const int32_t* tally_n_filters; /* [n_tallies]                                       */
  const int32_t* tally_filters;   /* [n_tallies][OMCGPU_MAX_FILTERS]                   */
  const int32_t* tally_strides;   /* [n_tallies][OMCGPU_MAX_FILTERS]                   */
  const int32_t* tally_n_scores;  /* [n_tallies]                                       */
  const int32_t* tally_scores;    /* [n_tallies][OMCGPU_MAX_SCORES]                    */
  const int64_t* tally_result_ofs;/* [n_tallies+1] into the dense VALUE buffer         */
} omcgpu_model;

/* ------------------------------------------------------------------------- */
/* per-generation inputs / outputs                                            */
/* ------------------------------------------------------------------------- */
typedef struct omcgpu_gen_in {
  int64_t n_particles;        /* source sites to transport on this device (work_per_rank) */
  int64_t id_offset;          /* simulation::work_index[rank]; particle id = id_offset+i+1 */
  int64_t particle_seed_base; /* (total_gen + overall_generation() - 1) * settings::n_particles */
  int64_t master_seed;        /* openmc_get_seed()                                       */
  uint64_t stride;            /* openmc_get_stride()                                     */
  double keff;                /* simulation::keff                                        */
  int32_t tallies_active;     /* score user tallies this generation                      */
  int32_t source_bank_clean;  /* 1: the host has not touched simulation::source_bank since the library wrote it
                                 (omcgpu_next_generation / omcgpu_sync_bank): the device-resident copy is current and the
                                 upload is skipped.  0: the host bank is authoritative (Python may edit it between batches,
                                 tests/unit_tests/test_lib.py:451-455) and is uploaded. */
} omcgpu_gen_in;

typedef struct omcgpu_gen_out {
  int64_t n_fission;          /* sites banked (fission bank already sorted by parent/progeny) */
  int64_t n_lost;             /* lost particles                                          */
  int64_t n_bank_full;        /* histories that found the fission bank full (sites dropped: the reference's warning, src/physics.cpp:236) */
  int64_t n_events;           /* total particle events processed                         */
  int64_t n_xs_lookups;       /* items processed by the calculate-xs kernel              */
  int64_t n_collisions;
  int64_t n_crossings;
  double global_tally[4];     /* generation increments, GlobalTally order                */
  double total_weight;        /* sum of source weights                                   */
  double ms_transport;        /* device time of the event loop (CUDA events)             */
  double ms_xs;               /* device time spent in the calculate-xs kernel            */
  double ms_surface;          /* sum of the launch durations of the crossing kernel (CUDA events on its stream) */
  double ms_collide;          /* the same for the collision kernel                       */
  double ms_score;            /* the same for the tally-scoring kernel                   */
  int64_t n_launch_surface;   /* launches behind ms_surface                              */
  int64_t n_launch_collide;
  int64_t n_launch_score;
  int64_t n_items_surface;    /* list items the crossing kernel processed (sweeps, without the tail) */
  int64_t n_items_collide;    /* the same for the collision kernel                       */
  /* with omcgpu_set_option("count_lookups", 1) -- [0] crossing kernel, [1] collision kernel,
   * [2] elsewhere (first flight, tail, stand-alone lookup kernel): */
  int64_t lookups_done[3];    /* macro-xs lookups actually evaluated (material / temperature changed) */
  int64_t nuclide_lookups[3]; /* sum over those of the material's nuclide count          */
  int64_t n_red;              /* RED.F64 issued by the scoring kernel's mesh ray trace   */
} omcgpu_gen_out;

/* ------------------------------------------------------------------------- */
/* entry points                                                               */
/* ------------------------------------------------------------------------- */

/* Upload the flattened model to `device`, allocate particle SoA / queues /
 * banks for up to max_particles histories in flight per generation (replaces
 * init_event_queues, src/event.cpp:29-38, and allocate_banks,
 * src/simulation.cpp:332-357).  Fails with OMCGPU_E_CUDA if no CUDA device is
 * present and OMCGPU_E_UNSUPPORTED if the model needs a feature the device path
 * refuses -- there is no CPU fallback. */
int omcgpu_init(const omcgpu_model* model, int device, int64_t max_particles);

/* Re-upload mutable properties of an initialised engine from the (re-flattened) host model: what
 * openmc_cell_set_temperature (include/openmc/capi.h:30) and openmc_material_set_densities /
 * openmc_material_set_density (:98) change between batches.  Array shapes must equal those given to
 * omcgpu_init; anything else needs a new omcgpu_init. */
#define OMCGPU_UPDATE_CELL_TEMPERATURES  1  /* cell_sqrtkT                         */
#define OMCGPU_UPDATE_MATERIAL_DENSITIES 2  /* mat_atom_density                    */
int omcgpu_update(const omcgpu_model* model, uint32_t what);

/* Release everything (replaces free_event_queues, src/finalize.cpp:57-59). */
int omcgpu_finalize(void);

/* Run-time knobs of an initialised engine (they survive until omcgpu_finalize):
 *   "count_lookups"  1: fill omcgpu_gen_out.n_lookups_done / n_nuclide_lookups (costs atomics)
 *   "tail_threshold" histories alive below which the generation is finished by the tail kernel
 *   "sync_every"     sweeps the host lets run between two counter read-backs
 *   "score_part"     mesh cells per part of a flight in the scoring kernel
 *   "probe_kT_neV"   cell temperature (kT in 1e-9 eV) the stand-alone lookup / collision probes use to
 *                    pick a nuclide's temperature table */
int omcgpu_set_option(const char* name, int64_t value);

/* How omcgpu_init arranged the event loop for this model: 2 = paired lists, crossing / collision
 * kernels that also do the lookup and the flight; 1 = merged alive list with one lookup + flight
 * kernel; 0 = merged alive list with the lookup as its own kernel over the (material, energy)-
 * sorted work list (many-nuclide models); -1 before omcgpu_init. */
int omcgpu_sweep_kind(void);

/* Last error text (like openmc_err_msg, src/error.cpp:33). */
const char* omcgpu_last_error(void);

/* Number of kernels launched by this library since omcgpu_init. */
int64_t omcgpu_kernel_launches(void);

/* One generation through host buffers -- the drop-in for
 * transport_event_based() (src/simulation.cpp:833-883) plus sort_fission_bank()
 * (src/bank.cpp:76-134): copies source_bank H2D, runs the event loop, sorts the
 * fission bank on the device, copies it D2H into fission_bank (capacity
 * fission_capacity sites) and fills `out`. */
int omcgpu_transport_generation(const omcgpu_gen_in* in, const omcgpu_site* source_bank,
  omcgpu_site* fission_bank, int64_t fission_capacity, omcgpu_gen_out* out);

/* One generation of a single-process run INCLUDING the sampling half of
 * synchronize_bank() (src/eigenvalue.cpp:152-178): source_bank (host,
 * in->n_particles sites) goes H2D, the event loop runs, the fission bank is
 * sorted and combed on the device with seed_id = simulation::total_gen +
 * overall_generation(), and the NEXT source bank comes back D2H into the same
 * buffer -- what libopenmc holds in simulation::source_bank after
 * finalize_generation() (src/simulation.cpp:522-571) when n_procs == 1.  The
 * sorted fission bank stays readable through omcgpu_get_fission_bank(). */
int omcgpu_next_generation(const omcgpu_gen_in* in, omcgpu_site* source_bank, int64_t seed_id,
  omcgpu_gen_out* out);

/* Page-lock a host bank (simulation::source_bank / fission_bank, allocated once
 * in allocate_banks(), src/simulation.cpp:332-357) so that the calls above DMA
 * straight from / into it instead of through the library's staging buffer. */
int omcgpu_host_register(void* p, int64_t bytes);
int omcgpu_host_unregister(void* p);

/* The same pieces separately, for a device-resident bank between generations. */
int omcgpu_set_source_bank(const omcgpu_site* source_bank, int64_t n);
int omcgpu_get_source_bank(omcgpu_site* source_bank, int64_t n);
int omcgpu_run_generation(const omcgpu_gen_in* in, omcgpu_gen_out* out);

/* Particle::mark_as_lost (src/particle.cpp:762-785) prints a warning with the particle's id and
 * writes a particle restart file per lost particle.  The device counts lost particles
 * (omcgpu_gen_out.n_lost) and keeps the state of the first OMCGPU_MAX_LOST_RECORDS of the last
 * generation: records[k] = {id, x, y, z, u, v, w, E} at the moment the particle was lost.
 * Returns the number of records written (<= capacity), or a negative error code. */
#define OMCGPU_LOST_RECORD 8
#define OMCGPU_MAX_LOST_RECORDS 64
int64_t omcgpu_lost_particles(double* records, int64_t capacity);

/* The device library comes in two builds of the same sources (openmc_b200/csrc/build.py):
 * libomcgpu.so, and libomcgpu_ext.so with the rarely needed features compiled in -- hexagonal
 * lattices (src/lattice.cpp:458-1009), axis cones and general quadrics, white and rotational
 * periodic boundaries, analog-estimator tallies, and secondary neutrons of reactions with an
 * integer yield (src/physics.cpp:1158-1162, src/particle.cpp:417-473).  Compiled into every kernel, their
 * dispatch costs 2-4 % even on models that use neither (profiles/r68-r70, r75), so the caller
 * asks, before omcgpu_init, which build the model needs: 1 = libomcgpu_ext.so.  omcgpu_init of
 * the other build refuses such a model by name.  Pure host code: no device is needed. */
int omcgpu_needs_extended_build(const omcgpu_model* model);
int omcgpu_get_fission_bank(omcgpu_site* fission_bank, int64_t capacity, int64_t* n);

/* Mesh::count_sites (src/mesh.cpp:941-997) over the fission bank the last generation left ON THE
 * DEVICE: the weight of the sites in every bin of a regular mesh (n_dim axes: shape, lower_left,
 * width), *sites_outside = 1 if some site lies outside it.  shannon_entropy (src/eigenvalue.cpp:579-607)
 * reads simulation::fission_bank, which stays on the device in the drop-in: the glue gets the
 * counts here and finishes the sum as the reference does.  Valid until the next generation starts. */
int omcgpu_count_fission_sites(int32_t n_dim, const int32_t* shape, const double* lower_left, const double* width,
  double* counts, int64_t n_bins, int32_t* sites_outside);

/* Uniform-comb resampling of the (sorted, device-resident) fission bank into
 * the device-resident source bank -- the sampling half of synchronize_bank()
 * (src/eigenvalue.cpp:152-178).  `start`/`total` are this rank's offset in and
 * the size of the global fission bank, n_particles the global source size,
 * seed_id = simulation::total_gen + overall_generation().  On return the
 * sampled sites [0, *n_sampled) are in the device source bank in global order.
 */
int omcgpu_comb_fission_bank(int64_t start, int64_t total, int64_t n_particles,
  int64_t seed_id, int64_t master_seed, uint64_t stride, int64_t* n_sampled);

/* The host half of that comb on its own (no device needed): the indices, relative to
 * `start`, of the sites among this rank's n_local sorted fission sites that the comb
 * selects -- bit-identical to the reference's recurrence tooth += teeth_distance. */
int omcgpu_comb_indices(int64_t start, int64_t n_local, int64_t total, int64_t n_particles,
  int64_t seed_id, int64_t master_seed, uint64_t stride, int32_t* idx, int64_t capacity,
  int64_t* n_sampled);

/* Device pointers of the SoA bank columns for the multi-GPU exchange
 * (src/eigenvalue.cpp:186-348): which = 0 source bank, 1 fission bank.
 * cols[0..8] = x y z u v w E time wgt (double*), cols[9] = delayed_group
 * (int32_t*), cols[10] = parent_id, cols[11] = progeny_id (int64_t*). */
int omcgpu_bank_columns_device(int which, void** cols, int64_t* n, int64_t* capacity);
int omcgpu_set_source_count(int64_t n);

/* ------------------------------------------------------------------------- */
/* Multi-GPU: one process per GPU of one box.  Geometry and data  */
/* are replicated, particle ids and seeds are global, rank r owns the global     */
/* source sites [work_index[r], work_index[r+1]) (calculate_work,                */
/* src/simulation.cpp:657-680).  The transport is NCCL (bound at run time from   */
/* libnccl.so.2) or three caller-supplied functions, e.g. a CUDA-aware MPI.      */
/* ------------------------------------------------------------------------- */
#define OMCGPU_COMM_ID_BYTES 128

/* Rank 0: a fresh NCCL unique id (ncclGetUniqueId) to hand to every rank by any side channel
 * (MPI_Bcast in libopenmc; a TCP store in bench.py). */
int omcgpu_comm_unique_id(char* id);

/* Join the communicator on the engine's device (after omcgpu_init); connects to every peer. */
int omcgpu_comm_init(const char* id, int rank, int n_ranks);

/* The same with a caller-supplied transport.  Buffers are DEVICE memory (the engine's stream is
 * synchronised before a call); every function returns 0 on success.
 *   allgather: every rank contributes n values, recv = [n_ranks][n]   (host memory)
 *   sendrecv:  n_ops simultaneous transfers with peers, complete on return
 *   allreduce: in-place sum of n doubles; root < 0: result on every rank, else on `root` only */
typedef int (*omcgpu_allgather_i64_fn)(void* user, const int64_t* send, int32_t n, int64_t* recv);
typedef int (*omcgpu_sendrecv_fn)(void* user, int32_t n_ops, const int32_t* peer, const int32_t* is_send,
  void* const* buf, const int64_t* bytes);
typedef int (*omcgpu_allreduce_f64_fn)(void* user, double* buf, int64_t n, int32_t root);
int omcgpu_comm_init_custom(int rank, int n_ranks, omcgpu_allgather_i64_fn allgather,
  omcgpu_sendrecv_fn sendrecv, omcgpu_allreduce_f64_fn allreduce, void* user);
int omcgpu_comm_finalize(void);

/* All-gather of n <= 4 int64 per rank (host values): also the barrier and the max-over-ranks
 * timing of bench.py. */
int omcgpu_comm_allgather_i64(const int64_t* send, int32_t n, int64_t* recv);
int64_t omcgpu_comm_bytes_exchanged(void);

/* synchronize_bank() (src/eigenvalue.cpp:86-365) over the ranks, on the device-resident banks left
 * by omcgpu_run_generation: all-gather of the fission-bank sizes, the comb with the common seed
 * (seed_id = simulation::total_gen + overall_generation()), all-gather of the sampled counts,
 * transfer of contiguous column ranges to their owners, and -- in the same call -- the sum over
 * the ranks of `globals` (the generation's k-collision / k-absorption / k-tracklength / leakage /
 * total weight; MPI_Allreduce of src/eigenvalue.cpp:63; in place, on every rank).  On return the
 * device source bank holds this rank's share of the next generation (*n_source sites) and, if
 * source_bank != NULL, so does the host's simulation::source_bank.  A rank without fission sites
 * or with an overflowing comb makes EVERY rank return the error. */
int omcgpu_sync_bank(int64_t n_particles, const int64_t* work_index, int64_t seed_id,
  int64_t master_seed, uint64_t stride, double* globals, int32_t n_globals,
  omcgpu_site* source_bank, int64_t* n_source);

/* reduce_tally_results() (src/tallies/tally.cpp:997-1025): device VALUE arrays of all tallies
 * summed onto rank `root` and zeroed elsewhere; root < 0: the sum on every rank. */
int omcgpu_reduce_tallies(int root);

/* Tally VALUE buffers: add the device VALUE array of tally `t` into
 * value[n_filter_bins*n_scores] (host) and zero it on the device -- the hand-off
 * into results_(:,:,VALUE) before Tally::accumulate (src/tallies/tally.cpp:821-878). */
int omcgpu_tally_results(int32_t tally, double* value, int64_t n);
int omcgpu_tally_values_device(double** values, int64_t* n);

/* Macroscopic cross-section lookup over a work list -- the kernel behind
 * process_calculate_xs_events (src/event.cpp:73-105) exposed on its own:
 * for i < n: (E[i], material[i]) -> out[i][4] = total, absorption, fission,
 * nu-fission and, when non-null, i_grid[i][k] for the k-th nuclide of the
 * material (max_nuc entries per particle).  MG: E holds the group index. */
int omcgpu_calculate_xs(const double* E, const int32_t* material, int64_t n,
  double* out, int32_t* i_grid, int32_t max_nuc);
/* Device-resident variant used by bench.py: pointers are device memory. */
int omcgpu_calculate_xs_device(const double* E, const int32_t* material, int64_t n,
  double* out, float* ms);

/* Geometry probes (openmc_find_cell, src/geometry.cpp:459-476, and
 * distance_to_boundary, :361-453) over n points, for bit-exact index parity. */
int omcgpu_find_cell(const double* xyz, const double* uvw, int64_t n,
  int32_t* cell, int32_t* instance, int32_t* material, int32_t* n_coord);
int omcgpu_distance_to_boundary(const double* xyz, const double* uvw, int64_t n,
  double* distance, int32_t* surface, int32_t* level, int32_t* lattice_trans);

/* Collision probe: collision() (src/physics.cpp:43-168; MG: src/physics_mg.cpp:26-60) on n
 * independent neutrons of weight `wgt` at the origin -- the device function of the collision
 * kernels on its own, for parity against vectors written by the reference's collision().
 * Neutron k: material[k], E[k] (MG: group), direction uvw[3k..], RNG state seed[k]; out: energy
 * (group), direction, weight, k-absorption score, RNG state, n_sites[k] banked fission sites;
 * the sites of all neutrons in (neutron, progeny) order in sites[0 .. *n_sites_total) with
 * parent_id = k + 1.  n <= max_particles. */
int omcgpu_collide(int64_t n, const int32_t* material, const double* E, const double* uvw,
  const uint64_t* seed, double keff, double wgt, double* E_out, double* uvw_out, double* wgt_out,
  double* k_abs_out, uint64_t* seed_out, int32_t* n_sites, omcgpu_site* sites,
  int64_t site_capacity, int64_t* n_sites_total);

/* RNG probes (src/random_lcg.cpp): out[i] = future_seed(n[i], seed[i]);
 * prn stream: out[i*count + j] = j-th prn of seed[i]. */
int omcgpu_future_seed(const uint64_t* n, const uint64_t* seed, int64_t count, uint64_t* out);
int omcgpu_prn(const uint64_t* seed, int64_t n_seeds, int32_t count, double* out);

#ifdef __cplusplus
}
#endif
#endif /* OMCGPU_H */
