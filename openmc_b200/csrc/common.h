// common.h -- shared definitions of the device engine.
//
// The event kernels are written once as `OMC_HD` functions over structure-of-
// arrays particle state.  nvcc compiles them for sm_100a (the product,
// libomcgpu.so).  A plain g++ compile of the very same sources with -DOMC_EMU
// exists ONLY for the CPU unit tests under tests/ (tests/_emu, never loaded by
// the package): it lets the host-side scheduler, queue logic and the device
// functions be checked against the oracle without a GPU.  It is not a fallback:
// omcgpu_init() of the product library fails when no CUDA device is present.
#pragma once

#include <stdint.h>

#include "../../include/omcgpu.h"

#if defined(__CUDACC__) && !defined(OMC_EMU)
#define OMC_CUDA 1
#define OMC_HD __device__ __forceinline__          /* device functions reading the constant-memory model */
#define OMC_BOTH __host__ __device__ __forceinline__ /* pure functions also used by host code (RNG) */
#define OMC_D __device__ __forceinline__
#define OMC_OOL __device__ __noinline__            /* out of line: called from many places, rarely hot (code size / icache) */
#else
#define OMC_CUDA 0
#define OMC_HD inline
#define OMC_BOTH inline
#define OMC_D inline
#define OMC_OOL inline
#include <cmath>
#include <cstring>
#endif

namespace omc {

// ---- constants: include/openmc/constants.h:49-54,61,81-83,112,337,377 ---------
constexpr double TINY_BIT = 1e-8;
constexpr double FP_PRECISION = 1e-14;
constexpr double FP_REL_PRECISION = 1e-5;
constexpr double FP_COINCIDENT = 1e-12;
constexpr int MAX_SAMPLE = 100000;
constexpr double PI = 3.141592653589793238462643383279502884;
constexpr double INFTY = 1.7976931348623157e308; // std::numeric_limits<double>::max()
constexpr double MASS_NEUTRON_EV = 939.56542052e6;
constexpr double C_LIGHT = 2.99792458e10;
constexpr int C_NONE = -1;
constexpr int MATERIAL_VOID = -1;
constexpr int SURFACE_NONE = 0;
constexpr int32_t OP_UNION = OMCGPU_OP_UNION;
constexpr int32_t OP_INTERSECTION = OMCGPU_OP_INTERSECTION;
constexpr int32_t OP_COMPLEMENT = OMCGPU_OP_COMPLEMENT;
constexpr int32_t OP_RIGHT_PAREN = OMCGPU_OP_RIGHT_PAREN;
constexpr int32_t OP_LEFT_PAREN = OMCGPU_OP_LEFT_PAREN;

OMC_HD double inf_d()
{
#if OMC_CUDA
  return __longlong_as_double(0x7ff0000000000000LL);
#else
  return INFINITY;
#endif
}

struct V3 {
  double x, y, z;
};
OMC_HD V3 v3(double x, double y, double z)
{
  V3 v;
  v.x = x;
  v.y = y;
  v.z = z;
  return v;
}
OMC_HD double dot(V3 a, V3 b)
{
  return a.x * b.x + a.y * b.y + a.z * b.z;
}
OMC_HD double comp(V3 a, int i)
{
  return i == 0 ? a.x : (i == 1 ? a.y : a.z);
}

// -----------------------------------------------------------------------------
// Particle state in HBM (last paragraph): a structure of
// arrays whose elements are 32-byte SECTORS, not scalars.  Work lists hold
// particle indices in arbitrary order, so a warp's accesses to a column are a
// gather; with one scalar per column every 8-byte access would move a whole
// 32-byte sector.  Fields that one event kernel reads together therefore share
// a sector (a "group": 4 doubles, or doubles and int32 mixed), and a gather
// moves only bytes the kernel uses.  `Col<T, S>` is a field inside a group:
// element i lives at p[i * S], S = group size in units of T.  Level-indexed
// fields are level-major: index (level * cap + i).
// -----------------------------------------------------------------------------
template<class T, int S>
struct Col {
  T* p;
#if defined(__CUDACC__) && !defined(OMC_EMU)
  __host__ __device__ __forceinline__
#else
  inline
#endif
  T& operator[](int64_t i) const { return p[i * S]; }
};

struct PState {
  int64_t cap; // stride between coordinate levels = particles allocated
  // A PState may also describe ONE particle staged in thread-local memory (the tail kernel keeps
  // the history it runs to completion there: cap = 1, index 0): `id_shift` is then the particle's
  // index in the generation, so that ids and seeds stay global
  int64_t id_shift;
  int32_t is_local;
  // many-nuclide materials: running sums of rho_k sigma_t,k after every XS_PREFIX_STRIDE nuclides,
  // written by the macro lookup and read by the nuclide-sampling walk of the collision, which
  // then starts at the right block instead of at nuclide 0 (n_prefix = 0: not kept)
  int32_t n_prefix;
  double* xs_prefix; // [cap][n_prefix], always a global column
  // group A[level]: position + which cell / lattice the level is in
  Col<double, 4> rx, ry, rz;
  Col<int32_t, 8> cell, lattice;
  // group B[level]: direction + universe
  Col<double, 4> ux, uy, uz;
  Col<int32_t, 8> universe;
  // group C[level] (16 bytes): lattice indices, touched only inside lattices
  Col<int32_t, 4> lat_x, lat_y, lat_z;
  // group P1: what a flight and a collision update
  Col<double, 4> E, wgt, time;
  Col<uint64_t, 4> seed;
  // group P2: macroscopic cross sections (written by k_xs)
  Col<double, 4> xs_total, xs_absorption, xs_fission, xs_nu_fission;
  // group P3: k-effective / leakage accumulators of the history
  Col<double, 4> k_abs, k_col, k_trk, k_leak;
  // group P4: integer state
  Col<int32_t, 8> n_coord, surface, material, material_last, cell_instance, cell_last0, g, n_event;
  // group P5: temperatures + the boundary found by the last flight
  Col<double, 4> sqrtkT, sqrtkT_last;
  Col<int32_t, 8> b_surface, b_level_trans, g_last;
  // plain column (scanned by cub for the fission-bank sort)
  int32_t* n_progeny;
  // plain column: the event the history waits for after this sweep (0 none / dead,
  // 1 collision, 2 crossing) -- input of the order-preserving list build
  uint8_t* next_ev;
};

// Flight records: what tracklength scoring needs from one flight, written by
// k_advance at the flight's position in the work list (coalesced) and by the
// tail kernel through an atomic append; consumed by k_score.  Self-contained,
// so scoring does not depend on particle state that later events overwrite.
constexpr int XS_PREFIX_STRIDE = 32;

struct Flights {
  int64_t cap;
  double *x, *y, *z;       // start point (r_last of include/openmc/particle_data.h)
  double *ux, *uy, *uz;    // direction at level 0
  double *total;           // |r1 - r0| (src/mesh.cpp:1008-1010)
  double *flux;            // wgt * flight distance
  double *E;
  double *xs_t, *xs_a, *xs_f, *xs_nf;
  int32_t *material, *g, *mg_ba;
  int32_t *pidx;           // particle (for cell/universe filters), -1 when logged by the tail
  int32_t *count;          // records appended by the tail kernel
};

// Source / fission banks as SoA columns (the 104-byte AoS SourceSite is only the
// host-side wire format).
struct Bank {
  int64_t cap;
  double *x, *y, *z, *u, *v, *w, *E, *time, *wgt;
  int32_t* delayed_group;
  int64_t *parent_id, *progeny_id;
};

// Work lists (event queues): indices of particles, filled by warp-aggregated
// atomic appends (replaces SharedArray::thread_safe_append,
// include/openmc/shared_array.h:71-89).
struct Queue {
  int32_t* idx;
  int32_t* count;      // device counter
  const int32_t* base; // keyed lists: both kinds share one sorted array, this list starts at *base (null: 0)
};

// Per-generation constants + counters visible to every kernel.
struct GenCtx {
  int64_t n;          // histories this generation
  int64_t id_offset;  // work_index
  int64_t seed_base;  // particle_seed_base
  int64_t master_seed;
  uint64_t stride;
  double keff;
  int32_t tallies_active;
  int32_t score_part; // mesh cells per part handed to one lane by k_score
  int32_t tally_needs_cells; // some tally has a cell/universe filter (needs live particle state)
  // device counters
  int32_t* n_fission;  // fission bank fill
  int32_t* n_lost;
  int32_t* n_bank_full; // histories that found the fission bank full
  unsigned long long* counters; // [0] events [1] xs lookups [2] collisions [3] crossings [4] lookups evaluated [5] their nuclides
  int32_t count_lookups;       // fill counters [4], [5]
  double* tally_values;
  int64_t tally_replica_stride; // doubles between two replicas of the VALUE arrays (k_score spreads its REDs)
  int32_t tally_replicas;
  int64_t fission_cap;
  // Collision records for the tallies scored AFTER a collision (analog estimator; collision
  // estimator of multigroup runs with survival biasing): column k of particle idx is
  // an_d[k * an_cap + idx] / an_i[k * an_cap + idx] (AN_* below).  Null when the model has no
  // such tally -- the collision kernels then carry one uniform branch and no stores.
  double* an_d;
  int32_t* an_i;
  int64_t an_cap;
  // the first lost_cap lost particles of the generation: [slot][OMCGPU_LOST_RECORD] =
  // {id, x, y, z, u, v, w, E} (what Particle::mark_as_lost reports, src/particle.cpp:762-785)
  double* lost_rec;
  int32_t lost_cap;
  // Secondary neutrons of reactions with an integer yield >= 2, e.g. (n,2n) (Particle::create_secondary
  // src/particle.cpp:75-93, inelastic_scatter src/physics.cpp:1158-1162, event_revive_from_secondary
  // src/particle.cpp:417-473): a pool of sites {x, y, z, u, v, w, E, wgt, time} handed out by an atomic
  // counter, and one LIFO list per particle through sec_next.  Null when no reaction of the model can
  // create them: the kernels then carry one uniform branch.
  double* sec_site;
  int32_t* sec_next;
  int32_t* sec_head;  // [particles of the generation] first site of the particle's list, -1 = none
  int32_t* sec_count; // [0] sites handed out  [1] sites that did not fit
  int32_t sec_cap;
};
constexpr int SEC_ND = 9;
enum { AN_RX, AN_RY, AN_RZ, AN_E_LAST, AN_WGT_LAST, AN_WGT_BANK, AN_MICRO_T, AN_MICRO_A, AN_MICRO_F, AN_MICRO_NF,
  AN_XS_T, AN_XS_A, AN_XS_F, AN_XS_NF, AN_E_POST, AN_ND };
enum { AN_SCATTER, AN_FISSION, AN_ALIVE, AN_G_LAST, AN_G_POST, AN_MATERIAL, AN_MG_BA, AN_FIS_BASE, AN_NI };

} // namespace omc
