// engine.cu -- event kernels, the event-loop scheduler and the C ABI of
// include/omcgpu.h.
//
// Replaces, for one generation:
//   transport_event_based            src/simulation.cpp:833-883
//   process_init_events              src/event.cpp:62-71      -> k_init
//   process_calculate_xs_events      src/event.cpp:73-105     -> k_xs
//   process_advance_particle_events  src/event.cpp:107-128    -> k_advance
//   process_surface_crossing_events  src/event.cpp:130-147    -> k_surface
//   process_collision_events         src/event.cpp:149-166    -> k_collide
//   process_death_events             src/event.cpp:168-177    -> k_reduce
//   sort_fission_bank                src/bank.cpp:76-134      -> scan + k_scatter_bank
//   synchronize_bank (sampling half) src/eigenvalue.cpp:152-178 -> k_gather_bank
#include "device.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <tuple>
#include <string>
#include <vector>

#if OMC_CUDA
#include <cooperative_groups.h>
#include <cooperative_groups/reduce.h>
#include <cub/cub.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/discard_iterator.h>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
#else
#include "cuda_emu.h"
#include <cmath>
#include <numeric>
#endif

namespace omc {

#ifndef OMC_BLOCK
#define OMC_BLOCK 256
#endif
constexpr int BLOCK = OMC_BLOCK;
#ifndef XA_MINBLOCKS
#define XA_MINBLOCKS 4 /* resident blocks per SM asked of k_xs_advance / k_surface_xa (register cap) */
#endif
#ifndef CX_MINBLOCKS
#define CX_MINBLOCKS 4 /* the same for k_collide / k_collide_xa (measured: profiles/README.md r10-r11, r21) */
#endif

// -----------------------------------------------------------------------------
// kernel plumbing: every kernel is `body(tid, args)` for tid < n_ub.  Under
// nvcc this is a __global__ launch; in the test-only host compile a loop.
// -----------------------------------------------------------------------------
static int64_t g_launches = 0;

#if OMC_CUDA
// name##_geo: the kernel walks the CSG tables and is launched with the staged part of the geometry
// blob as dynamic shared memory (device.h, stage_geometry / gld)
#define OMC_KERNEL_MB(name, ArgT, body, MINBLOCKS) \
  constexpr bool name##_geo = false; \
  __global__ void __launch_bounds__(BLOCK, MINBLOCKS) name(int64_t n_ub, ArgT a) \
  { \
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; \
    if (tid < n_ub) \
      body(tid, a); \
  }
#define OMC_KERNEL_MB_GEO(name, ArgT, body, MINBLOCKS) \
  constexpr bool name##_geo = true; \
  __global__ void __launch_bounds__(BLOCK, MINBLOCKS) name(int64_t n_ub, ArgT a) \
  { \
    stage_geometry(); \
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; \
    if (tid < n_ub) \
      body(tid, a); \
  }
#define OMC_KERNEL(name, ArgT, body) OMC_KERNEL_MB(name, ArgT, body, 1)
#define OMC_KERNEL_GEO(name, ArgT, body) OMC_KERNEL_MB_GEO(name, ArgT, body, 1)
// The same body behind a persistent grid whose warps fetch the next 32 list positions from a device
// cursor (a.cursor, zero at launch) until the list is exhausted.  With one block per 256 positions a
// block keeps its registers and warp slots until its slowest warp is done -- the event kernels'
// lanes differ by an order of magnitude in work (cell search, rejection loops), and ncu showed 46 %
// of the warp slots occupied; here a warp that is done simply takes the next chunk.
#define OMC_KERNEL_DYN(name, ArgT, body, MINBLOCKS) \
  constexpr bool name##_geo = true; \
  __global__ void __launch_bounds__(BLOCK, MINBLOCKS) name(int64_t n_ub, ArgT a) \
  { \
    stage_geometry(); \
    const int lane = threadIdx.x & 31; \
    int64_t cnt = *a.q_in.count; \
    cnt = cnt < n_ub ? cnt : n_ub; \
    for (;;) { \
      int base = 0; \
      if (lane == 0) \
        base = atomicAdd(a.cursor, 32); \
      base = __shfl_sync(0xffffffffu, base, 0); \
      if (base >= cnt) \
        break; \
      if (base + lane < cnt) \
        body((int64_t)base + lane, a); \
      __syncwarp(); \
    } \
  }
#define OMC_LAUNCH_DYN(name, n_ub, stream, args, ctas_per_sm) \
  do { \
    if ((n_ub) > 0) { \
      unsigned grid_ = (unsigned)(((n_ub) + BLOCK - 1) / BLOCK); \
      const unsigned cap_ = (unsigned)(dyn_grid_sms() * (ctas_per_sm)); \
      name<<<grid_ < cap_ ? grid_ : cap_, BLOCK, name##_geo ? geo_smem_bytes() : 0, stream>>>((int64_t)(n_ub), args); \
      ++g_launches; \
    } \
  } while (0)
static int dyn_grid_sms();
static unsigned geo_smem_bytes();
#define OMC_LAUNCH(name, n_ub, stream, args) \
  do { \
    if ((n_ub) > 0) { \
      name<<<(unsigned)(((n_ub) + BLOCK - 1) / BLOCK), BLOCK, name##_geo ? geo_smem_bytes() : 0, stream>>>((int64_t)(n_ub), \
        args); \
      ++g_launches; \
    } \
  } while (0)
#else
#define OMC_KERNEL_MB(name, ArgT, body, MINBLOCKS) OMC_KERNEL(name, ArgT, body)
#define OMC_KERNEL_MB_GEO(name, ArgT, body, MINBLOCKS) OMC_KERNEL(name, ArgT, body)
#define OMC_KERNEL_GEO(name, ArgT, body) OMC_KERNEL(name, ArgT, body)
#define OMC_KERNEL_DYN(name, ArgT, body, MINBLOCKS) OMC_KERNEL(name, ArgT, body)
#define OMC_LAUNCH_DYN(name, n_ub, stream, args, ctas_per_sm) OMC_LAUNCH(name, n_ub, stream, args)
#define OMC_KERNEL(name, ArgT, body) \
  void name(int64_t n_ub, ArgT a) \
  { \
    for (int64_t tid = 0; tid < n_ub; ++tid) \
      body(tid, a); \
  }
#define OMC_LAUNCH(name, n_ub, stream, args) \
  do { \
    if ((n_ub) > 0) { \
      name((int64_t)(n_ub), args); \
      ++g_launches; \
    } \
  } while (0)
#endif

// kernels that contain the cross-section lookup exist in two flavours: `name` (one
// nuclide at a time) and `name_b` (lockstep walk for many-nuclide materials, larger
// register budget); G.batch_xs is decided per model in omcgpu_init
// With few histories alive (the last ~100 sweeps of a generation, the tail) a sweep costs
// the latency of ONE event chain, occupancy is irrelevant, and the lockstep walk -- which
// shortens the lookup's 4 dependent gathers per nuclide to 4 per batch -- is used for every
// material (xs_batch_min = 1).
#define LAUNCH_XS(name, n_ub, stream, args) \
  do { \
    const bool low_ = (int64_t)(n_ub) < low_count_threshold(); \
    (args).xs_batch_min = low_ ? 1 : OMC_XS_BATCH_MIN; \
    if (batch_xs_enabled() || low_) \
      OMC_LAUNCH(name##_b, n_ub, stream, args); \
    else \
      OMC_LAUNCH(name, n_ub, stream, args); \
  } while (0)
// crossing / collision kernels of the paired sweep: the persistent, work-fetching flavour when
// OMCGPU_DYNAMIC is on (and the model does not need the lockstep-lookup flavour)
#define LAUNCH_XA(name, n_ub, stream, args) \
  do { \
    if (dynamic_enabled() && !batch_xs_enabled() && (int64_t)(n_ub) >= low_count_threshold()) { \
      (args).xs_batch_min = OMC_XS_BATCH_MIN; \
      OMC_LAUNCH_DYN(name##_dyn, n_ub, stream, args, XA_MINBLOCKS); \
    } else { \
      LAUNCH_XS(name, n_ub, stream, args); \
    } \
  } while (0)
static bool dynamic_enabled();
static bool batch_xs_enabled();
static int64_t low_count_threshold();
static int tail_batch_min();

OMC_D int32_t list_item(const Queue& q, int64_t tid)
{
  return q.idx[(q.base ? *q.base : 0) + tid];
}

// warp-aggregated append (ballot + one atomic per warp), cf. north_star
OMC_D void queue_append(const Queue& q, int32_t value)
{
#if OMC_CUDA
  cg::coalesced_group g = cg::coalesced_threads();
  int32_t base = 0;
  if (g.thread_rank() == 0)
    base = atomicAdd(q.count, (int32_t)g.size());
  base = g.shfl(base, 0);
  q.idx[base + g.thread_rank()] = value;
#else
  q.idx[(*q.count)++] = value;
#endif
}

// The event kernels are chains of dependent gathers (work-list entry -> particle
// sectors -> model data), and the particle's sectors mostly miss L2.  Asking L2
// for all of them as soon as the particle index is known turns a chain of DRAM
// round trips into one (the later loads find their sector in L2 or in flight);
// a prefetch holds no register.
OMC_D void prefetch_l2(const void* p)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(__cvta_generic_to_global(p)));
#else
  (void)p;
#endif
}

#if OMC_CUDA
__constant__ int c_prefetch = 0; // OMCGPU_PREFETCH=1 turns the sector prefetch on (A/B measurements)
#endif

OMC_D void prefetch_particle(const PState& ps, int64_t i, bool collision)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  if (!c_prefetch)
    return;
  prefetch_l2(&ps.n_coord[i]);  // P4
  prefetch_l2(&ps.E[i]);        // P1
  prefetch_l2(&ps.xs_total[i]); // P2
  prefetch_l2(&ps.k_abs[i]);    // P3
  prefetch_l2(&ps.sqrtkT[i]);   // P5
  const int L = DM.n_coord_levels < 3 ? DM.n_coord_levels : 3;
  for (int l = 0; l < L; ++l) {
    prefetch_l2(&ps.rx[(int64_t)l * ps.cap + i]); // A[l]
    prefetch_l2(&ps.ux[(int64_t)l * ps.cap + i]); // B[l]
    if (DM.n_lattices > 0 && l > 0)
      prefetch_l2(&ps.lat_x[(int64_t)l * ps.cap + i]); // C[l]
  }
  if (collision)
    prefetch_l2(ps.n_progeny + i);
#else
  (void)ps;
  (void)i;
  (void)collision;
#endif
}

// -----------------------------------------------------------------------------
// event bodies
// -----------------------------------------------------------------------------
struct EvArgs {
  PState ps;
  GenCtx gc;
  Bank src;  // source bank (k_init)
  Bank fb;   // fission bank (k_collide)
  Queue q_in, q_a, q_b; // input list, and up to two output lists
  Flights fl;           // flight records for k_score
  int32_t fl_buf;       // which of the two record buffers `fl` is (c_fl[fl_buf])
  const int32_t* n_flights; // device count of records (its ring slot outlives the sweep)
  int32_t first_event;      // k_tail: event the histories of q_in are waiting for (NEXT_NONE: a flight)
  int32_t xs_batch_min;     // lockstep kernels: nuclides in a material from which the lockstep walk is used
  int32_t ordered;          // survivors are flagged in ps.next_ev (lists built in index order) instead of appended
  int32_t* cursor;          // work cursor of a persistent (work-fetching) launch, zero when the kernel starts
};

enum { NEXT_NONE = -2, NEXT_DEAD = -1, NEXT_COLLISION = 0, NEXT_SURFACE = 1 };
constexpr int32_t PIDX_LOGGED = -1;    // flight logged by the tail: no particle state to consult
constexpr int32_t PIDX_NO_FLIGHT = -2; // work-list position whose history made no flight this sweep

// rare, out of line, VALUES only (an out-of-line helper that takes `const PState&` makes its
// caller keep the 1 KB of column pointers in its stack frame: DESIGN.md, stack-frame pitfall --
// measured again here, profiles/r67: 1272-byte frames, +50 % on every event kernel)
OMC_OOL void record_lost(double* rec, int64_t id, V3 r, V3 u, double E)
{
  rec[0] = (double)id;
  rec[1] = r.x;
  rec[2] = r.y;
  rec[3] = r.z;
  rec[4] = u.x;
  rec[5] = u.y;
  rec[6] = u.z;
  rec[7] = E;
}
OMC_D void mark_lost(const PState& ps, const GenCtx& gc, int64_t i)
{ // Particle::mark_as_lost, src/particle.cpp:762-785 (count + kill; the warning is the host's)
  const int32_t slot = atomic_add_i32(gc.n_lost, 1);
#ifndef OMC_NO_LOST_REC /* measurement build */
  if (gc.lost_rec && slot < gc.lost_cap)
    record_lost(gc.lost_rec + (int64_t)slot * OMCGPU_LOST_RECORD, gc.id_offset + ps.id_shift + i + 1, get_r(ps, i, 0),
      get_u(ps, i, 0), ps.E[i]);
#else
  (void)slot;
#endif
  ps.wgt[i] = 0.0;
}

// initialize_history + Particle::from_source + first locate
// (src/simulation.cpp:573-649, src/particle.cpp:114-158, :176-193)
OMC_D bool ev_init(const PState& ps, const GenCtx& gc, const Bank& src, int64_t i)
{
  const int L = DM.n_coord_levels < OMCGPU_MAX_COORD ? DM.n_coord_levels : OMCGPU_MAX_COORD;
  for (int l = 0; l < L; ++l)
    reset_level(ps, i, l);
  set_r(ps, i, 0, v3(src.x[i], src.y[i], src.z[i]));
  set_u(ps, i, 0, v3(src.u[i], src.v[i], src.w[i]));
  ps.n_coord[i] = 1;
  ps.surface[i] = SURFACE_NONE;
  ps.material[i] = C_NONE;
  ps.material_last[i] = C_NONE;
  ps.sqrtkT[i] = -1.0;
  ps.sqrtkT_last[i] = 0.0;
  ps.cell_instance[i] = 0;
  const double wgt = src.wgt[i];
  ps.wgt[i] = wgt;
  ps.time[i] = src.time[i];
  if (DM.run_CE) {
    ps.E[i] = src.E[i];
    ps.g[i] = 0;
  } else {
    const int g = (int)src.E[i];
    ps.g[i] = g;
    ps.E[i] = ld(DM.mg_energy_bin_avg + g);
  }
  ps.n_progeny[i] = 0;
  ps.n_event[i] = 0;
  if (gc.sec_head != nullptr)
    gc.sec_head[ps.id_shift + i] = -1;
  ps.k_abs[i] = 0.0;
  ps.k_col[i] = 0.0;
  ps.k_trk[i] = 0.0;
  ps.k_leak[i] = 0.0;
  ps.xs_total[i] = 0.0;
  ps.xs_absorption[i] = 0.0;
  ps.xs_fission[i] = 0.0;
  ps.xs_nu_fission[i] = 0.0;
  ps.b_surface[i] = 0;
  ps.b_level_trans[i] = 0;
  const int64_t id = gc.id_offset + i + 1;
  const int64_t particle_seed = gc.seed_base + id;
  ps.seed[i] = future_seed((uint64_t)particle_seed * gc.stride, (uint64_t)gc.master_seed);
  if (wgt == 0.0)
    return false;
  if (!find_cell_from_root(ps, i, SURFACE_NONE)) {
    mark_lost(ps, gc, i);
    return false;
  }
  ps.cell_last0[i] = LV(cell, 0);
  return true;
}

// Particle::event_calculate_xs, src/particle.cpp:160-230 (locate done at birth)
// d_counters: [0] events [1] lookups [2] collisions [3] crossings (the tail kernel's share);
// pairs {lookups evaluated, their nuclides} per kernel kind; REDs issued by the scoring kernel
constexpr int CNT_SURFACE = 4, CNT_COLLIDE = 6, CNT_OTHER = 8, CNT_RED = 10, N_COUNTERS = 12;

// lookups actually evaluated and their nuclide count (omcgpu_set_option("count_lookups", 1):
// bench.py's byte model needs the mean number of nuclides per lookup; off in timed generations)
OMC_D void count_lookup(unsigned long long* counters, int n_nuc)
{
#if OMC_CUDA
  cg::coalesced_group g = cg::coalesced_threads();
  const unsigned long long tot = cg::reduce(g, (unsigned long long)n_nuc, cg::plus<unsigned long long>());
  if (g.thread_rank() == 0) {
    atomicAdd(counters, (unsigned long long)g.size());
    atomicAdd(counters + 1, tot);
  }
#else
  counters[0] += 1;
  counters[1] += (unsigned long long)n_nuc;
#endif
}

template<bool BX = false>
OMC_D void ev_xs(const PState& ps, int64_t i, int batch_min = OMC_XS_BATCH_MIN, unsigned long long* count = nullptr)
{
  const int m = ps.material[i];
  if (m == MATERIAL_VOID) {
    ps.xs_total[i] = 0.0;
    ps.xs_absorption[i] = 0.0;
    ps.xs_fission[i] = 0.0;
    ps.xs_nu_fission[i] = 0.0;
    return;
  }
  if (DM.run_CE) {
    if (m != ps.material_last[i] || ps.sqrtkT[i] != ps.sqrtkT_last[i]) {
      double t, a, f, nf;
      ce_macro_xs<BX>(m, ps.E[i], t, a, f, nf, nullptr, 0, batch_min,
        ps.n_prefix ? ps.xs_prefix + (ps.id_shift + i) * ps.n_prefix : nullptr);
      if (count)
        count_lookup(count, ld(DM.mat_nuc_ofs + m + 1) - ld(DM.mat_nuc_ofs + m));
      ps.xs_total[i] = t;
      ps.xs_absorption[i] = a;
      ps.xs_fission[i] = f;
      ps.xs_nu_fission[i] = nf;
    }
  } else {
    const int G = DM.mg_n_groups;
    const int n_coord = ps.n_coord[i];
    const int t = mg_temperature_index(m, ps.sqrtkT[i]);
    const int a = mg_angle_index(m, get_u(ps, i, n_coord - 1));
    const int ba = mg_ba(m, t, a);
    const int g = ps.g[i];
    // Cell::density_mult of the particle's cell instance (src/mgxs.cpp:620-625; extended build: continuous-
    // energy models fold it into their device materials at omcgpu_init and never come here)
    double mult = 1.0;
#if OMC_EXTENDED
    if (DM.cell_density_mult != nullptr) {
      const int c = LV(cell, n_coord - 1);
      const int d0 = ld(DM.cell_density_ofs + c), nd = ld(DM.cell_density_ofs + c + 1) - d0;
      if (nd > 0)
        mult = ld(DM.cell_density_mult + d0 + (nd > 1 ? ps.cell_instance[i] : 0));
    }
#endif
    ps.xs_total[i] = ld(DM.mg_total + ba * G + g) * mult;
    ps.xs_absorption[i] = ld(DM.mg_absorption + ba * G + g) * mult;
    ps.xs_nu_fission[i] = ld(DM.mgm_fissionable + m) ? ld(DM.mg_nu_fission + ba * G + g) * mult : 0.;
  }
}

// what the scores need from a particle in flight (read after the move, before
// the next event changes it)
OMC_D void load_score_ctx(const PState& ps, int64_t i, int n_coord, double distance, ScoreCtx& sc)
{
  const int m = ps.material[i];
  sc.flux = ps.wgt[i] * distance;
  sc.xs_total = ps.xs_total[i];
  sc.xs_absorption = ps.xs_absorption[i];
  sc.xs_fission = ps.xs_fission[i];
  sc.xs_nu_fission = ps.xs_nu_fission[i];
  sc.material = m;
  sc.g = ps.g[i];
  sc.mg_ba = 0;
  if (!DM.run_CE && m >= 0) // score_general_mg uses the lowest level's direction
    sc.mg_ba = mg_ba(m, mg_temperature_index(m, ps.sqrtkT[i]),
      mg_angle_index(m, get_u(ps, i, n_coord - 1)));
}

// flight record for k_score
OMC_D void record_flight(const Flights& fl, int64_t slot, const PState& ps, int64_t i, int n_coord,
  V3 r0, V3 u0, double E, double distance, int32_t pidx)
{
  ScoreCtx sc;
  load_score_ctx(ps, i, n_coord, distance, sc);
  fl.x[slot] = r0.x;
  fl.y[slot] = r0.y;
  fl.z[slot] = r0.z;
  fl.ux[slot] = u0.x;
  fl.uy[slot] = u0.y;
  fl.uz[slot] = u0.z;
  fl.total[slot] = track_length(r0, get_r(ps, i, 0));
  fl.flux[slot] = sc.flux;
  fl.E[slot] = E;
  fl.xs_t[slot] = sc.xs_total;
  fl.xs_a[slot] = sc.xs_absorption;
  fl.xs_f[slot] = sc.xs_fission;
  fl.xs_nf[slot] = sc.xs_nu_fission;
  fl.material[slot] = sc.material;
  fl.g[slot] = sc.g;
  fl.mg_ba[slot] = sc.mg_ba;
  fl.pidx[slot] = pidx;
}

// reserve one slot of the tail's flight log (warp-aggregated); -1 when full
OMC_D int64_t flight_log_reserve(const Flights& fl)
{
#if OMC_CUDA
  cg::coalesced_group g = cg::coalesced_threads();
  int32_t base = 0;
  if (g.thread_rank() == 0)
    base = atomicAdd(fl.count, (int32_t)g.size());
  base = g.shfl(base, 0);
  const int64_t slot = (int64_t)base + g.thread_rank();
#else
  const int64_t slot = (*fl.count)++;
#endif
  return slot < fl.cap ? slot : -1;
}

// Particle::event_advance, src/particle.cpp:232-286
// Tally scoring: slot >= 0 -- the flight is recorded at fl[slot] for k_score
// (sweeps); slot == SLOT_LOG -- appended to the tail's flight log, scored by this
// lane only if the log is full or a tally needs the particle's cells.
constexpr int64_t SLOT_LOG = -1;
// TAIL: the caller may pass SLOT_LOG (only the tail kernel does); the sweep kernels always own a
// record slot, and their code then holds no out-of-line scoring call that would force a staged
// particle's view into memory
struct AdvanceOut { // what the keyed list build wants to know of the flight (registers, not re-read)
  int material, b_surface, b_level_trans;
};
template<bool TAIL = false>
OMC_D int ev_advance(const PState& ps, const GenCtx& gc, const Flights& fl, int64_t i, int64_t slot,
  AdvanceOut* ao = nullptr)
{
  const int n_coord = ps.n_coord[i];
  const int surface = ps.surface[i];
  bool neg_lat;
  const Boundary b = distance_to_boundary(ps, i, n_coord, surface, neg_lat);
  if (neg_lat) {
    mark_lost(ps, gc, i);
    return NEXT_DEAD;
  }
  const double xs_total = ps.xs_total[i];
  double d_coll;
  if (xs_total == 0.0) {
    d_coll = inf_d();
  } else {
    uint64_t seed = ps.seed[i];
    d_coll = -log(prn(seed)) / xs_total;
    ps.seed[i] = seed;
  }
  const double distance = b.distance < d_coll ? b.distance : d_coll;
  const V3 r0 = get_r(ps, i, 0);
  const V3 u0 = get_u(ps, i, 0);
  for (int l = 0; l < n_coord; ++l) { // GeometryState::move_distance
    LV(rx, l) += distance * LV(ux, l);
    LV(ry, l) += distance * LV(uy, l);
    LV(rz, l) += distance * LV(uz, l);
  }
  const int m = ps.material[i];
  const double E = ps.E[i];
  const int g = ps.g[i];
  int ba = 0;
  double speed;
  if (DM.run_CE) { // Particle::speed, src/particle.cpp:46-73
    speed = C_LIGHT * sqrt(E * (E + 2 * MASS_NEUTRON_EV)) / (E + MASS_NEUTRON_EV);
  } else if (m >= 0) {
    ba = mg_ba(m, mg_temperature_index(m, ps.sqrtkT[i]), mg_angle_index(m, u0));
    speed = 1.0 / ld(DM.mg_inverse_velocity + ba * DM.mg_n_groups + g);
  } else {
    speed = inf_d();
  }
  ps.time[i] += distance / speed;
  const double wgt = ps.wgt[i];
  const double nu_fission = ps.xs_nu_fission[i];
  if (gc.tallies_active && DM.n_tallies > 0) {
    int32_t pidx = (int32_t)(ps.id_shift + i); // the particle's index in the generation
    if (slot == SLOT_LOG) {
      pidx = PIDX_LOGGED;
      if (!gc.tally_needs_cells)
        slot = flight_log_reserve(fl);
    }
    if (!TAIL || slot >= 0) {
      record_flight(fl, slot, ps, i, n_coord, r0, u0, E, distance, pidx);
    } else {
      ScoreCtx sc;
      load_score_ctx(ps, i, n_coord, distance, sc);
      score_tracklength(ps, i, n_coord, r0, get_r(ps, i, 0), u0, E, sc, gc.tally_values);
    }
  }
  ps.k_trk[i] += wgt * distance * nu_fission;
  ps.b_surface[i] = b.surface;
  ps.b_level_trans[i] = b.level | (b.trans << 8);
  if (ao) {
    ao->material = m;
    ao->b_surface = b.surface;
    ao->b_level_trans = b.level | (b.trans << 8);
  }
  return (d_coll > b.distance) ? NEXT_SURFACE : NEXT_COLLISION;
}

// Particle::event_revive_from_secondary, src/particle.cpp:417-473: a dead particle with secondary
// neutrons on its list becomes the last one created (Particle::from_source on the site: the random
// number streams, the id, the progeny counter and the k accumulators carry on)
OMC_D bool ev_revive(const PState& ps, const GenCtx& gc, int64_t i)
{
  const int n_event = ps.n_event[i] + 1;
  ps.n_event[i] = n_event;
  if (n_event == DM.max_particle_events)
    ps.wgt[i] = 0.0;
  if (ps.wgt[i] != 0.0)
    return true;
#if !OMC_SECONDARIES
  (void)gc;
  return false;
#else
  if (gc.sec_site == nullptr) // uniform: no reaction of the model creates secondaries
    return false;
  const int64_t idx = ps.id_shift + i;
  for (;;) {
    const int32_t slot = gc.sec_head[idx];
    if (slot < 0)
      return false;
    gc.sec_head[idx] = gc.sec_next[slot];
    atomic_add_u64(gc.counters, (unsigned long long)ps.n_event[i]); // the dead particle's events (counters[0])
    const double* s = gc.sec_site + (int64_t)slot * SEC_ND;
    const int L = DM.n_coord_levels < OMCGPU_MAX_COORD ? DM.n_coord_levels : OMCGPU_MAX_COORD;
    for (int l = 0; l < L; ++l)
      reset_level(ps, i, l);
    set_r(ps, i, 0, v3(s[0], s[1], s[2]));
    set_u(ps, i, 0, v3(s[3], s[4], s[5]));
    ps.n_coord[i] = 1;
    ps.surface[i] = SURFACE_NONE;
    ps.material[i] = C_NONE;
    ps.material_last[i] = C_NONE;
    ps.cell_instance[i] = 0;
    ps.E[i] = s[6];
    ps.g[i] = 0;
    ps.wgt[i] = s[7];
    ps.time[i] = s[8];
    ps.n_event[i] = 0;
    if (find_cell_from_root(ps, i, SURFACE_NONE)) {
      ps.cell_last0[i] = LV(cell, 0);
      return true;
    }
    mark_lost(ps, gc, i); // event_calculate_xs could not locate it: the next secondary, if any
  }
#endif
}

// Rare boundary conditions, out of line so that the crossing kernel does not carry them.  They take
// and return VALUES: a `const PState&` argument would make the caller keep a 1 KB copy of the column
// pointers in its stack frame (measured: +30 % on the crossing kernel, profiles/r59).
struct PosDir {
  V3 r, u;
};
// RotationalPeriodicBC::handle_particle, src/boundary_condition.cpp:242-287: position and direction
// rotated about x / y / z by the angle whose cosine and sine the flattener took with the host's libm
OMC_OOL PosDir rotate_periodic(V3 r, V3 u, int s, int axis)
{
  const double* t = DM.surf_periodic_shift + 3 * s;
  const double c = gld(t), sn = gld(t + 1);
  PosDir o;
  o.r = r;
  o.u = u;
  if (axis == 0) { // independent y, dependent z
    o.r.y = c * r.y - sn * r.z; o.r.z = sn * r.y + c * r.z;
    o.u.y = c * u.y - sn * u.z; o.u.z = sn * u.y + c * u.z;
  } else if (axis == 1) { // independent x, dependent z (the reference's convention for y)
    o.r.x = c * r.x - sn * r.z; o.r.z = sn * r.x + c * r.z;
    o.u.x = c * u.x - sn * u.z; o.u.z = sn * u.x + c * u.z;
  } else { // independent x, dependent y
    o.r.x = c * r.x - sn * r.y; o.r.y = sn * r.x + c * r.y;
    o.u.x = c * u.x - sn * u.y; o.u.y = sn * u.x + c * u.y;
  }
  return o;
}
// WhiteBC::handle_particle, src/boundary_condition.cpp:53-62 -> new direction, advanced seed
struct DirSeed {
  V3 u;
  uint64_t seed;
};
OMC_OOL DirSeed white_reflect(V3 r, V3 u, uint64_t seed, int s)
{
  DirSeed o;
  o.seed = seed;
  const V3 d = diffuse_reflect(surf_normal(s, r), u, o.seed);
  const double nrm = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
  o.u = v3(d.x / nrm, d.y / nrm, d.z / nrm);
  return o;
}

// Particle::event_cross_surface + cross_surface + BCs + cross_lattice:
// src/particle.cpp:288-329, 549-710 ; src/geometry.cpp:301-357 ;
// src/boundary_condition.cpp:18-47
OMC_D bool ev_surface(const PState& ps, const GenCtx& gc, int64_t i)
{
  ps.cell_last0[i] = LV(cell, 0);
  const int blt = ps.b_level_trans[i];
  const int n_coord = blt & 0xff;
  const int trans = blt >> 8;
  int surface = ps.b_surface[i];
  ps.surface[i] = surface;
  ps.n_coord[i] = n_coord;
  if (trans != TRANS_NONE) {
    // ---- cross_lattice ------------------------------------------------------
    const int lvl = n_coord - 1;
    const int lat = LV(lattice, lvl);
    LatIdx li;
    li.x = LV(lat_x, lvl) + ((trans & 3) - 1);
    li.y = LV(lat_y, lvl) + (((trans >> 2) & 3) - 1);
    li.z = LV(lat_z, lvl) + (((trans >> 4) & 3) - 1);
    LV(lat_x, lvl) = li.x;
    LV(lat_y, lvl) = li.y;
    LV(lat_z, lvl) = li.z;
    bool found = false;
    if (lat_is_hex(lat)) { // hexagonal lattice: out of line
      if (hex_cross(ps, i, lvl))
        found = find_cell_from(ps, i, n_coord, surface);
    } else {
      V3 r = get_r(ps, i, lvl - 1);
      const int c_up = LV(cell, lvl - 1);
      const double* t = DM.cell_translation + 3 * c_up;
      r.x -= gld(t);
      r.y -= gld(t + 1);
      r.z -= gld(t + 2);
      if (gld(DM.cell_rotated + c_up)) // src/geometry.cpp:324-326
        r = cell_rotate(c_up, r);
      set_r(ps, i, lvl, rect_local_position(lat, r, li));
      if (rect_valid(lat, li)) {
        LV(universe, lvl) = lat_universe(lat, li);
        found = find_cell_from(ps, i, n_coord, surface);
      }
    }
    if (!found) {
      ps.n_coord[i] = 1;
      found = find_cell_from(ps, i, 1, surface);
      if (!found)
        mark_lost(ps, gc, i);
    }
  } else {
    const int s = abs(surface) - 1;
    const int bc = gld(DM.surf_bc + s);
    if (bc == OMCGPU_BC_VACUUM) {
      ps.k_leak[i] += ps.wgt[i];
      ps.wgt[i] = 0.0;
    } else if (bc == OMCGPU_BC_PERIODIC) {
      // TranslationalPeriodicBC::handle_particle + Particle::cross_periodic_bc,
      // src/boundary_condition.cpp:129-155, src/particle.cpp:712-760
      if (n_coord != 1) {
        mark_lost(ps, gc, i);
      } else {
        const int axis = gld(DM.surf_periodic_axis + s);
        if (axis < 0) {
          const double* t = DM.surf_periodic_shift + 3 * s;
          LV(rx, 0) += gld(t);
          LV(ry, 0) += gld(t + 1);
          LV(rz, 0) += gld(t + 2);
        } else {
#if OMC_RARE_SURFACES
          const PosDir o = rotate_periodic(get_r(ps, i, 0), get_u(ps, i, 0), s, axis);
          LV(rx, 0) = o.r.x;
          LV(ry, 0) = o.r.y;
          LV(rz, 0) = o.r.z;
          set_u(ps, i, 0, o.u);
#endif
        }
        const int partner = gld(DM.surf_periodic + s) + 1;
        // translational: same side of the partner; rotational: the other side (both normals of a
        // rotational pair point into the geometry), src/boundary_condition.cpp:137-143, 252-258
        surface = (surface > 0) == (axis < 0) ? partner : -partner;
        ps.surface[i] = surface;
        if (!find_cell_from(ps, i, 1, surface))
          mark_lost(ps, gc, i);
      }
    } else if (bc == OMCGPU_BC_REFLECTIVE || bc == OMCGPU_BC_WHITE) {
      if (n_coord != 1) {
        mark_lost(ps, gc, i);
      } else {
#if OMC_RARE_SURFACES
        if (bc == OMCGPU_BC_WHITE) {
          const DirSeed o = white_reflect(get_r(ps, i, 0), get_u(ps, i, 0), ps.seed[i], s);
          ps.seed[i] = o.seed;
          set_u(ps, i, 0, o.u);
        } else
#endif
        {
          const V3 r = get_r(ps, i, 0);
          V3 u = get_u(ps, i, 0);
          V3 n = surf_normal(s, r); // Position::reflect, include/openmc/position.h:207-213
          const double projection = dot(n, u);
          const double magnitude = dot(n, n);
          const double f = 2.0 * projection / magnitude;
          u = v3(u.x - n.x * f, u.y - n.y * f, u.z - n.z * f);
          const double nrm = sqrt(u.x * u.x + u.y * u.y + u.z * u.z);
          u = v3(u.x / nrm, u.y / nrm, u.z / nrm);
          set_u(ps, i, 0, u);
        }
        surface = -surface;
        ps.surface[i] = surface;
        if (!find_cell_from(ps, i, 1, surface))
          mark_lost(ps, gc, i);
      }
    } else {
      bool found = find_cell_from(ps, i, n_coord, surface);
      if (!found) {
        ps.n_coord[i] = 1;
        found = find_cell_from(ps, i, 1, surface);
      }
      if (!found) {
        ps.surface[i] = SURFACE_NONE;
        ps.n_coord[i] = 1;
        LV(rx, 0) += TINY_BIT * LV(ux, 0);
        LV(ry, 0) += TINY_BIT * LV(uy, 0);
        LV(rz, 0) += TINY_BIT * LV(uz, 0);
        if (!find_cell_from(ps, i, 1, SURFACE_NONE))
          mark_lost(ps, gc, i);
      }
    }
  }
  return ev_revive(ps, gc, i);
}

// the replica of the tally VALUE arrays this block adds into (see k_score)
OMC_D double* score_values(const GenCtx& gc)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  return gc.tally_values + (int64_t)(blockIdx.x % gc.tally_replicas) * gc.tally_replica_stride;
#else
  return gc.tally_values;
#endif
}

// Tallies with the collision estimator, scored by the collision kernel before collision()
// (score_collision_tally, src/tallies/tally_scoring.cpp:2543-2620).  Out of line and fed from the
// particle's columns, not from ev_collide's registers: models without such tallies (the default
// estimator is tracklength) pay one uniform branch.
OMC_OOL void ev_collide_score(const PState& ps, const GenCtx& gc, int64_t i)
{
  ScoreCtx sc;
  const int n_coord = ps.n_coord[i];
  const int material = ps.material[i];
  sc.xs_total = ps.xs_total[i];
  sc.flux = ps.wgt[i] / sc.xs_total; // wgt_last / Sigma_t, src/tallies/tally_scoring.cpp:2546-2549
  sc.xs_absorption = ps.xs_absorption[i];
  sc.xs_fission = ps.xs_fission[i];
  sc.xs_nu_fission = ps.xs_nu_fission[i];
  sc.material = material;
  sc.g = ps.g[i];
  sc.mg_ba = 0;
  if (!DM.run_CE && material >= 0)
    sc.mg_ba = mg_ba(material, mg_temperature_index(material, ps.sqrtkT[i]),
      mg_angle_index(material, get_u(ps, i, n_coord - 1)));
  score_collision(ps, i, n_coord, get_r(ps, i, 0), ps.E[i], sc, score_values(gc));
}

// Particle::event_collide, src/particle.cpp:331-415
// SCORE: score the collision-estimator tallies here (the tail kernel); the list kernels leave
// that to k_collision_score, launched ahead of them only when the model has such tallies
template<bool BX = false, bool SCORE = true>
OMC_D bool ev_collide(const PState& ps, const GenCtx& gc, const Bank& fb, int64_t i)
{
  if (SCORE && gc.tallies_active && DM.n_collision_tallies > 0)
    ev_collide_score(ps, gc, i);
  Collider p;
  p.r = get_r(ps, i, 0);
  p.u = get_u(ps, i, 0);
  p.E = ps.E[i];
  p.wgt = ps.wgt[i];
  p.time = ps.time[i];
  p.g = ps.g[i];
  p.material = ps.material[i];
  p.seed = ps.seed[i];
  p.n_progeny = ps.n_progeny[i];
  p.id = gc.id_offset + ps.id_shift + i + 1;
  p.xs_total = ps.xs_total[i];
  p.xs_absorption = ps.xs_absorption[i];
  p.xs_nu_fission = ps.xs_nu_fission[i];
  p.k_abs = ps.k_abs[i];
  ps.k_col[i] += p.wgt * p.xs_nu_fission / p.xs_total;
  ps.surface[i] = SURFACE_NONE;
  // tallies scored after the collision (analog estimator): the state before it goes into the
  // particle's collision record, collision_ce / collision_mg add what happened
  const bool record = an_on(gc);
  const int64_t ridx = ps.id_shift + i;
  if (record) {
    an_put(gc, ridx, AN_RX, p.r.x);
    an_put(gc, ridx, AN_RY, p.r.y);
    an_put(gc, ridx, AN_RZ, p.r.z);
    an_put(gc, ridx, AN_E_LAST, p.E);
    an_put(gc, ridx, AN_WGT_LAST, p.wgt);
    an_put(gc, ridx, AN_WGT_BANK, 0.0);
    an_put(gc, ridx, AN_MICRO_T, 0.0);
    an_put(gc, ridx, AN_MICRO_A, 0.0);
    an_put(gc, ridx, AN_MICRO_F, 0.0);
    an_put(gc, ridx, AN_MICRO_NF, 0.0);
    an_put(gc, ridx, AN_XS_T, p.xs_total);
    an_put(gc, ridx, AN_XS_A, p.xs_absorption);
    an_put(gc, ridx, AN_XS_F, ps.xs_fission[i]);
    an_put(gc, ridx, AN_XS_NF, p.xs_nu_fission);
    an_puti(gc, ridx, AN_SCATTER, 0);
    an_puti(gc, ridx, AN_FISSION, 0);
    an_puti(gc, ridx, AN_G_LAST, p.g);
    an_puti(gc, ridx, AN_MATERIAL, p.material);
    an_puti(gc, ridx, AN_MG_BA, 0);
  }
  if (DM.run_CE) {
    bool no_nuclide;
    collision_ce<BX>(p, gc, fb, no_nuclide, ps.n_prefix ? ps.xs_prefix + (ps.id_shift + i) * ps.n_prefix : nullptr);
    if (no_nuclide)
      atomic_add_i32(gc.n_lost, 1);
    if (p.E < DM.energy_cutoff) // src/physics.cpp:67-71
      p.wgt = 0.0;
  } else {
    const int n_coord = ps.n_coord[i];
    const int ba = mg_ba(p.material, mg_temperature_index(p.material, ps.sqrtkT[i]),
      mg_angle_index(p.material, get_u(ps, i, n_coord - 1)));
    if (record)
      an_puti(gc, ridx, AN_MG_BA, ba);
    collision_mg(p, ba, gc, fb);
  }
  if (record) {
    an_puti(gc, ridx, AN_ALIVE, p.wgt != 0.0 ? 1 : 0);
    an_puti(gc, ridx, AN_G_POST, p.g);
    an_put(gc, ridx, AN_E_POST, p.E);
  }
  const int n_coord = ps.n_coord[i];
  { // directions of all levels follow level 0, through the rotations of rotated fills
    // (src/particle.cpp:396-406)
    V3 ul = p.u;
    set_u(ps, i, 0, ul);
    for (int l = 1; l < n_coord; ++l) {
      const int c_up = LV(cell, l - 1);
      if (gld(DM.cell_rotated + c_up))
        ul = cell_rotate(c_up, ul);
      set_u(ps, i, l, ul);
    }
  }
  ps.E[i] = p.E;
  ps.g[i] = p.g;
  ps.wgt[i] = p.wgt;
  ps.seed[i] = p.seed;
  ps.n_progeny[i] = p.n_progeny;
  ps.k_abs[i] = p.k_abs;
  ps.material_last[i] = C_NONE;
  if (SCORE && record)
    score_after_collision(ps, i, n_coord, gc, ridx, score_values(gc), fb.E);
  return ev_revive(ps, gc, i);
}

// The particle columns a list kernel receives are the GLOBAL ones (EvArgs::ps.is_local == 0; the tail
// kernel builds its thread-local views itself).  Telling the compiler lets it drop the
// thread-local branches of find_cell_from / hex_cross / hex_distance_at_level from those kernels.
#if OMC_CUDA && defined(__CUDA_ARCH__)
#define ASSUME_GLOBAL_COLUMNS(a) __builtin_assume((a).ps.is_local == 0)
#else
#define ASSUME_GLOBAL_COLUMNS(a) ((void)0)
#endif

// ---- kernels -----------------------------------------------------------------
OMC_D void body_init(int64_t tid, const EvArgs& a)
{
  ASSUME_GLOBAL_COLUMNS(a);
  if (tid >= a.gc.n)
    return;
  if (ev_init(a.ps, a.gc, a.src, tid))
    queue_append(a.q_a, (int32_t)tid);
}
OMC_KERNEL_GEO(k_init, EvArgs, body_init)

// Sort key of a calculate-xs work item: (material, lethargy) -- particles of one material
// with neighbouring energies sit next to each other in the sorted list, so a warp's
// gathers of the hash window, the probes and the xs rows of a nuclide fall into the same
// few sectors (many-nuclide materials: the tables are far larger than L2).  Positions
// beyond the list's count get the largest key.
struct KeyArgs {
  PState ps;
  Queue q_in;
  uint32_t* keys;
  int32_t* vals;
};
OMC_D void body_xs_keys(int64_t tid, const KeyArgs& a)
{
  uint32_t key = 0xffffffffu;
  int32_t i = 0;
  if (tid < *a.q_in.count) {
    i = list_item(a.q_in, tid);
    const int m = a.ps.material[i];
    if (m >= 0 && DM.run_CE) {
      double u = log(a.ps.E[i] / DM.energy_min) * DM.fine_inv_spacing * 8.0; // 1/8 of a fine-hash bin
      u = u < 0.0 ? 0.0 : (u > 16777215.0 ? 16777215.0 : u);
      key = ((uint32_t)(m < 254 ? m : 254) << 24) | (uint32_t)u;
    } else {
      key = 0xff000000u;
    }
  }
  a.keys[tid] = key;
  a.vals[tid] = i;
}
OMC_KERNEL(k_xs_keys, KeyArgs, body_xs_keys)

template<bool BX>
OMC_D void body_xs(int64_t tid, const EvArgs& a)
{
  ASSUME_GLOBAL_COLUMNS(a);
  if (tid >= *a.q_in.count)
    return;
  ev_xs<BX>(a.ps, list_item(a.q_in, tid), a.xs_batch_min, a.gc.count_lookups ? a.gc.counters + CNT_OTHER : nullptr);
}
OMC_KERNEL_MB(k_xs, EvArgs, body_xs<false>, 2)
OMC_KERNEL_MB(k_xs_b, EvArgs, body_xs<true>, 2)

// a work-list position that records no flight this sweep must say so: k_score
// walks positions, and the record left there by an earlier sweep is stale
OMC_D void no_flight(const EvArgs& a, int64_t tid)
{
  if (a.gc.tallies_active && DM.n_tallies > 0 && tid < a.fl.cap)
    a.fl.pidx[tid] = PIDX_NO_FLIGHT;
}

// What the list build reads of a history after its flight.  ordered == 1: 1 collision, 2 crossing,
// 0 dead (three-way partition in index order).  ordered == 2 (keyed lists): a class byte -- 0..31
// the material a collision happens in, 32..254 what a crossing will do (the surface token it
// crosses, or the level of a lattice step), 255 dead.  The lists are then stable-sorted by class
// (index order inside a class), so that the lanes of a warp sample the same nuclides, cross the
// same surface, meet the same boundary condition and search the same candidate cells.
constexpr int FLAG_DEAD_KEYED = 255, FLAG_FIRST_CROSSING = 32;
OMC_D uint8_t event_flag(int ordered, int next, const AdvanceOut& ao)
{
  if (ordered != 2)
    return next == NEXT_COLLISION ? 1 : (next == NEXT_SURFACE ? 2 : 0);
  if (next == NEXT_COLLISION) {
    const int m = ao.material;
    return (uint8_t)(m < 0 || m > 31 ? 31 : m);
  }
  if (next == NEXT_SURFACE) {
    const int blt = ao.b_level_trans;
    const int code = (blt >> 8) != TRANS_NONE ? (blt & 0xff) : 8 + ao.b_surface + DM.n_surfaces;
    return (uint8_t)(FLAG_FIRST_CROSSING + (code < 222 ? code : 222));
  }
  return FLAG_DEAD_KEYED;
}
OMC_D uint8_t dead_flag(int ordered)
{
  return ordered == 2 ? FLAG_DEAD_KEYED : 0;
}

// flight of particle i recorded at work-list position tid, then routed to the
// collision (q_a) or crossing (q_b) list
OMC_D void advance_and_route(int64_t tid, const EvArgs& a, int32_t i)
{
  AdvanceOut ao = {0, 0, 0};
  const int next = ev_advance(a.ps, a.gc, a.fl, i, tid, &ao);
  if (a.ordered) {
    a.ps.next_ev[i] = event_flag(a.ordered, next, ao);
    if (next == NEXT_DEAD)
      no_flight(a, tid);
    return;
  }
  if (next == NEXT_COLLISION)
    queue_append(a.q_a, i);
  else if (next == NEXT_SURFACE)
    queue_append(a.q_b, i);
  else
    no_flight(a, tid);
}

OMC_D void body_advance(int64_t tid, const EvArgs& a)
{
  ASSUME_GLOBAL_COLUMNS(a);
  if (tid >= *a.q_in.count)
    return;
  advance_and_route(tid, a, list_item(a.q_in, tid));
}
OMC_KERNEL_GEO(k_advance, EvArgs, body_advance)

// calculate-xs and advance events of the same work list in one pass: every
// alive particle goes through both in every sweep, so unless the xs list is to be
// re-ordered in between, one kernel saves a launch and one round trip of the
// particle's sectors through L2
template<bool BX>
OMC_D void body_xs_advance(int64_t tid, const EvArgs& a)
{
  ASSUME_GLOBAL_COLUMNS(a);
  if (tid >= *a.q_in.count)
    return;
  const int32_t i = list_item(a.q_in, tid);
  prefetch_particle(a.ps, i, false);
  ev_xs<BX>(a.ps, i, a.xs_batch_min, a.gc.count_lookups ? a.gc.counters + CNT_OTHER : nullptr);
  advance_and_route(tid, a, i);
}
OMC_KERNEL_MB_GEO(k_xs_advance, EvArgs, body_xs_advance<false>, XA_MINBLOCKS)
OMC_KERNEL_MB_GEO(k_xs_advance_b, EvArgs, body_xs_advance<true>, 2)

// Sweep mode 2 (run_sweeps_paired): a crossing or a collision is always followed,
// for the survivor, by a cross-section lookup and a flight.  One kernel per list
// does all three for its particles -- every lane of a warp walks the same event
// sequence, so nothing diverges that did not before -- and routes the survivor
// straight into the NEXT sweep's collision / crossing list.  A particle's
// sectors then travel through L2 once per sweep instead of twice, and the merged
// "alive" list (one append + one gather per particle and sweep) disappears.
template<bool BX>
OMC_D void body_surface_xa(int64_t tid, const EvArgs& a)
{
  ASSUME_GLOBAL_COLUMNS(a);
  if (tid >= *a.q_in.count)
    return;
  const int32_t i = list_item(a.q_in, tid);
  prefetch_particle(a.ps, i, false);
  if (ev_surface(a.ps, a.gc, i)) {
    ev_xs<BX>(a.ps, i, a.xs_batch_min, a.gc.count_lookups ? a.gc.counters + CNT_SURFACE : nullptr);
    advance_and_route(tid, a, i);
  } else {
    no_flight(a, tid);
    if (a.ordered)
      a.ps.next_ev[i] = dead_flag(a.ordered);
  }
}
OMC_KERNEL_MB_GEO(k_surface_xa, EvArgs, body_surface_xa<false>, XA_MINBLOCKS)
OMC_KERNEL_MB_GEO(k_surface_xa_b, EvArgs, body_surface_xa<true>, 2)
OMC_KERNEL_DYN(k_surface_xa_dyn, EvArgs, body_surface_xa<false>, XA_MINBLOCKS)

// collision-estimator tallies of the particles on a collision list, before the collision kernel
// changes them (launched only for models that have such tallies: the collision kernels themselves
// carry none of the scoring code)
OMC_D void body_collision_score(int64_t tid, const EvArgs& a)
{
  ASSUME_GLOBAL_COLUMNS(a);
  if (tid >= *a.q_in.count)
    return;
  ev_collide_score(a.ps, a.gc, list_item(a.q_in, tid));
}
OMC_KERNEL(k_collision_score, EvArgs, body_collision_score)
#define LAUNCH_COLLISION_SCORE(n_ub, stream, args) \
  do { \
    if (G.n_collision_tallies > 0 && (args).gc.tallies_active) \
      OMC_LAUNCH(k_collision_score, n_ub, stream, args); \
  } while (0)

// tallies scored AFTER the collision (analog estimator), from the records the collision kernel
// left: launched behind it on the same stream, only for models that have such tallies
OMC_D void body_after_collision_score(int64_t tid, const EvArgs& a)
{
  ASSUME_GLOBAL_COLUMNS(a);
  if (tid >= *a.q_in.count)
    return;
  const int32_t i = list_item(a.q_in, tid);
  score_after_collision(a.ps, i, a.ps.n_coord[i], a.gc, a.ps.id_shift + i, score_values(a.gc), a.fb.E);
}
OMC_KERNEL(k_after_collision_score, EvArgs, body_after_collision_score)
#define LAUNCH_AFTER_COLLISION_SCORE(n_ub, stream, args) \
  do { \
    if (G.n_after_collision_tallies > 0 && (args).gc.tallies_active) \
      OMC_LAUNCH(k_after_collision_score, n_ub, stream, args); \
  } while (0)

template<bool BX>
OMC_D void body_collide_xa(int64_t tid, const EvArgs& a)
{
  ASSUME_GLOBAL_COLUMNS(a);
  if (tid >= *a.q_in.count)
    return;
  const int32_t i = list_item(a.q_in, tid);
  prefetch_particle(a.ps, i, true);
  if (ev_collide<BX, false>(a.ps, a.gc, a.fb, i)) {
    ev_xs<BX>(a.ps, i, a.xs_batch_min, a.gc.count_lookups ? a.gc.counters + CNT_COLLIDE : nullptr);
    advance_and_route(tid, a, i);
  } else {
    no_flight(a, tid);
    if (a.ordered)
      a.ps.next_ev[i] = dead_flag(a.ordered);
  }
}
OMC_KERNEL_MB_GEO(k_collide_xa, EvArgs, body_collide_xa<false>, CX_MINBLOCKS)
OMC_KERNEL_MB_GEO(k_collide_xa_b, EvArgs, body_collide_xa<true>, 2)
OMC_KERNEL_DYN(k_collide_xa_dyn, EvArgs, body_collide_xa<false>, CX_MINBLOCKS)


// -----------------------------------------------------------------------------
// k_score -- tracklength tally scoring for the flights k_advance just recorded
// (score_tracklength_tally, src/tallies/tally_scoring.cpp:2416-2541).
//
// One lane evaluates the filters of one flight.  The mesh ray trace, whose
// length differs wildly from flight to flight, is NOT done lane-per-flight:
// every flight is cut into parts of about SCORE_PART mesh cells, a warp prefix
// sum turns the 32 flights of a warp into one list of parts, and the lanes
// trace the parts round-robin from a shared-memory copy of the flight data --
// so a warp's work is the sum of its flights' cells, not 32 x the longest one.
// -----------------------------------------------------------------------------
constexpr int SCORE_NVAL = OMCGPU_MAX_SCORES;
#ifndef SCORE_MINBLOCKS
#define SCORE_MINBLOCKS 3 /* resident blocks per SM asked of k_score (80 registers; measured r21: +3 % at 4e6 histories) */
#endif

// simple tally: one bin per non-mesh filter -> base index.  NS = scores actually emitted
// per mesh cell (the loop below is compile-time: no issue slots for absent scores)
template<int NS>
struct PartEmit {
  double* r0;     // results + base * ns: the part's row of bin 0
  int bin_stride; // doubles between two mesh bins = mesh_stride * ns
  double val[NS];
  OMC_D void set(double* results, int64_t base, int mesh_stride, int ns)
  {
    r0 = results + base * ns;
    bin_stride = mesh_stride * ns;
  }
  OMC_D void operator()(int bin, double w)
  {
    double* r = r0 + (int64_t)bin * bin_stride; // one wide multiply-add per mesh cell
#pragma unroll
    for (int s = 0; s < NS; ++s) // compile-time indices keep val in registers
      if (val[s] != 0.0)
        atomic_add_f64(r + s, val[s] * w);
  }
};

// part boundaries: identical expression for the end of part c and the start of c+1
OMC_HD double part_bound(int c, int nc, double total)
{
  return c >= nc ? total : (c * total) / nc;
}

// rare path kept out of line so that it does not cost the hot path registers
#if OMC_CUDA
__device__ __noinline__
#else
inline
#endif
void score_mesh_generic(const TallyEval& te, int t, const ScoreCtx& sc, double* values, V3 r0, V3 u,
  double total)
{
  GenericEmit em;
  em.te = &te;
  em.t = t;
  em.sc = &sc;
  em.values = values;
  mesh_march(te.mesh, r0, u, total, 0.0, total, em);
}

// one flight record in registers
struct Flight {
  V3 r0, u;
  double total, E;
  ScoreCtx sc;
  int32_t pidx;
};
OMC_D Flight load_flight(const Flights& fl, int64_t j)
{
  Flight f;
  f.r0 = v3(fl.x[j], fl.y[j], fl.z[j]);
  f.u = v3(fl.ux[j], fl.uy[j], fl.uz[j]);
  f.total = fl.total[j];
  f.E = fl.E[j];
  f.sc.flux = fl.flux[j];
  f.sc.xs_total = fl.xs_t[j];
  f.sc.xs_absorption = fl.xs_a[j];
  f.sc.xs_fission = fl.xs_f[j];
  f.sc.xs_nu_fission = fl.xs_nf[j];
  f.sc.material = fl.material[j];
  f.sc.g = fl.g[j];
  f.sc.mg_ba = fl.mg_ba[j];
  f.pidx = fl.pidx[j];
  return f;
}

// lane-side preparation of tally t for one flight: evaluates the filters,
// scores tallies without a mesh filter at once, and returns the number of parts
// the mesh ray trace is cut into (0: nothing to trace cooperatively)
OMC_D int score_prepare(const PState& ps, const Flight& f, int t, double* values, int score_part,
  int64_t& base, double* val, unsigned long long* red_counter = nullptr)
{
  if (ld(DM.tally_estimator + t) != OMCGPU_ESTIMATOR_TRACKLENGTH)
    return 0; // collision-estimator tallies are scored by the collision kernel
  TallyEval te;
  const int64_t i = f.pidx; // cell / universe filters read the particle's coordinate levels
  const int n_coord = f.pidx >= 0 ? ps.n_coord[i] : 0;
  if (!tally_prepare(ps, i, n_coord, t, f.E, f.sc, te))
    return 0;
  if (te.mesh_pos < 0) {
    tally_emit(te, t, 0, 1.0, f.sc, values);
    return 0;
  }
  if (f.total == 0.0)
    return 0;
  bool simple = true;
  base = 0;
  for (int k = 0; k < te.nf; ++k) {
    if (k == te.mesh_pos)
      continue;
    if (te.nb[k] != 1)
      simple = false;
    base += (int64_t)te.bins[k][0] * te.stride[k];
  }
  if (!simple) { // several matches per filter: this lane traces the whole flight
    score_mesh_generic(te, t, f.sc, values, f.r0, f.u, f.total);
    return 0;
  }
  const int ns = ld(DM.tally_n_scores + t);
  bool any = false;
#pragma unroll
  for (int s = 0; s < SCORE_NVAL; ++s) {
    val[s] = 0.0;
    if (s < ns) {
      bool skip;
      const double v = score_value(ld(DM.tally_scores + OMCGPU_MAX_SCORES * t + s), f.sc, skip);
      if (!skip)
        val[s] = v;
      any = any || val[s] != 0.0;
    }
  }
  if (!any)
    return 0;
  if (f.total < 2 * TINY_BIT)
    return 1;
  const V3 r1 = v3(f.r0.x + f.total * f.u.x, f.r0.y + f.total * f.u.y, f.r0.z + f.total * f.u.z);
  const int cells = mesh_cells_visited(te.mesh, f.r0, r1, f.u);
  if (red_counter) { // cells x scores that add something: the REDs this flight issues (exact inside the mesh)
    int nnz = 0;
    for (int s = 0; s < SCORE_NVAL; ++s)
      nnz += val[s] != 0.0 ? 1 : 0;
#if OMC_CUDA && defined(__CUDA_ARCH__)
    atomicAdd(red_counter, (unsigned long long)(cells * nnz));
#else
    *red_counter += (unsigned long long)(cells * nnz);
#endif
  }
  const int nc = (cells + score_part - 1) / score_part;
  return nc < 1 ? 1 : nc;
}

// The same preparation for the second version of k_score: the lane also computes the per-flight
// constants of the mesh ray trace (mesh_ray_setup: the only divisions of the trace), so that the
// parts can be started by any lane without a division.  Flights the DDA does not cover (they
// start outside the mesh, or are shorter than 2 TINY_BIT) are traced here by their own lane
// with the reference's loop.  Returns the number of parts; dpart = length of a part.
OMC_D int score_prepare2(const PState& ps, const Flight& f, int t, double* values, int score_part,
  int64_t& base, double* val, MeshRay& ray, double& dpart, unsigned long long* red_counter = nullptr)
{
  if (ld(DM.tally_estimator + t) != OMCGPU_ESTIMATOR_TRACKLENGTH)
    return 0; // collision-estimator tallies are scored by the collision kernel
  TallyEval te;
  const int64_t i = f.pidx; // cell / universe filters read the particle's coordinate levels
  const int n_coord = f.pidx >= 0 ? ps.n_coord[i] : 0;
  if (!tally_prepare(ps, i, n_coord, t, f.E, f.sc, te))
    return 0;
  if (te.mesh_pos < 0) {
    tally_emit(te, t, 0, 1.0, f.sc, values);
    return 0;
  }
  if (f.total == 0.0)
    return 0;
  bool simple = true;
  base = 0;
  for (int k = 0; k < te.nf; ++k) {
    if (k == te.mesh_pos)
      continue;
    if (te.nb[k] != 1)
      simple = false;
    base += (int64_t)te.bins[k][0] * te.stride[k];
  }
  if (!simple) { // several matches per filter: this lane traces the whole flight
    score_mesh_generic(te, t, f.sc, values, f.r0, f.u, f.total);
    return 0;
  }
  const int ns = ld(DM.tally_n_scores + t);
  bool any = false;
#pragma unroll
  for (int s = 0; s < SCORE_NVAL; ++s) {
    val[s] = 0.0;
    if (s < ns) {
      bool skip;
      const double v = score_value(ld(DM.tally_scores + OMCGPU_MAX_SCORES * t + s), f.sc, skip);
      if (!skip)
        val[s] = v;
      any = any || val[s] != 0.0;
    }
  }
  if (!any)
    return 0;
  if (f.total < 2 * TINY_BIT || !mesh_ray_setup(te.mesh, f.r0, f.u, ray)) {
    PartEmit<SCORE_NVAL> em;
#pragma unroll
    for (int s = 0; s < SCORE_NVAL; ++s)
      em.val[s] = val[s];
    em.set(values + ld(DM.tally_result_ofs + t), base, te.stride[te.mesh_pos], ns);
    mesh_march_exact(te.mesh, f.r0, f.u, f.total, 0.0, f.total, em);
    return 0;
  }
  const int cells = mesh_ray_cells(ray, f.total);
  if (red_counter) { // cells x scores that add something: the REDs this flight issues
    int nnz = 0;
    for (int s = 0; s < SCORE_NVAL; ++s)
      nnz += val[s] != 0.0 ? 1 : 0;
#if OMC_CUDA && defined(__CUDA_ARCH__)
    atomicAdd(red_counter, (unsigned long long)(cells * nnz));
#else
    *red_counter += (unsigned long long)(cells * nnz);
#endif
  }
  int nc = (cells + score_part - 1) / score_part;
  nc = nc < 1 ? 1 : nc;
  dpart = f.total / nc;
  return nc;
}

// part c of nc: [c dpart, (c + 1) dpart], the last one ends at total exactly
OMC_HD double part_begin(int c, double dpart)
{
  return c * dpart;
}
OMC_HD double part_end(int c, int nc, double dpart, double total)
{
  return c + 1 >= nc ? total : (c + 1) * dpart;
}

// position of the mesh filter and its stride are properties of the tally
OMC_HD void tally_mesh_info(int t, int& mesh, int& mesh_stride)
{
  mesh = -1;
  mesh_stride = 0;
  const int nf = ld(DM.tally_n_filters + t);
  for (int k = 0; k < nf; ++k) {
    const int f = ld(DM.tally_filters + OMCGPU_MAX_FILTERS * t + k);
    if (ld(DM.filter_type + f) == OMCGPU_FILTER_MESH) {
      mesh = ld(DM.filter_mesh + f);
      mesh_stride = ld(DM.tally_strides + OMCGPU_MAX_FILTERS * t + k);
    }
  }
}

#if OMC_CUDA
// Everything one lane does for its own flight and tally t, out of line: nothing
// of the flight stays in registers while the lanes trace parts afterwards.
// sd points at this lane's column of the warp's staging area (stride 32 doubles).

__constant__ Flights c_fl[4]; // [list kind][parity]: k_score(sweep s) overlaps the event kernels of sweep s+1

__device__ __noinline__ int score_lane(int64_t j, int buf, int t, double* values, int score_part,
  double* sd, int64_t* sbase, unsigned long long* red_counter)
{
  const Flight f = load_flight(c_fl[buf], j);
  double val[SCORE_NVAL];
  int64_t base = 0;
  const int nc = score_prepare(c_ps, f, t, values, score_part, base, val, red_counter);
  if (nc > 0) {
    sd[0 * 32] = f.r0.x;
    sd[1 * 32] = f.r0.y;
    sd[2 * 32] = f.r0.z;
    sd[3 * 32] = f.u.x;
    sd[4 * 32] = f.u.y;
    sd[5 * 32] = f.u.z;
    sd[6 * 32] = f.total;
#pragma unroll
    for (int s = 0; s < SCORE_NVAL; ++s)
      sd[(7 + s) * 32] = val[s];
    *sbase = base;
  }
  return nc;
}

// the lanes of a warp trace the parts of the warp's 32 flights round-robin
template<int NS>
__device__ __forceinline__ void score_parts(int mesh, int mesh_stride, int ns, double* results, int n_parts,
  int lane, const int* incl, const double (*sd)[32], const int64_t* sbase)
{
  PartEmit<NS> em;
  for (int item = lane; item < n_parts; item += 32) {
    int o = 0; // owner: first lane whose inclusive count exceeds item
    for (int step = 16; step > 0; step >>= 1)
      if (incl[o + step - 1] <= item)
        o += step;
    const int c = item - (o > 0 ? incl[o - 1] : 0);
    const int onc = incl[o] - (o > 0 ? incl[o - 1] : 0);
    const V3 pr0 = v3(sd[0][o], sd[1][o], sd[2][o]);
    const V3 pu = v3(sd[3][o], sd[4][o], sd[5][o]);
    const double ptotal = sd[6][o];
#pragma unroll
    for (int s = 0; s < NS; ++s)
      em.val[s] = sd[7 + s][o];
    em.set(results, sbase[o], mesh_stride, ns);
    mesh_march(mesh, pr0, pu, ptotal, part_bound(c, onc, ptotal), part_bound(c + 1, onc, ptotal), em);
  }
}

__global__ void __launch_bounds__(BLOCK, SCORE_MINBLOCKS) k_score(int64_t n_ub, EvArgs a)
{
  constexpr int WARPS = BLOCK / 32;
  __shared__ double s_d[WARPS][7 + SCORE_NVAL][32]; // r0, u, total, val[]
  __shared__ int64_t s_base[WARPS][32];
  __shared__ int s_incl[WARPS][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = *a.n_flights < a.fl.cap ? *a.n_flights : a.fl.cap;
  // Small tallies are replicated: REDs to one address serialise in the L2 slice that owns
  // it (measured, profiles/r13_red_peak.txt: 143 G RED/s into 30 000 doubles, 210 G/s into
  // >= 300 000), so every block adds into its own replica; omcgpu_tally_results folds them.
  double* values = a.gc.tally_values + (int64_t)(blockIdx.x % a.gc.tally_replicas) * a.gc.tally_replica_stride;
  unsigned long long* red_counter = a.gc.count_lookups ? a.gc.counters + CNT_RED : nullptr;
  // grid-stride over blocks of flights: with a capped grid (G.score_ctas) the kernel
  // occupies only part of every SM for the whole sweep, so that its RED traffic
  // shares the SM with the latency-bound event kernels of the next sweep
  // a warp's 32 flights at a time: either a static block-strided walk, or (a.cursor: persistent
  // grid) the next chunk fetched from a device cursor, so that no warp slot idles while another
  // warp of its block traces long flights
  int64_t wbase = (int64_t)blockIdx.x * BLOCK + 32 * warp;
  for (;; wbase += (int64_t)gridDim.x * BLOCK) {
  if (a.cursor) {
    int fetched = 0;
    if (lane == 0)
      fetched = atomicAdd(a.cursor, 32);
    wbase = __shfl_sync(0xffffffffu, fetched, 0);
  }
  if (wbase >= n || wbase >= n_ub)
    break;
  const int64_t tid = wbase + lane;
  const bool active = tid < n && a.fl.pidx[tid] != PIDX_NO_FLIGHT;
  for (int t = 0; t < DM.n_tallies; ++t) {
    __syncwarp(); // lanes leave the part loop of the previous tally at different times
    int nc = 0;
    if (active) // loads the flight, evaluates the filters, publishes r0/u/total/val/base
      nc = score_lane(tid, a.fl_buf, t, values, a.gc.score_part, &s_d[warp][0][lane], &s_base[warp][lane], red_counter);
    int mesh, mesh_stride;
    tally_mesh_info(t, mesh, mesh_stride);
    if (mesh < 0)
      continue; // warp-uniform: a property of the tally
    int incl = nc; // inclusive warp scan of the part counts
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d)
        incl += v;
    }
    const int n_parts = __shfl_sync(0xffffffffu, incl, 31);
    if (n_parts == 0)
      continue;
    s_incl[warp][lane] = incl;
    __syncwarp();
    const int ns = ld(DM.tally_n_scores + t);
    double* results = values + ld(DM.tally_result_ofs + t);
    // warp-uniform: a property of the tally
    if (ns == 1)
      score_parts<1>(mesh, mesh_stride, ns, results, n_parts, lane, s_incl[warp], s_d[warp], s_base[warp]);
    else if (ns <= 3)
      score_parts<3>(mesh, mesh_stride, ns, results, n_parts, lane, s_incl[warp], s_d[warp], s_base[warp]);
    else
      score_parts<SCORE_NVAL>(mesh, mesh_stride, ns, results, n_parts, lane, s_incl[warp], s_d[warp], s_base[warp]);
  }
  __syncwarp(); // the staging area is reused by the next block of flights
  }
}
// ---- second version: per-flight ray constants in the staging area, short parts, lanes in step ------
// staging area of one warp, one column per lane (= per flight)
constexpr int SD_TNEXT0 = 0, SD_TDEL = 3, SD_INV_TDEL = 6, SD_TOTAL = 9, SD_INV_TOTAL = 10, SD_DPART = 11,
              SD_VAL = 12, SD_N = SD_VAL + SCORE_NVAL;
constexpr int SI_LEFT = 0, SI_STRIDE = 3, SI_BIN0 = 6, SI_N = 7;
__device__ __noinline__ int score_lane2(int64_t j, int buf, int t, double* values, int score_part,
  double* sd, int* si, int64_t* sbase, unsigned long long* red_counter)
{
  const Flight f = load_flight(c_fl[buf], j);
  double val[SCORE_NVAL];
  int64_t base = 0;
  MeshRay ray;
  double dpart = 0.0;
  const int nc = score_prepare2(c_ps, f, t, values, score_part, base, val, ray, dpart, red_counter);
  if (nc > 0) {
#pragma unroll
    for (int k = 0; k < MESH_RAY_AXES; ++k) {
      sd[(SD_TNEXT0 + k) * 32] = ray.tnext0[k];
      sd[(SD_TDEL + k) * 32] = ray.tdel[k];
      sd[(SD_INV_TDEL + k) * 32] = ray.inv_tdel[k];
      si[(SI_LEFT + k) * 32] = ray.left[k];
      si[(SI_STRIDE + k) * 32] = ray.stride[k];
    }
    si[SI_BIN0 * 32] = ray.bin0;
    sd[SD_TOTAL * 32] = f.total;
    sd[SD_INV_TOTAL * 32] = 1.0 / f.total;
    sd[SD_DPART * 32] = dpart;
#pragma unroll
    for (int s = 0; s < SCORE_NVAL; ++s)
      sd[(SD_VAL + s) * 32] = val[s];
    *sbase = base;
  }
  return nc;
}

// the lanes of a warp trace the parts of the warp's 32 flights in rounds of 32 parts: every lane
// starts its part (multiplies only), then all lanes step through their few cells together
// (out of line, one function per variant: the hot loops get the kernel's whole register budget
// instead of sharing it with everything k_score2 keeps live -- inlined, they reloaded spilled
// values from local memory in every DDA step, profiles/r64)
constexpr int SCORE_WARPS = BLOCK / 32;
__shared__ double s2_d[SCORE_WARPS][SD_N][32];
__shared__ int s2_i[SCORE_WARPS][SI_N][32];
__shared__ int64_t s2_base[SCORE_WARPS][32];
__shared__ int s2_incl[SCORE_WARPS][32];
struct MeshRayStaged { // Ray view (mesh_part_march) of column o of the warp's staging area
  int warp, o;
  __device__ __forceinline__ double tnext0(int k) const { return s2_d[warp][SD_TNEXT0 + k][o]; }
  __device__ __forceinline__ double tdel(int k) const { return s2_d[warp][SD_TDEL + k][o]; }
  __device__ __forceinline__ double inv_tdel(int k) const { return s2_d[warp][SD_INV_TDEL + k][o]; }
  __device__ __forceinline__ int left(int k) const { return s2_i[warp][SI_LEFT + k][o]; }
  __device__ __forceinline__ int stride(int k) const { return s2_i[warp][SI_STRIDE + k][o]; }
  __device__ __forceinline__ int bin0() const { return s2_i[warp][SI_BIN0][o]; }
};
template<int NS, int ND>
__device__ __noinline__ void score_parts2(int mesh_stride, int ns, double* results, int n_parts)
{
  const int warp = threadIdx.x >> 5;
  PartEmit<NS> em;
  const int n_round = (n_parts + 31) & ~31;
  for (int item = threadIdx.x & 31; item < n_round; item += 32) {
    const bool have = item < n_parts;
    MeshRayStaged ray;
    ray.warp = warp;
    ray.o = 0;
    double ta = 0.0, tb = 0.0;
    if (have) {
      const int* incl = s2_incl[warp];
      int o = 0; // owner: first lane whose inclusive count exceeds item
      for (int step = 16; step > 0; step >>= 1)
        if (incl[o + step - 1] <= item)
          o += step;
      const int before = o > 0 ? incl[o - 1] : 0;
      const int c = item - before;
      const int onc = incl[o] - before;
      const double dpart = s2_d[warp][SD_DPART][o];
      ta = part_begin(c, dpart);
      tb = part_end(c, onc, dpart, s2_d[warp][SD_TOTAL][o]);
      const double inv_total = s2_d[warp][SD_INV_TOTAL][o];
#pragma unroll
      for (int s = 0; s < NS; ++s) // score per unit length: the march emits lengths
        em.val[s] = s2_d[warp][SD_VAL + s][o] * inv_total;
      em.set(results, s2_base[warp][o], mesh_stride, ns);
      ray.o = o;
    }
    __syncwarp();
    if (have)
      mesh_part_march<ND, true>(ray, ta, tb, 1.0, em);
    __syncwarp();
  }
}

__global__ void __launch_bounds__(BLOCK, SCORE_MINBLOCKS) k_score2(int64_t n_ub, EvArgs a)
{
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = *a.n_flights < a.fl.cap ? *a.n_flights : a.fl.cap;
  double* values = a.gc.tally_values + (int64_t)(blockIdx.x % a.gc.tally_replicas) * a.gc.tally_replica_stride;
  unsigned long long* red_counter = a.gc.count_lookups ? a.gc.counters + CNT_RED : nullptr;
  int64_t wbase = (int64_t)blockIdx.x * BLOCK + 32 * warp;
  for (;; wbase += (int64_t)gridDim.x * BLOCK) {
    if (a.cursor) {
      int fetched = 0;
      if (lane == 0)
        fetched = atomicAdd(a.cursor, 32);
      wbase = __shfl_sync(0xffffffffu, fetched, 0);
    }
    if (wbase >= n || wbase >= n_ub)
      break;
    const int64_t tid = wbase + lane;
    const bool active = tid < n && a.fl.pidx[tid] != PIDX_NO_FLIGHT;
    for (int t = 0; t < DM.n_tallies; ++t) {
      __syncwarp();
      int nc = 0;
      if (active)
        nc = score_lane2(tid, a.fl_buf, t, values, a.gc.score_part, &s2_d[warp][0][lane], &s2_i[warp][0][lane],
          &s2_base[warp][lane], red_counter);
      int mesh, mesh_stride;
      tally_mesh_info(t, mesh, mesh_stride);
      if (mesh < 0)
        continue; // warp-uniform: a property of the tally
      int incl = nc; // inclusive warp scan of the part counts
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d)
          incl += v;
      }
      const int n_parts = __shfl_sync(0xffffffffu, incl, 31);
      if (n_parts == 0)
        continue;
      s2_incl[warp][lane] = incl;
      __syncwarp();
      const int ns = ld(DM.tally_n_scores + t);
      const bool d2 = ld(DM.mesh_ndim + mesh) == 2; // 1-D meshes take the 3-axis loop (absent axes are never crossed)
      double* results = values + ld(DM.tally_result_ofs + t);
#define SCORE_PARTS2(NS_) \
  do { \
    if (d2) \
      score_parts2<NS_, 2>(mesh_stride, ns, results, n_parts); \
    else \
      score_parts2<NS_, 3>(mesh_stride, ns, results, n_parts); \
  } while (0)
      if (ns == 1)
        SCORE_PARTS2(1);
      else if (ns <= 3)
        SCORE_PARTS2(3);
      else
        SCORE_PARTS2(SCORE_NVAL);
#undef SCORE_PARTS2
    }
    __syncwarp(); // the staging area is reused by the next block of flights
  }
}

#ifndef SCORE_PART_V2
#define SCORE_PART_V2 6 /* mesh cells per part, k_score2 */
#endif
static int g_score_grid_cap = 0; // 0: one block per BLOCK flights; else at most this many blocks
static int g_score_version = 1;  // OMCGPU_SCORE_VERSION=2: k_score2 (measured slower, profiles/README.md r63-r65)
#define LAUNCH_SCORE(n_ub, stream, args) \
  do { \
    if ((n_ub) > 0) { \
      unsigned grid_ = (unsigned)(((n_ub) + BLOCK - 1) / BLOCK); \
      if (g_score_grid_cap > 0 && grid_ > (unsigned)g_score_grid_cap) \
        grid_ = (unsigned)g_score_grid_cap; \
      if ((args).cursor && grid_ > (unsigned)(dyn_grid_sms() * SCORE_MINBLOCKS)) \
        grid_ = (unsigned)(dyn_grid_sms() * SCORE_MINBLOCKS); \
      if (g_score_version == 1) \
        k_score<<<grid_, BLOCK, 0, stream>>>((int64_t)(n_ub), args); \
      else \
        k_score2<<<grid_, BLOCK, 0, stream>>>((int64_t)(n_ub), args); \
      ++g_launches; \
    } \
  } while (0)
#else
// host compile (tests only): same preparation, same parts, traced one after another
#ifndef SCORE_PART_V2
#define SCORE_PART_V2 6
#endif
static int g_score_version_host = 1;
#define g_score_version g_score_version_host
void k_score(int64_t n_ub, EvArgs a)
{
  const int64_t n = *a.n_flights < a.fl.cap ? *a.n_flights : a.fl.cap;
  for (int64_t tid = 0; tid < n_ub && tid < n; ++tid) {
    if (a.fl.pidx[tid] == PIDX_NO_FLIGHT)
      continue;
    const Flight f = load_flight(a.fl, tid);
    for (int t = 0; t < DM.n_tallies; ++t) {
      PartEmit<SCORE_NVAL> em;
      int64_t base = 0;
      int mesh, mesh_stride;
      tally_mesh_info(t, mesh, mesh_stride);
      if (g_score_version_host == 1) {
        const int nc = score_prepare(a.ps, f, t, a.gc.tally_values, a.gc.score_part, base, em.val);
        em.set(a.gc.tally_values + DM.tally_result_ofs[t], base, mesh_stride, DM.tally_n_scores[t]);
        for (int c = 0; c < nc; ++c)
          mesh_march(mesh, f.r0, f.u, f.total, part_bound(c, nc, f.total), part_bound(c + 1, nc, f.total), em);
        continue;
      }
      // the arithmetic of k_score2: ray constants once per flight, parts started without a division
      MeshRay ray;
      double dpart = 0.0;
      const int nc = score_prepare2(a.ps, f, t, a.gc.tally_values, a.gc.score_part, base, em.val, ray, dpart);
      if (nc == 0)
        continue;
      em.set(a.gc.tally_values + DM.tally_result_ofs[t], base, mesh_stride, DM.tally_n_scores[t]);
      MeshRayRegs view;
      view.r = &ray;
      const double inv_total = 1.0 / f.total;
      for (int c = 0; c < nc; ++c) {
        if (DM.mesh_ndim[mesh] == 2)
          mesh_part_march<2>(view, part_begin(c, dpart), part_end(c, nc, dpart, f.total), inv_total, em);
        else
          mesh_part_march<3>(view, part_begin(c, dpart), part_end(c, nc, dpart, f.total), inv_total, em);
      }
    }
  }
}
#define LAUNCH_SCORE(n_ub, stream, args) \
  do { \
    if ((n_ub) > 0) { \
      k_score((int64_t)(n_ub), args); \
      ++g_launches; \
    } \
  } while (0)
#endif

OMC_D void body_surface(int64_t tid, const EvArgs& a)
{
  ASSUME_GLOBAL_COLUMNS(a);
  if (tid >= *a.q_in.count)
    return;
  const int32_t i = list_item(a.q_in, tid);
  prefetch_particle(a.ps, i, false);
  const bool alive = ev_surface(a.ps, a.gc, i);
  if (a.ordered)
    a.ps.next_ev[i] = alive ? 3 : 0; // 3: waits for its next flight
  else if (alive)
    queue_append(a.q_a, i);
}
OMC_KERNEL_GEO(k_surface, EvArgs, body_surface)

template<bool BX>
OMC_D void body_collide(int64_t tid, const EvArgs& a)
{
  ASSUME_GLOBAL_COLUMNS(a);
  if (tid >= *a.q_in.count)
    return;
  const int32_t i = list_item(a.q_in, tid);
  prefetch_particle(a.ps, i, true);
  const bool alive = ev_collide<BX, false>(a.ps, a.gc, a.fb, i);
  if (a.ordered)
    a.ps.next_ev[i] = alive ? 3 : 0;
  else if (alive)
    queue_append(a.q_a, i);
}
OMC_KERNEL_MB_GEO(k_collide, EvArgs, body_collide<false>, CX_MINBLOCKS)
OMC_KERNEL_MB_GEO(k_collide_b, EvArgs, body_collide<true>, 2)

// The tail of a generation (few particles left, launch-latency bound): one
// thread runs the remaining events of one history to completion with the very
// same event functions (cf. transport_history_based_single_particle,
// src/simulation.cpp:804-821).
OMC_D void body_tail(int64_t tid, const EvArgs& a)
{
  if (tid >= *a.q_in.count)
    return;
  const int32_t i_global = list_item(a.q_in, tid);
  LocalParticle<OMCGPU_MAX_COORD> lp;
  load_local(a.ps, i_global, lp);
  const PState ps = local_view(lp, a.ps, i_global);
  const int64_t i = 0;
  unsigned long long n_xs = 0, n_coll = 0, n_cross = 0;
  int next = a.first_event; // what the histories of this list are waiting for
  for (;;) {
    if (next == NEXT_NONE) {
      ev_xs<true>(ps, i, a.xs_batch_min, a.gc.count_lookups ? a.gc.counters + CNT_OTHER : nullptr); // few long-running threads: chain length, not registers, matters
      ++n_xs;
      next = ev_advance<true>(ps, a.gc, a.fl, i, SLOT_LOG);
    }
    bool alive;
    if (next == NEXT_COLLISION) {
      alive = ev_collide<true>(ps, a.gc, a.fb, i);
      ++n_coll;
    } else if (next == NEXT_SURFACE) {
      alive = ev_surface(ps, a.gc, i);
      ++n_cross;
    } else {
      break;
    }
    if (!alive)
      break;
    next = NEXT_NONE;
  }
  store_local_results(a.ps, i_global, lp);
  atomic_add_u64(a.gc.counters + 1, n_xs);
  atomic_add_u64(a.gc.counters + 2, n_coll);
  atomic_add_u64(a.gc.counters + 3, n_cross);
}
OMC_KERNEL_GEO(k_tail, EvArgs, body_tail)
// few, long-running threads: small blocks spread them over all SMs
constexpr int TAIL_BLOCK = 64;
#if OMC_CUDA
#define LAUNCH_TAIL(n_ub, stream, args) \
  do { \
    (args).xs_batch_min = tail_batch_min(); \
    if ((n_ub) > 0) { \
      k_tail<<<(unsigned)(((n_ub) + TAIL_BLOCK - 1) / TAIL_BLOCK), TAIL_BLOCK, geo_smem_bytes(), stream>>>((int64_t)(n_ub), \
        args); \
      ++g_launches; \
    } \
  } while (0)
#else
#define LAUNCH_TAIL(n_ub, stream, args) \
  do { \
    (args).xs_batch_min = tail_batch_min(); \
    OMC_LAUNCH(k_tail, n_ub, stream, args); \
  } while (0)
#endif

// ---- keyed lists: where the collision classes end and the dead begin in the sorted flags ----
struct BoundsArgs {
  const uint8_t* keys; // sorted ascending
  int64_t n;
  int32_t* counts; // [0] collisions, [1] crossings
};
OMC_D int64_t first_not_below(const uint8_t* k, int64_t n, int v)
{
  int64_t lo = 0, len = n;
  while (len > 0) {
    const int64_t half = len >> 1;
    if (k[lo + half] < v) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  return lo;
}
OMC_D void body_bounds(int64_t tid, const BoundsArgs& a)
{
  if (tid == 0) {
    const int64_t c = first_not_below(a.keys, a.n, FLAG_FIRST_CROSSING);
    a.counts[0] = (int32_t)c;
  } else if (tid == 1) {
    const int64_t c = first_not_below(a.keys, a.n, FLAG_FIRST_CROSSING);
    const int64_t d = first_not_below(a.keys, a.n, FLAG_DEAD_KEYED);
    a.counts[1] = (int32_t)(d - c);
  }
}
OMC_KERNEL(k_bounds, BoundsArgs, body_bounds)

struct IotaArgs {
  int32_t* v;
  int64_t n;
};
OMC_D void body_iota(int64_t tid, const IotaArgs& a)
{
  if (tid < a.n)
    a.v[tid] = (int32_t)tid;
}
OMC_KERNEL(k_iota, IotaArgs, body_iota)

// ---- deterministic reduction of the per-particle accumulators ------------------
// (event_death, src/particle.cpp:475-512: global_tally_* += ..., total_weight)
struct ReduceArgs {
  PState ps;
  Bank src;
  int64_t n;
  double* partial; // [n_chunks][6]
  int64_t chunk;
};
OMC_D void body_reduce(int64_t tid, const ReduceArgs& a)
{
  const int64_t lo = tid * a.chunk;
  int64_t hi = lo + a.chunk;
  if (hi > a.n)
    hi = a.n;
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = lo; i < hi; ++i) {
    s[OMCGPU_K_COLLISION] += a.ps.k_col[i];
    s[OMCGPU_K_ABSORPTION] += a.ps.k_abs[i];
    s[OMCGPU_K_TRACKLENGTH] += a.ps.k_trk[i];
    s[OMCGPU_LEAKAGE] += a.ps.k_leak[i];
    s[4] += a.src.wgt[i];
    s[5] += (double)a.ps.n_event[i];
  }
  for (int k = 0; k < 6; ++k)
    a.partial[tid * 6 + k] = s[k];
}
OMC_KERNEL(k_reduce, ReduceArgs, body_reduce)

// ---- Mesh::count_sites over a bank (Shannon entropy), src/mesh.cpp:941-997 -----------------------
struct CountArgs {
  Bank bank;
  int64_t n;
  int32_t nd, shape[3];
  double ll[3], w[3];
  double* counts;
  int32_t* outside;
};
OMC_D void body_count_sites(int64_t tid, const CountArgs& a)
{
  if (tid >= a.n)
    return;
  const double r[3] = {a.bank.x[tid], a.bank.y[tid], a.bank.z[tid]};
  int bin = 0, cells = 1;
  for (int k = 0; k < a.nd; ++k) { // StructuredMesh::get_bin :874-929
    const int idx = mesh_index(r[k], a.ll[k], a.w[k]);
    if (idx < 1 || idx > a.shape[k]) {
      *a.outside = 1;
      return;
    }
    bin += (idx - 1) * cells;
    cells *= a.shape[k];
  }
  atomic_add_f64(a.counts + bin, a.bank.wgt[tid]);
}
OMC_KERNEL(k_count_sites, CountArgs, body_count_sites)

// ---- fold the replicas of the tally VALUE arrays into replica 0 --------------------
struct FoldArgs {
  double* v;
  int64_t n, stride;
  int32_t replicas;
};
OMC_D void body_fold_tally(int64_t tid, const FoldArgs& a)
{
  if (tid >= a.n)
    return;
  double s = a.v[tid];
  for (int r = 1; r < a.replicas; ++r) {
    s += a.v[r * a.stride + tid];
    a.v[r * a.stride + tid] = 0.0;
  }
  a.v[tid] = s;
}
OMC_KERNEL(k_fold_tally, FoldArgs, body_fold_tally)

// ---- bank kernels --------------------------------------------------------------
struct BankArgs {
  Bank a, b;
  omcgpu_site* aos;
  const int32_t* offs;    // exclusive scan of n_progeny
  const int32_t* gather;  // comb indices
  int64_t n;
  int64_t id_offset;
};
OMC_D void bank_copy(const Bank& d, int64_t j, const Bank& s, int64_t k)
{
  d.x[j] = s.x[k];
  d.y[j] = s.y[k];
  d.z[j] = s.z[k];
  d.u[j] = s.u[k];
  d.v[j] = s.v[k];
  d.w[j] = s.w[k];
  d.E[j] = s.E[k];
  d.time[j] = s.time[k];
  d.wgt[j] = s.wgt[k];
  d.delayed_group[j] = s.delayed_group[k];
  d.parent_id[j] = s.parent_id[k];
  d.progeny_id[j] = s.progeny_id[k];
}
OMC_D void body_scatter_bank(int64_t tid, const BankArgs& a)
{ // sort_fission_bank, src/bank.cpp:111-125
  if (tid >= a.n)
    return;
  const int64_t parent = a.a.parent_id[tid] - 1 - a.id_offset;
  const int64_t dest = (int64_t)a.offs[parent] + a.a.progeny_id[tid];
  bank_copy(a.b, dest, a.a, tid);
}
OMC_KERNEL(k_scatter_bank, BankArgs, body_scatter_bank)

OMC_D void body_gather_bank(int64_t tid, const BankArgs& a)
{ // temp_sites[index_temp] = fission_bank[idx], src/eigenvalue.cpp:166-177
  if (tid >= a.n)
    return;
  bank_copy(a.b, tid, a.a, a.gather[tid]);
}
OMC_KERNEL(k_gather_bank, BankArgs, body_gather_bank)

OMC_D void body_unpack(int64_t tid, const BankArgs& a)
{
  if (tid >= a.n)
    return;
  const omcgpu_site s = a.aos[tid];
  a.a.x[tid] = s.r[0];
  a.a.y[tid] = s.r[1];
  a.a.z[tid] = s.r[2];
  a.a.u[tid] = s.u[0];
  a.a.v[tid] = s.u[1];
  a.a.w[tid] = s.u[2];
  a.a.E[tid] = s.E;
  a.a.time[tid] = s.time;
  a.a.wgt[tid] = s.wgt;
  a.a.delayed_group[tid] = s.delayed_group;
  a.a.parent_id[tid] = s.parent_id;
  a.a.progeny_id[tid] = s.progeny_id;
}
OMC_KERNEL(k_unpack, BankArgs, body_unpack)

OMC_D void body_pack(int64_t tid, const BankArgs& a)
{
  if (tid >= a.n)
    return;
  omcgpu_site s;
  s.r[0] = a.a.x[tid];
  s.r[1] = a.a.y[tid];
  s.r[2] = a.a.z[tid];
  s.u[0] = a.a.u[tid];
  s.u[1] = a.a.v[tid];
  s.u[2] = a.a.w[tid];
  s.E = a.a.E[tid];
  s.time = a.a.time[tid];
  s.wgt = a.a.wgt[tid];
  s.delayed_group = a.a.delayed_group[tid];
  s.surf_id = 0;
  s.particle = 0;
  s.parent_nuclide = -1;
  s.parent_id = a.a.parent_id[tid];
  s.progeny_id = a.a.progeny_id[tid];
  a.aos[tid] = s;
}
OMC_KERNEL(k_pack, BankArgs, body_pack)

// ---- stand-alone probes ---------------------------------------------------------
struct XsArgs {
  const double* E;
  const int32_t* material;
  int64_t n;
  double* out;
  int32_t* i_grid;
  int32_t max_nuc;
  int32_t xs_batch_min;
};
template<bool BX>
OMC_D void body_probe_xs(int64_t tid, const XsArgs& a)
{
  if (tid >= a.n)
    return;
  const int m = a.material[tid];
  double t = 0, ab = 0, f = 0, nf = 0;
  if (DM.run_CE) {
    ce_macro_xs<BX>(m, a.E[tid], t, ab, f, nf, a.i_grid ? a.i_grid + tid * a.max_nuc : nullptr,
      a.max_nuc, a.xs_batch_min);
  } else {
    const int G = DM.mg_n_groups;
    const int g = (int)a.E[tid];
    const int ba = mg_ba(m, mg_temperature_index(m, gld(DM.cell_sqrtkT)), 0);
    t = ld(DM.mg_total + ba * G + g);
    ab = ld(DM.mg_absorption + ba * G + g);
    nf = ld(DM.mgm_fissionable + m) ? ld(DM.mg_nu_fission + ba * G + g) : 0.;
  }
#if OMC_CUDA && defined(__CUDA_ARCH__)
  reinterpret_cast<double2*>(a.out)[2 * tid] = make_double2(t, ab);
  reinterpret_cast<double2*>(a.out)[2 * tid + 1] = make_double2(f, nf);
#else
  a.out[4 * tid + 0] = t;
  a.out[4 * tid + 1] = ab;
  a.out[4 * tid + 2] = f;
  a.out[4 * tid + 3] = nf;
#endif
}
OMC_KERNEL_MB(k_probe_xs, XsArgs, body_probe_xs<false>, 2)
OMC_KERNEL_MB(k_probe_xs_b, XsArgs, body_probe_xs<true>, 2)

// collision() of src/physics.cpp:43-168 (CE) / src/physics_mg.cpp:26-60 (MG) on n independent
// neutrons -- the device function the collision kernels call, exposed on its own so that it can
// be held to vectors written by the reference's collision() (tests/golden/refcoll_ce.npz)
struct CollideArgs {
  PState ps;
  GenCtx gc;
  Bank fb;
  int64_t n;
  const int32_t* material;
  const double* E;
  const double* uvw;
  const uint64_t* seed;
  double wgt;
  double* E_out;
  double* uvw_out;
  double* wgt_out;
  double* k_abs_out;
  uint64_t* seed_out;
};
OMC_D void body_probe_collide(int64_t tid, const CollideArgs& a)
{
  if (tid >= a.n)
    return;
  Collider p;
  p.r = v3(0., 0., 0.);
  p.u = v3(a.uvw[3 * tid], a.uvw[3 * tid + 1], a.uvw[3 * tid + 2]);
  p.E = a.E[tid];
  p.wgt = a.wgt;
  p.time = 0.0;
  p.k_abs = 0.0;
  p.g = 0;
  p.material = a.material[tid];
  p.seed = a.seed[tid];
  p.n_progeny = 0;
  p.id = tid + 1;
  if (DM.run_CE) {
    double t, ab, f, nf;
    ce_macro_xs<false>(p.material, p.E, t, ab, f, nf, nullptr, 0);
    p.xs_total = t;
    p.xs_absorption = ab;
    p.xs_nu_fission = nf;
    bool no_nuclide = false;
    collision_ce<false>(p, a.gc, a.fb, no_nuclide);
    if (no_nuclide)
      atomic_add_i32(a.gc.n_lost, 1);
  } else {
    const int G_ = DM.mg_n_groups;
    p.g = (int)a.E[tid];
    const int ba = mg_ba(p.material, 0, mg_angle_index(p.material, p.u));
    p.xs_total = ld(DM.mg_total + ba * G_ + p.g);
    p.xs_absorption = ld(DM.mg_absorption + ba * G_ + p.g);
    p.xs_nu_fission = ld(DM.mgm_fissionable + p.material) ? ld(DM.mg_nu_fission + ba * G_ + p.g) : 0.;
    collision_mg(p, ba, a.gc, a.fb);
  }
  a.E_out[tid] = DM.run_CE ? p.E : (double)p.g;
  a.uvw_out[3 * tid] = p.u.x;
  a.uvw_out[3 * tid + 1] = p.u.y;
  a.uvw_out[3 * tid + 2] = p.u.z;
  a.wgt_out[tid] = p.wgt;
  a.k_abs_out[tid] = p.k_abs;
  a.seed_out[tid] = p.seed;
  a.ps.n_progeny[tid] = p.n_progeny;
}
OMC_KERNEL(k_probe_collide, CollideArgs, body_probe_collide)

struct GeoArgs {
  PState ps;
  const double* xyz;
  const double* uvw;
  int64_t n;
  int32_t *cell, *instance, *material, *n_coord;
  double* distance;
  int32_t *surface, *level, *lat_trans;
};
OMC_D void body_probe_geom(int64_t tid, const GeoArgs& a)
{
  if (tid >= a.n)
    return;
  const PState& ps = a.ps;
  const int64_t i = tid;
  const int L = DM.n_coord_levels < OMCGPU_MAX_COORD ? DM.n_coord_levels : OMCGPU_MAX_COORD;
  for (int l = 0; l < L; ++l)
    reset_level(ps, i, l);
  set_r(ps, i, 0, v3(a.xyz[3 * tid], a.xyz[3 * tid + 1], a.xyz[3 * tid + 2]));
  if (a.uvw)
    set_u(ps, i, 0, v3(a.uvw[3 * tid], a.uvw[3 * tid + 1], a.uvw[3 * tid + 2]));
  else
    set_u(ps, i, 0, v3(0.0, 0.0, 1.0));
  ps.material[i] = C_NONE;
  ps.sqrtkT[i] = -1.0;
  ps.cell_instance[i] = 0;
  const bool found = find_cell_from_root(ps, i, SURFACE_NONE);
  const int nc = ps.n_coord[i];
  if (a.cell) {
    a.cell[tid] = found ? LV(cell, nc - 1) : -1;
    a.instance[tid] = found ? ps.cell_instance[i] : -1;
    a.material[tid] = found ? ps.material[i] : -2;
    a.n_coord[tid] = found ? nc : 0;
  }
  if (a.distance) {
    if (!found) {
      a.distance[tid] = -1.0;
      a.surface[tid] = 0;
      a.level[tid] = 0;
      a.lat_trans[3 * tid] = a.lat_trans[3 * tid + 1] = a.lat_trans[3 * tid + 2] = 0;
    } else {
      bool neg;
      const Boundary b = distance_to_boundary(ps, i, nc, SURFACE_NONE, neg);
      a.distance[tid] = b.distance;
      a.surface[tid] = b.surface;
      a.level[tid] = b.level;
      a.lat_trans[3 * tid] = (b.trans & 3) - 1;
      a.lat_trans[3 * tid + 1] = ((b.trans >> 2) & 3) - 1;
      a.lat_trans[3 * tid + 2] = ((b.trans >> 4) & 3) - 1;
    }
  }
}
OMC_KERNEL_GEO(k_probe_geom, GeoArgs, body_probe_geom)

struct RngArgs {
  const uint64_t* n;
  const uint64_t* seed;
  int64_t count;
  int32_t per;
  uint64_t* out_u;
  double* out_d;
};
OMC_D void body_probe_rng(int64_t tid, const RngArgs& a)
{
  if (tid >= a.count)
    return;
  if (a.out_u) {
    a.out_u[tid] = future_seed(a.n[tid], a.seed[tid]);
  } else {
    uint64_t s = a.seed[tid];
    for (int j = 0; j < a.per; ++j)
      a.out_d[tid * a.per + j] = prn(s);
  }
}
OMC_KERNEL(k_probe_rng, RngArgs, body_probe_rng)

// =============================================================================
// host side
// =============================================================================
static char g_err[512] = "";

static int fail(int code, const char* fmt, const char* detail = "")
{
  std::snprintf(g_err, sizeof(g_err), fmt, detail);
  return code;
}

#define CUDA_OK(call) \
  do { \
    cudaError_t e_ = (call); \
    if (e_ != cudaSuccess) \
      return fail(OMCGPU_E_CUDA, "CUDA error: %s", cudaGetErrorString(e_)); \
  } while (0)

constexpr int RING = 32;              // sweeps whose counters stay readable
constexpr int RING_OFS = 8;           // d_counts: [0] alive after k_init, [4] fission sites, [5] lost, [6] tail flight log, [7] bank-full hits
constexpr int SLOT = 8;               // ints per ring slot: [0] collisions [1] crossings [2] next alive (merged sweep)
                                      // [4] [5] work cursors of the crossing / collision kernel, [6] [7] of their scoring launches
constexpr int N_COUNTS = RING_OFS + SLOT * RING;

struct Engine {
  bool ready = false;
  omcgpu_model hm;            // scalars + DEVICE pointers
  std::vector<void*> dev_allocs;
  int64_t cap = 0;            // max particles per generation
  int64_t fis_cap = 0;        // fission bank capacity (3 * cap)
  PState ps;
  Flights fl[4];              // [0],[1]: double buffer of sweep mode 1; mode 2 uses [kind * 2 + parity]
  bool tally_needs_cells = false;
  bool overlap = true;        // run k_score / k_surface / k_collide on their own streams
  cudaStream_t s_score = 0, s_surf = 0, s_coll = 0;
  cudaEvent_t e_adv = nullptr, e_surf = nullptr, e_coll = nullptr, e_score[2] = {nullptr, nullptr};
  Bank src, fis, fis_sorted;
  int64_t n_src = 0, n_fis = 0;
  Queue q[4];
  Queue q_alive;              // sweep mode 2: histories alive after k_init
  int32_t* d_counts = nullptr;   // [0..3] queue counts, [4] n_fission, [5] n_lost
  unsigned long long* d_counters = nullptr; // [8]: events, xs, collisions, crossings (tail), lookups done, nuclide lookups
  bool count_lookups = false;
  int32_t* h_counts = nullptr;   // pinned
  unsigned long long* h_counters = nullptr;
  int32_t* d_offs = nullptr;     // progeny scan
  int32_t* d_gather = nullptr;
  void* d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  double* d_partial = nullptr;
  double* h_partial = nullptr;
  int64_t n_chunks = 0;
  omcgpu_site* d_aos = nullptr;
  omcgpu_site* h_aos = nullptr;  // pinned staging
  double* d_tally = nullptr;
  int64_t n_tally = 0;
  std::vector<int64_t> tally_ofs; // host copy of tally_result_ofs
  int n_cell_sqrtkT = 0, n_mat_nuclides = 0; // array lengths omcgpu_update checks
  int n_collision_tallies = 0;       // tallies scored before the collision (collision estimator)
  int n_after_collision_tallies = 0; // ... after it, from the collision records (analog estimator)
  double* d_lost_rec = nullptr;      // state of the first lost particles of the generation (mark_lost)
  double* d_sec_site = nullptr;      // secondary-neutron pool (GenCtx::sec_*), allocated when the model can create them
  int32_t* d_sec_next = nullptr;
  int32_t* d_sec_head = nullptr;
  int32_t* d_sec_count = nullptr;
  int32_t sec_cap = 0;
  int64_t n_lost_last = 0;           // lost particles of the last generation
  double* d_an_d = nullptr;          // collision records, AN_ND / AN_NI columns of `cap` entries
  int32_t* d_an_i = nullptr;
  // several temperatures per nuclide: device materials are (model material, cell sqrt(kT)) pairs
  int n_base_materials = 0;
  std::vector<int32_t> mat_base;    // [device material] -> model material (empty: identity)
  std::vector<double> mat_sqrtkT;   // [device material] the cell temperature it was resolved for
  std::vector<double> mat_mult;     // [device material] the cell density multiplier folded into its atom densities
  bool temps_expanded = false;      // the device materials were resolved per cell temperature (temperature tables)
  const char* geo_blob = nullptr; // CSG tables in one block (device.h, gld)
  size_t geo_bytes = 0;
  unsigned geo_smem = 0;        // bytes of it the geometry kernels stage in shared memory
  bool geo_spilled = false;     // a geometry table did not fit the block (sizing bug)
  int tally_replicas = 1;       // copies of the VALUE arrays k_score spreads its REDs over
  int64_t tally_stride = 0;
  cudaStream_t stream = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evx0[RING] = {}, evx1[RING] = {}, evp0 = nullptr, evp1 = nullptr;
  // per-kernel device times of a sweep, each on the stream its kernel runs on: [0] crossing, [1] collision, [2] scoring
  cudaEvent_t evk0[4][RING] = {}, evk1[4][RING] = {}; // [3]: scoring of the collision list's flights
  bool fuse_xs = true;          // k_xs + k_advance as one kernel (ms_xs then covers both)
  bool sort_xs = false;         // sweep mode 1, unfused: k_xs runs over the list sorted by (material, energy)
  bool sort_coll = false;       // OMCGPU_SORT_COLL=1: k_collide in that order too (measured on C5: 211 vs 202 ms, off)
  uint32_t *d_keys = nullptr, *d_keys_out = nullptr;
  int32_t *d_vals = nullptr, *d_vals_out = nullptr;
  void* d_sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  bool ordered = true;          // sweep mode 2: lists rebuilt in particle-index order after every sweep
  bool dynamic_score = false;   // OMCGPU_DYNAMIC_SCORE=1: the same for the scoring kernel
  bool dynamic = false;         // OMCGPU_DYNAMIC=1: persistent crossing / collision kernels whose warps fetch work from a cursor
  bool keyed = false;           // ... and stable-sorted by event class (material / surface crossed): OMCGPU_KEYED=1
  uint8_t* d_key_out = nullptr;
  int32_t* d_iota = nullptr;
  void* d_ksort_tmp = nullptr;
  size_t ksort_tmp_bytes = 0;
  void* d_part_tmp = nullptr;   // cub::DevicePartition scratch
  size_t part_tmp_bytes = 0;
  int64_t low_count = 0;        // OMCGPU_LOW_COUNT: below this many alive histories sweeps and tail use the lockstep
                                // lookup for every material (measured: no gain, profiles/r19_matrix.log; default off)
  bool batch_xs = false;        // some material has >= OMC_XS_BATCH_MIN nuclides: lockstep lookup kernels
  int sweep_mode = 2;           // 1: xs+advance | surface | collide over a merged alive list; 2: paired lists
  int score_ctas = 0;           // >0: k_score grid capped at this many blocks per SM (persistent, co-resident)
  int n_sm = 148;
  int64_t sync_every = 8;       // sweeps the host lets run between two counter read-backs
  int64_t tail_threshold = 16384;
  double probe_kT = 0.0;
  int score_part = 16;          // measured: 4 / 8 / 16 -> 138.9 / 131.9 / 131.3 ms (profiles/r11_matrix.log)
  int64_t id_offset_last = 0;
  FILE* trace = nullptr;        // OMCGPU_TRACE=<file>: one line per sweep {generation, sweep, collisions, crossings}
  int64_t generation_no = 0;
};
static Engine G;
static unsigned geo_smem_bytes()
{
  return G.geo_smem;
}
static bool batch_xs_enabled()
{
  return G.batch_xs;
}
static bool dynamic_enabled()
{
  return G.dynamic;
}
static int dyn_grid_sms()
{
  return G.n_sm;
}
static int64_t low_count_threshold()
{
  return G.low_count;
}
static int tail_batch_min()
{
  return G.low_count > 0 ? 1 : OMC_XS_BATCH_MIN;
}

// The continuous-energy tables (energy grids, xs rows, hash tables) are carved from one
// allocation so that ONE access-policy window can ask L2 to keep them resident while
// particle state and flight records stream through (OMCGPU_L2_PERSIST=1).
static char* g_arena = nullptr;
static size_t g_arena_size = 0, g_arena_used = 0, g_arena_align = 256;
static bool g_arena_on = false, g_arena_spilled = false;

template<class T>
static int dev_alloc(T** p, size_t count)
{
  if (g_arena_on) {
    const size_t bytes = (std::max<size_t>(count, 1) * sizeof(T) + g_arena_align - 1) / g_arena_align * g_arena_align;
    if (g_arena_used + bytes <= g_arena_size) {
      *p = reinterpret_cast<T*>(g_arena + g_arena_used);
      cudaMemset(*p, 0, bytes);
      g_arena_used += bytes;
      return 0;
    }
    g_arena_spilled = true;
  }
  void* q = nullptr;
  cudaError_t e = cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T));
  if (e != cudaSuccess)
    return fail(OMCGPU_E_ALLOCATE, "cudaMalloc failed: %s", cudaGetErrorString(e));
  cudaMemset(q, 0, std::max<size_t>(count, 1) * sizeof(T));
  G.dev_allocs.push_back(q);
  *p = static_cast<T*>(q);
  return 0;
}

template<class T>
static int upload(const T** field, size_t count)
{
  const T* host = *field;
  T* d = nullptr;
  int rc = dev_alloc(&d, count);
  if (rc)
    return rc;
  if (host && count) {
    cudaError_t e = cudaMemcpy(d, host, count * sizeof(T), cudaMemcpyHostToDevice);
    if (e != cudaSuccess)
      return fail(OMCGPU_E_CUDA, "H2D copy failed: %s", cudaGetErrorString(e));
  }
  *field = d;
  return 0;
}

static int alloc_bank(Bank& b, int64_t cap)
{
  b.cap = cap;
  int rc = 0;
  rc |= dev_alloc(&b.x, cap);
  rc |= dev_alloc(&b.y, cap);
  rc |= dev_alloc(&b.z, cap);
  rc |= dev_alloc(&b.u, cap);
  rc |= dev_alloc(&b.v, cap);
  rc |= dev_alloc(&b.w, cap);
  rc |= dev_alloc(&b.E, cap);
  rc |= dev_alloc(&b.time, cap);
  rc |= dev_alloc(&b.wgt, cap);
  rc |= dev_alloc(&b.delayed_group, cap);
  rc |= dev_alloc(&b.parent_id, cap);
  rc |= dev_alloc(&b.progeny_id, cap);
  return rc;
}

// smallest double E whose HOST hash bin int(log(E/E_min)/spacing) is >= k
static double hash_edge_for(int k, double e_min, double spacing)
{
  auto bin = [&](double E) { return (int)(std::log(E / e_min) / spacing); };
  double E = e_min * std::exp(k * spacing);
  // walk to the exact transition in units of ulp
  while (bin(E) >= k)
    E = std::nextafter(E, 0.0);
  while (bin(E) < k)
    E = std::nextafter(E, INFINITY);
  return E;
}

// Several temperatures per nuclide (TemperatureMethod::NEAREST, src/nuclide.cpp:682-691): the table
// a lookup uses depends on the nuclide and on the cell's kT only, so it is resolved HERE, once per
// (material, cell temperature) pair: every such pair becomes a device material whose entries are
// table indices, and the cells point at those.  The kernels then run exactly the one-temperature
// code (no per-lookup temperature search, no kT in the collision registers).
struct TemperatureExpansion {
  std::vector<int32_t> cell_mat_ofs, cell_material, mat_id, mat_fissionable, mat_nuc_ofs, mat_nuclide, mat_base;
  std::vector<double> mat_atom_density, mat_sqrtkT, mat_mult;
};
// Cell density multipliers (Cell::density_mult, src/cell.cpp:104-111; Material::atom_density(i, rho),
// include/openmc/material.h:104) are resolved the same way: a (material, multiplier) pair is a device
// material whose atom densities are atom_density_(i) * multiplier -- the very product the reference
// forms at every lookup -- so the kernels carry neither the multiplier nor its test
// (src/particle.cpp:209: a change of multiplier is a change of device material).
static bool has_density_multipliers(const omcgpu_model* m)
{
  if (!m->cell_density_ofs || !m->cell_density_mult)
    return false;
  for (int k = 0; k < m->cell_density_ofs[m->n_cells]; ++k)
    if (m->cell_density_mult[k] != 1.0)
      return true;
  return false;
}
static int nearest_table(const omcgpu_model* m, int d, double kT)
{
  int best = m->nuct_ofs[d];
  double max_diff = INFINITY;
  for (int t = m->nuct_ofs[d]; t < m->nuct_ofs[d + 1]; ++t) {
    const double diff = std::fabs(m->nuc_kT[t] - kT);
    if (diff < max_diff) {
      best = t;
      max_diff = diff;
    }
  }
  return best;
}
static void expand_temperatures(const omcgpu_model* m, TemperatureExpansion& X)
{
  std::map<std::tuple<int32_t, double, double>, int32_t> derived;
  const bool temps = m->n_nuclide_temps > 0, dens = has_density_multipliers(m);
  X.cell_mat_ofs.push_back(0);
  for (int c = 0; c < m->n_cells; ++c) {
    const int m0 = m->cell_mat_ofs[c], nmat = m->cell_mat_ofs[c + 1] - m0;
    const int k0 = m->cell_sqrtkT_ofs[c], nkt = temps ? m->cell_sqrtkT_ofs[c + 1] - k0 : 0;
    const int d0 = dens ? m->cell_density_ofs[c] : 0, nd = dens ? m->cell_density_ofs[c + 1] - d0 : 0;
    const int n = nmat == 0 ? 0 : std::max(std::max(nmat, nd), std::max(nkt, 1));
    for (int k = 0; k < n; ++k) {
      const int32_t mat = m->cell_material[m0 + (nmat > 1 ? k : 0)];
      if (mat < 0) {
        X.cell_material.push_back(mat);
        continue;
      }
      const double sq = nkt > 0 ? m->cell_sqrtkT[k0 + (nkt > 1 ? k : 0)] : 0.0; // (the key of a model without temperature tables)
      const double mult = nd > 0 ? m->cell_density_mult[d0 + (nd > 1 ? k : 0)] : 1.0;
      auto it = derived.find({mat, sq, mult});
      if (it == derived.end()) {
        it = derived.emplace(std::make_tuple(mat, sq, mult), (int32_t)X.mat_base.size()).first;
        X.mat_base.push_back(mat);
        X.mat_sqrtkT.push_back(sq);
        X.mat_mult.push_back(mult);
      }
      X.cell_material.push_back(it->second);
    }
    X.cell_mat_ofs.push_back((int32_t)X.cell_material.size());
  }
  X.mat_nuc_ofs.push_back(0);
  for (size_t d = 0; d < X.mat_base.size(); ++d) {
    const int b = X.mat_base[d];
    const double kT = X.mat_sqrtkT[d] * X.mat_sqrtkT[d];
    X.mat_id.push_back(m->mat_id[b]);
    X.mat_fissionable.push_back(m->mat_fissionable[b]);
    for (int k = m->mat_nuc_ofs[b]; k < m->mat_nuc_ofs[b + 1]; ++k) {
      X.mat_nuclide.push_back(temps ? nearest_table(m, m->mat_nuclide[k], kT) : m->mat_nuclide[k]);
      X.mat_atom_density.push_back(m->mat_atom_density[k] * X.mat_mult[d]); // atom_density_(i) * rho_multiplier
    }
    X.mat_nuc_ofs.push_back((int32_t)X.mat_nuclide.size());
  }
}

static int upload_model(const omcgpu_model* m)
{
  G.hm = *m;
  omcgpu_model& h = G.hm;
  TemperatureExpansion X;
  G.n_base_materials = m->n_materials;
  G.mat_base.clear();
  G.mat_sqrtkT.clear();
  h.mat_base = nullptr;
  G.mat_mult.clear();
  G.temps_expanded = h.run_CE && m->n_nuclide_temps > 0;
  // continuous energy: host-side input only, folded into the device materials below; multigroup: the
  // lookup of the extended build reads them (uploaded after the geometry block)
  const bool mg_density = !h.run_CE && has_density_multipliers(m);
  h.cell_density_ofs = nullptr;
  h.cell_density_mult = nullptr;
  if (h.run_CE && (m->n_nuclide_temps > 0 || has_density_multipliers(m))) {
    expand_temperatures(m, X);
    h.n_materials = (int32_t)X.mat_base.size();
    h.cell_mat_ofs = X.cell_mat_ofs.data();
    h.cell_material = X.cell_material.data();
    h.mat_id = X.mat_id.data();
    h.mat_fissionable = X.mat_fissionable.data();
    h.mat_nuc_ofs = X.mat_nuc_ofs.data();
    h.mat_nuclide = X.mat_nuclide.data();
    h.mat_atom_density = X.mat_atom_density.data();
    h.mat_base = X.mat_base.data();
    h.n_nuclide_temps = 0; // the device sees resolved tables only
    h.nuct_ofs = nullptr;
    G.mat_base = X.mat_base;
    G.mat_sqrtkT = X.mat_sqrtkT;
    G.mat_mult = X.mat_mult;
  }
  const int ns = h.n_surfaces, nc = h.n_cells, nu = h.n_universes, nl = h.n_lattices;
  const int nm = h.n_materials;
  int rc = 0;
#define UP(field, count) rc = rc ? rc : upload(&h.field, (size_t)(count))
  // the CSG tables go into one block, hottest first: the kernels that walk the geometry stage its
  // first OMCGPU_GEO_SMEM bytes (default 24 KB, 0 = off) in shared memory (device.h, gld)
  {
    // upper bound of the geometry tables uploaded between here and lat_offsets (16-byte slots)
    const size_t n_tok_ = nc ? m->cell_region_ofs[nc] : 0, n_cm_ = nc ? h.cell_mat_ofs[nc] : 0,
                 n_kt_ = nc ? m->cell_sqrtkT_ofs[nc] : 0, n_co_ = nc ? m->cell_offset_ofs[nc] : 0,
                 n_uc_ = nu ? m->univ_cell_ofs[nu] : 0, n_lu_ = nl ? m->lat_univ_ofs[nl] : 0,
                 n_lo_ = nl ? m->lat_offset_ofs[nl] : 0;
    g_arena_size = 40 * 16 + (size_t)ns * (4 * 5 + 8 * 3 + 8 * OMCGPU_SURF_NCOEF + 2 * 4) + 80 * (size_t)h.n_quadrics +
                   (size_t)nc * (4 * 12 + 8 * 12) + 4 * (n_tok_ + n_cm_ + n_co_ + n_uc_ + n_lu_ + n_lo_) + 8 * n_kt_ +
                   4 * (2 * (size_t)ns + 1) + 8 * n_tok_ + /* half-space candidate lists: <= 2 entries per token */
                   (size_t)nu * 4 + (size_t)nl * (4 * 9 + 8 * 6) + 4096;
    g_arena_size = (g_arena_size + 255) / 256 * 256;
    g_arena_used = 0;
    g_arena_align = 16;
    void* q = nullptr;
    if (cudaMalloc(&q, g_arena_size) != cudaSuccess)
      return fail(OMCGPU_E_ALLOCATE, "cudaMalloc of the geometry block failed%s");
    G.dev_allocs.push_back(q);
    g_arena = static_cast<char*>(q);
    g_arena_on = true;
    g_arena_spilled = false;
    G.geo_blob = g_arena;
  }
  UP(surf_type, ns);
  UP(surf_bc, ns);
  UP(surf_periodic, ns);
  UP(surf_periodic_shift, (size_t)ns * 3);
  UP(surf_periodic_axis, ns);
  UP(surf_id, ns);
  UP(surf_coef, (size_t)ns * OMCGPU_SURF_NCOEF);
  UP(quad_coef, (size_t)h.n_quadrics * 10);
  const int n_tok = nc ? m->cell_region_ofs[nc] : 0;
  const int n_cmat = nc ? h.cell_mat_ofs[nc] : 0;
  const int n_ckt = nc ? m->cell_sqrtkT_ofs[nc] : 0;
  const int n_coff = nc ? m->cell_offset_ofs[nc] : 0;
  UP(cell_id, nc);
  UP(cell_type, nc);
  UP(cell_universe, nc);
  UP(cell_fill, nc);
  UP(cell_simple, nc);
  UP(cell_region_ofs, nc + 1);
  UP(region_tokens, n_tok);
  UP(cell_mat_ofs, nc + 1);
  UP(cell_material, n_cmat);
  UP(cell_sqrtkT_ofs, nc + 1);
  UP(cell_sqrtkT, n_ckt);
  G.n_cell_sqrtkT = n_ckt;
  UP(cell_translation, (size_t)nc * 3);
  UP(cell_rotated, nc);
  UP(cell_rotation, (size_t)nc * 9);
  UP(cell_distribcell_index, nc);
  UP(cell_offset_ofs, nc + 1);
  UP(cell_offset, n_coff);
  const int n_uc = nu ? m->univ_cell_ofs[nu] : 0;
  UP(univ_cell_ofs, nu + 1);
  UP(univ_cells, n_uc);
  // surface half-space -> cells that name it (a cell with unions / complements is
  // listed under both sides of every surface it names)
  std::vector<std::vector<int32_t>> by_tok(2 * (size_t)ns);
  for (int c = 0; c < nc; ++c)
    for (int k = m->cell_region_ofs[c]; k < m->cell_region_ofs[c + 1]; ++k) {
      const int32_t tok = m->region_tokens[k];
      if (tok >= OMCGPU_OP_UNION || tok == 0)
        continue;
      const int sfc = std::abs(tok) - 1;
      for (int side = 0; side < 2; ++side)
        if (!m->cell_simple[c] || side == (tok > 0 ? 1 : 0)) {
          auto& v = by_tok[2 * (size_t)sfc + side];
          if (v.empty() || v.back() != c)
            v.push_back(c);
        }
    }
  std::vector<int32_t> tok_ofs(2 * (size_t)ns + 1, 0), tok_cells;
  for (size_t t = 0; t < by_tok.size(); ++t) {
    tok_cells.insert(tok_cells.end(), by_tok[t].begin(), by_tok[t].end());
    tok_ofs[t + 1] = (int32_t)tok_cells.size();
  }
  h.tok_cell_ofs = tok_ofs.data();
  h.tok_cells = tok_cells.data();
  UP(tok_cell_ofs, tok_ofs.size());
  UP(tok_cells, tok_cells.size());
  const int n_lu = nl ? m->lat_univ_ofs[nl] : 0;
  const int n_lo = nl ? m->lat_offset_ofs[nl] : 0;
  UP(lat_shape, (size_t)nl * 3);
  UP(lat_is3d, nl);
  UP(lat_type, nl);
  h.n_hex_lattices = 0;
  for (int l = 0; l < nl; ++l)
    h.n_hex_lattices += m->lat_type[l] != OMCGPU_LATTICE_RECT ? 1 : 0;
  UP(lat_outer, nl);
  UP(lat_pitch, (size_t)nl * 3);
  UP(lat_lower_left, (size_t)nl * 3);
  UP(lat_univ_ofs, nl + 1);
  UP(lat_universes, n_lu);
  UP(lat_offset_ofs, nl + 1);
  UP(lat_offsets, n_lo);
  {
    g_arena_on = false;
    g_arena_align = 256;
    G.geo_bytes = g_arena_used;
    G.geo_spilled = g_arena_spilled;
    size_t budget = 24576;
    if (const char* t = std::getenv("OMCGPU_GEO_SMEM"))
      budget = (size_t)std::max(0, std::atoi(t));
    G.geo_smem = (unsigned)(std::min(G.geo_bytes, budget) / 16 * 16);
#if OMC_CUDA
    CUDA_OK(cudaMemcpyToSymbol(c_geo_blob, &G.geo_blob, sizeof(char*)));
#endif
  }
  if (mg_density) {
    h.cell_density_ofs = m->cell_density_ofs;
    h.cell_density_mult = m->cell_density_mult;
    UP(cell_density_ofs, nc + 1);
    UP(cell_density_mult, m->cell_density_ofs[nc]);
  }
  const int n_mn = nm ? h.mat_nuc_ofs[nm] : 0;
  if (h.mat_base)
    UP(mat_base, nm);
  UP(mat_id, nm);
  UP(mat_fissionable, nm);
  UP(mat_nuc_ofs, nm + 1);
  UP(mat_nuclide, n_mn);
  UP(mat_atom_density, n_mn);
  G.n_mat_nuclides = n_mn;
  if (h.run_CE) {
    const int nn = h.n_nuclides;
    const int64_t n_pts = m->nuc_grid_ofs[nn];
    // hash-edge table from the HOST libm (see ce_log_index)
    std::vector<double> edge(h.n_log_bins + 2);
    edge[0] = 0.0;
    for (int k = 1; k <= h.n_log_bins; ++k)
      edge[k] = hash_edge_for(k, h.energy_min, h.log_spacing);
    edge[h.n_log_bins + 1] = INFINITY;
    h.hash_edge = edge.data();
    UP(hash_edge, h.n_log_bins + 2);
    UP(nuc_awr, nn);
    UP(nuc_kT, nn);
    UP(nuc_fissionable, nn);
    UP(nuc_grid_ofs, nn + 1);
    {
      int per = OMCGPU_FINE_PER_LOG_BIN;
      if (const char* t = std::getenv("OMCGPU_FINE_PER_LOG_BIN"))
        per = std::max(1, std::atoi(t));
      g_arena_size = (size_t)n_pts * 8 + (size_t)n_pts * 32 + (size_t)nn * (h.n_log_bins + 1) * 4 +
                     (size_t)nn * h.n_log_bins * per * 8 + 4096;
      g_arena_used = 0;
      void* q = nullptr;
      if (cudaMalloc(&q, g_arena_size) != cudaSuccess)
        return fail(OMCGPU_E_ALLOCATE, "cudaMalloc of the cross-section arena failed%s");
      G.dev_allocs.push_back(q);
      g_arena = static_cast<char*>(q);
      g_arena_on = true;
    }
    UP(nuc_energy, n_pts);
    UP(nuc_xs, (size_t)n_pts * 4);
    UP(nuc_grid_index, (size_t)nn * (h.n_log_bins + 1));
    // fine hash (see ce_fine_index): per nuclide one merge of the sorted grid with
    // the sorted, widened bin edges
    std::vector<int32_t> fine;
    {
      int per = OMCGPU_FINE_PER_LOG_BIN;
      if (const char* t = std::getenv("OMCGPU_FINE_PER_LOG_BIN"))
        per = std::max(1, std::atoi(t));
      const int nf = h.n_log_bins * per;
      h.n_fine_bins = nf;
      const double total = h.log_spacing * h.n_log_bins;
      h.fine_inv_spacing = nf / total;
      const double sp = total / nf, widen = 1e-6;
      std::vector<double> e_lo(nf), e_hi(nf);
      for (int j = 0; j < nf; ++j) {
        e_lo[j] = j == 0 ? 0.0 : h.energy_min * std::exp((j - widen) * sp) * (1.0 - 1e-12);
        e_hi[j] = j == nf - 1 ? INFINITY : h.energy_min * std::exp((j + 1 + widen) * sp) * (1.0 + 1e-12);
      }
      fine.resize((size_t)nn * nf * 2);
      for (int k = 0; k < nn; ++k) {
        const double* e = m->nuc_energy + m->nuc_grid_ofs[k];
        const int64_t ng = m->nuc_grid_ofs[k + 1] - m->nuc_grid_ofs[k];
        int64_t c_lo = 0, c_hi = 0; // number of grid points below the edge (edges ascend)
        for (int j = 0; j < nf; ++j) {
          while (c_lo < ng && e[c_lo] < e_lo[j])
            ++c_lo;
          while (c_hi < ng && e[c_hi] < e_hi[j])
            ++c_hi;
          fine[((size_t)j * nn + k) * 2 + 0] = (int32_t)c_lo; // bin-major: the nuclides of a material
          fine[((size_t)j * nn + k) * 2 + 1] = (int32_t)c_hi; // share the sectors of one bin
        }
      }
      h.fine_window = fine.data();
    }
    UP(fine_window, fine.size());
    g_arena_on = false;
    UP(nuc_rx_ofs, nn + 1);
    UP(nuc_fission_rx, nn);
    UP(nuc_n_precursor, nn);
    UP(nuc_total_nu_fn, nn);
    const int nr = h.n_reactions, np = h.n_products, nf = h.n_functions;
    const int64_t n_rxs = nr ? m->rx_xs_ofs[nr] : 0;
    UP(rx_mt, nr);
    UP(rx_scatter_in_cm, nr);
    UP(rx_inelastic, nr);
    UP(rx_threshold, nr);
    UP(rx_xs_ofs, nr + 1);
    UP(rx_xs, n_rxs);
    UP(rx_q_value, nr);
    UP(rx_prod_ofs, nr + 1);
    UP(prod_emission, np);
    UP(prod_decay_rate, np);
    UP(prod_yield_fn, np);
    UP(prod_elaw, np);
    UP(prod_par, (size_t)np * 2);
    UP(prod_fn, (size_t)np * 2);
    const int n_fd = nf ? m->fn_ofs[nf] : 0;
    UP(fn_type, nf);
    UP(fn_ofs, nf + 1);
    UP(fn_data, n_fd);
  } else {
    const int Gn = h.mg_n_groups, DG = h.mg_n_delayed;
    const int n_blk = m->mgm_blk_ofs[nm];
    const int n_ba = m->blk_ang_ofs[n_blk];
    UP(mg_energy_bin_avg, Gn);
    UP(mg_energy_bins, Gn + 1);
    UP(mgm_n_temps, nm);
    UP(mgm_n_pol, nm);
    UP(mgm_n_azi, nm);
    UP(mgm_isotropic, nm);
    UP(mgm_fissionable, nm);
    UP(mgm_blk_ofs, nm + 1);
    UP(mgm_kTs, n_blk);
    UP(blk_ang_ofs, n_blk + 1);
    UP(mg_total, (size_t)n_ba * Gn);
    UP(mg_absorption, (size_t)n_ba * Gn);
    UP(mg_nu_fission, (size_t)n_ba * Gn);
    UP(mg_prompt_nu_fission, (size_t)n_ba * Gn);
    UP(mg_fission, (size_t)n_ba * Gn);
    UP(mg_inverse_velocity, (size_t)n_ba * Gn);
    UP(mg_delayed_nu_fission, (size_t)n_ba * DG * Gn);
    UP(mg_chi_prompt, (size_t)n_ba * Gn * Gn);
    UP(mg_chi_delayed, (size_t)n_ba * DG * Gn * Gn);
    UP(mg_decay_rate, (size_t)n_ba * DG);
    // ragged scatter sizes
    size_t n_e = 0, n_d = 0, n_mu = 0;
    for (int ba = 0; ba < n_ba; ++ba) {
      const int order = m->sc_order[ba];
      for (int g = 0; g < Gn; ++g) {
        const int ng = m->sc_gmax[ba * Gn + g] - m->sc_gmin[ba * Gn + g] + 1;
        n_e = std::max<size_t>(n_e, (size_t)m->sc_gofs[ba * Gn + g] + ng);
        n_d = std::max<size_t>(n_d, (size_t)m->sc_dofs[ba * Gn + g] + (size_t)ng * order);
      }
      const int fmt = m->sc_format[ba];
      const int nmu = fmt == OMCGPU_SCATT_TABULAR ? order : (fmt == OMCGPU_SCATT_HISTOGRAM ? order + 1 : 1);
      n_mu = std::max<size_t>(n_mu, (size_t)m->sc_mu_ofs[ba] + nmu);
    }
    UP(sc_format, n_ba);
    UP(sc_order, n_ba);
    UP(sc_gmin, (size_t)n_ba * Gn);
    UP(sc_gmax, (size_t)n_ba * Gn);
    UP(sc_gofs, (size_t)n_ba * Gn);
    UP(sc_dofs, (size_t)n_ba * Gn);
    UP(sc_scattxs, (size_t)n_ba * Gn);
    UP(sc_energy, n_e);
    UP(sc_mult, n_e);
    UP(sc_maxval, n_e);
    UP(sc_dist, n_d);
    UP(sc_fmu, n_d);
    UP(sc_mu_ofs, n_ba);
    UP(sc_mu, n_mu);
    UP(sc_dmu, n_ba);
  }
  const int nme = h.n_meshes, nfl = h.n_filters, nt = h.n_tallies;
  UP(mesh_ndim, nme);
  UP(mesh_shape, (size_t)nme * 3);
  UP(mesh_lower_left, (size_t)nme * 3);
  UP(mesh_width, (size_t)nme * 3);
  const int n_fdat = nfl ? m->filter_data_ofs[nfl] : 0;
  UP(filter_type, nfl);
  UP(filter_n_bins, nfl);
  UP(filter_mesh, nfl);
  UP(filter_matches_groups, nfl);
  UP(filter_data_ofs, nfl + 1);
  UP(filter_idata, n_fdat);
  UP(filter_ddata, n_fdat);
  h.n_collision_tallies = 0;
  G.n_after_collision_tallies = 0;
  for (int t = 0; t < nt; ++t) { // (the same split as device.h tally_scored_after_collision)
    const bool after = m->tally_estimator[t] == OMCGPU_ESTIMATOR_ANALOG ||
                       (m->tally_estimator[t] == OMCGPU_ESTIMATOR_COLLISION && !m->run_CE && m->survival_biasing);
    G.n_after_collision_tallies += after ? 1 : 0;
    h.n_collision_tallies += m->tally_estimator[t] == OMCGPU_ESTIMATOR_COLLISION && !after ? 1 : 0;
  }
  G.n_collision_tallies = h.n_collision_tallies;
  UP(tally_estimator, nt);
  G.n_tally = nt ? m->tally_result_ofs[nt] : 0;
  G.tally_ofs.assign(m->tally_result_ofs, m->tally_result_ofs + (nt ? nt + 1 : 0));
  UP(tally_n_filters, nt);
  UP(tally_filters, (size_t)nt * OMCGPU_MAX_FILTERS);
  UP(tally_strides, (size_t)nt * OMCGPU_MAX_FILTERS);
  UP(tally_n_scores, nt);
  UP(tally_scores, (size_t)nt * OMCGPU_MAX_SCORES);
  UP(tally_result_ofs, nt + 1);
#undef UP
  if (rc)
    return rc;
#if OMC_CUDA
#if !defined(OMC_GEO_NEVER)
  { // (device.h, gld): geometry pointers become offsets into the geometry block
#if defined(OMC_GEO_ALWAYS)
    if (G.geo_smem < G.geo_bytes)
      return fail(OMCGPU_E_UNSUPPORTED, "OMC_GEO_ALWAYS build: the geometry block does not fit the shared-memory budget%s");
#else
    if (G.geo_smem < G.geo_bytes)
      G.geo_smem = 0; // all or nothing: a block larger than the budget is read from global memory
    if (G.geo_spilled)
      return fail(OMCGPU_E_ALLOCATE, "internal: the geometry block was sized too small%s");
#endif
    omcgpu_model hd = h;
#define REBASE(field) hd.field = reinterpret_cast<decltype(hd.field)>(reinterpret_cast<const char*>(h.field) - G.geo_blob)
    REBASE(surf_type); REBASE(surf_bc); REBASE(surf_periodic); REBASE(surf_periodic_shift); REBASE(surf_periodic_axis); REBASE(surf_id);
    REBASE(surf_coef); REBASE(quad_coef); REBASE(cell_id); REBASE(cell_type); REBASE(cell_universe); REBASE(cell_fill);
    REBASE(cell_simple); REBASE(cell_region_ofs); REBASE(region_tokens); REBASE(cell_mat_ofs); REBASE(cell_material);
    REBASE(cell_sqrtkT_ofs); REBASE(cell_sqrtkT); REBASE(cell_translation); REBASE(cell_rotated); REBASE(cell_rotation);
    REBASE(cell_distribcell_index); REBASE(cell_offset_ofs); REBASE(cell_offset); REBASE(univ_cell_ofs);
    REBASE(univ_cells); REBASE(tok_cell_ofs); REBASE(tok_cells); REBASE(lat_shape); REBASE(lat_is3d); REBASE(lat_type); REBASE(lat_outer);
    REBASE(lat_pitch); REBASE(lat_lower_left); REBASE(lat_univ_ofs); REBASE(lat_universes); REBASE(lat_offset_ofs);
    REBASE(lat_offsets);
#undef REBASE
    CUDA_OK(cudaMemcpyToSymbol(c_dm, &hd, sizeof(omcgpu_model)));
  }
#else
  CUDA_OK(cudaMemcpyToSymbol(c_dm, &h, sizeof(omcgpu_model)));
#endif
#else
  c_dm = h;
#endif
  return 0;
}

// inelastic_scatter creates yield - 1 secondary neutrons when a product's yield is an integer >= 2,
// e.g. (n,2n) (src/physics.cpp:1158-1162): can a yield function of the model take such a value on
// an interval?  (a constant, or a flat piece / flat extrapolation of a tabulated function)
static bool model_creates_secondaries(const omcgpu_model* m)
{
  if (!m->run_CE)
    return false;
  for (int r = 0; r < m->n_reactions; ++r) {
    if (!m->rx_inelastic[r] || m->rx_prod_ofs[r + 1] <= m->rx_prod_ofs[r])
      continue;
    const int fn = m->prod_yield_fn[m->rx_prod_ofs[r]];
    if (fn < 0 || fn >= m->n_functions)
      continue;
    const double* d = m->fn_data + m->fn_ofs[fn];
    const int n = m->fn_ofs[fn + 1] - m->fn_ofs[fn];
    auto integral = [](double y) { return y >= 2.0 && std::floor(y) == y; };
    if (m->fn_type[fn] == OMCGPU_FN_POLYNOMIAL) {
      bool constant = true;
      for (int k = 1; k < n; ++k)
        constant = constant && d[k] == 0.0;
      if (n >= 1 && (!constant || integral(d[0])))
        return true; // (a non-constant polynomial may cross an integer: keep the pool)
    } else {
      const int np = n / 2;
      const double* y = d + np;
      for (int k = 0; k < np; ++k)
        if (integral(y[k]))
          return true;
    }
  }
  return false;
}

} // namespace omc (the C entry point below is used by check_supported)
extern "C" int omcgpu_needs_extended_build(const omcgpu_model* m)
{
  if (!m)
    return 0;
  for (int l = 0; l < m->n_lattices; ++l)
    if (m->lat_type && m->lat_type[l] != OMCGPU_LATTICE_RECT)
      return 1; // hexagonal lattices
  for (int s = 0; s < m->n_surfaces; ++s) {
    if (m->surf_type[s] >= OMCGPU_SURF_XCONE)
      return 1; // axis cones, general quadrics
    if (m->surf_bc[s] == OMCGPU_BC_WHITE || (m->surf_bc[s] == OMCGPU_BC_PERIODIC && m->surf_periodic_axis[s] >= 0))
      return 1; // white and rotational periodic boundaries
  }
  for (int t = 0; t < m->n_tallies; ++t)
    if (m->tally_estimator[t] == OMCGPU_ESTIMATOR_ANALOG ||
        (m->tally_estimator[t] == OMCGPU_ESTIMATOR_COLLISION && !m->run_CE && m->survival_biasing))
      return 1; // tallies scored after the collision from the collision records
  if (!m->run_CE && omc::has_density_multipliers(m))
    return 1; // cell density multipliers of a multigroup model (continuous energy: resolved at omcgpu_init)
  return omc::model_creates_secondaries(m) ? 1 : 0; // secondary neutrons of integer yields
}
namespace omc {

static int check_supported(const omcgpu_model* m)
{
  if (m->n_coord_levels > OMCGPU_MAX_COORD)
    return fail(OMCGPU_E_UNSUPPORTED, "geometry nests deeper than OMCGPU_MAX_COORD%s");
  for (int s = 0; s < m->n_surfaces; ++s) {
    if (m->surf_type[s] < 0 || m->surf_type[s] > OMCGPU_SURF_QUADRIC)
      return fail(OMCGPU_E_UNSUPPORTED, "surface type not supported on the device%s");
    if (m->surf_bc[s] < 0 || m->surf_bc[s] > OMCGPU_BC_WHITE)
      return fail(OMCGPU_E_UNSUPPORTED, "boundary condition not supported on the device%s");
    if (m->surf_bc[s] == OMCGPU_BC_PERIODIC &&
        (m->surf_type[s] > OMCGPU_SURF_PLANE || m->surf_periodic[s] < 0 || m->surf_periodic[s] >= m->n_surfaces ||
          m->surf_periodic_axis[s] < -1 || m->surf_periodic_axis[s] > 2))
      return fail(OMCGPU_E_UNSUPPORTED, "periodic boundary: only between two planes%s");
  }
  if (omcgpu_needs_extended_build(m) && !OMC_EXTENDED)
    return fail(OMCGPU_E_UNSUPPORTED, "hexagonal lattices, axis cones, general quadrics, white / rotational periodic "
                                      "boundaries, analog-estimator tallies, multigroup density multipliers and secondary neutrons are "
                                      "compiled into libomcgpu_ext.so (-DOMC_EXTENDED=1), not into this build: load that "
                                      "library (omcgpu_needs_extended_build)%s");
  for (int t = 0; t < m->n_tallies; ++t) {
    if (m->tally_n_filters[t] > OMCGPU_MAX_FILTERS || m->tally_n_scores[t] > OMCGPU_MAX_SCORES)
      return fail(OMCGPU_E_UNSUPPORTED, "tally has too many filters or scores%s");
    if (m->tally_estimator[t] != OMCGPU_ESTIMATOR_TRACKLENGTH && m->tally_estimator[t] != OMCGPU_ESTIMATOR_COLLISION &&
        m->tally_estimator[t] != OMCGPU_ESTIMATOR_ANALOG)
      return fail(OMCGPU_E_UNSUPPORTED, "unknown tally estimator%s");
    int n_mesh = 0;
    for (int k = 0; k < m->tally_n_filters[t]; ++k) {
      const int ft = m->filter_type[m->tally_filters[t * OMCGPU_MAX_FILTERS + k]];
      if (ft == OMCGPU_FILTER_MESH)
        ++n_mesh;
      if (ft == OMCGPU_FILTER_ENERGYOUT) {
        if (m->tally_estimator[t] != OMCGPU_ESTIMATOR_ANALOG)
          return fail(OMCGPU_E_UNSUPPORTED, "an energyout filter needs the analog estimator%s");
      }
    }
    if (n_mesh > 1)
      return fail(OMCGPU_E_UNSUPPORTED, "more than one mesh filter per tally%s");
  }
  if (m->run_CE) {
    for (int p = 0; p < m->n_products; ++p)
      if (m->prod_elaw[p] < 0 || m->prod_elaw[p] > OMCGPU_ELAW_WATT)
        return fail(OMCGPU_E_UNSUPPORTED, "secondary energy law not supported on the device%s");
  }
  return 0;
}

static EvArgs make_args(const omcgpu_gen_in* in)
{
  EvArgs a;
  std::memset(&a, 0, sizeof(a));
  a.ps = G.ps;
  a.src = G.src;
  a.fb = G.fis;
  if (in) {
    a.gc.n = in->n_particles;
    a.gc.id_offset = in->id_offset;
    a.gc.seed_base = in->particle_seed_base;
    a.gc.master_seed = in->master_seed;
    a.gc.stride = in->stride;
    a.gc.keff = in->keff;
    a.gc.tallies_active = in->tallies_active;
  }
  a.gc.n_fission = G.d_counts + 4;
  a.gc.n_lost = G.d_counts + 5;
  a.gc.n_bank_full = G.d_counts + 7;
  a.gc.counters = G.d_counters;
  a.gc.count_lookups = G.count_lookups ? 1 : 0;
  a.gc.tally_values = G.d_tally;
  a.gc.tally_replicas = G.tally_replicas;
  a.gc.tally_replica_stride = G.tally_stride;
  a.gc.fission_cap = G.fis_cap;
  a.gc.score_part = G.score_part;
  a.gc.tally_needs_cells = G.tally_needs_cells ? 1 : 0;
  a.gc.an_d = G.d_an_d;
  a.gc.an_i = G.d_an_i;
  a.gc.an_cap = G.cap;
  a.gc.lost_rec = G.d_lost_rec;
  a.gc.lost_cap = OMCGPU_MAX_LOST_RECORDS;
  a.gc.sec_site = G.d_sec_site;
  a.gc.sec_next = G.d_sec_next;
  a.gc.sec_head = G.d_sec_head;
  a.gc.sec_count = G.d_sec_count;
  a.gc.sec_cap = G.sec_cap;
  a.fl = G.fl[0];
  a.fl_buf = 0;
  a.first_event = NEXT_NONE;
  return a;
}

static int read_counts()
{
  CUDA_OK(cudaMemcpyAsync(G.h_counts, G.d_counts, N_COUNTS * sizeof(int32_t), cudaMemcpyDeviceToHost, G.stream));
  CUDA_OK(cudaStreamSynchronize(G.stream));
  return 0;
}

static int sort_fission_bank(int64_t n, int64_t id_offset)
{
  // exclusive scan of progeny_per_particle, then scatter (src/bank.cpp:76-134)
#if OMC_CUDA
  CUDA_OK(cub::DeviceScan::ExclusiveSum(G.d_scan_tmp, G.scan_tmp_bytes, G.ps.n_progeny, G.d_offs,
    (int)n, G.stream));
  ++g_launches;
#else
  std::exclusive_scan(G.ps.n_progeny, G.ps.n_progeny + n, G.d_offs, 0);
#endif
  BankArgs b;
  std::memset(&b, 0, sizeof(b));
  b.a = G.fis;
  b.b = G.fis_sorted;
  b.offs = G.d_offs;
  b.n = G.n_fis;
  b.id_offset = id_offset;
  OMC_LAUNCH(k_scatter_bank, G.n_fis, G.stream, b);
  return 0;
}

struct SweepStats {
  int64_t n_xs = 0, n_coll = 0, n_cross = 0;
  double ms_xs = 0.0;
  double ms_kernel[3] = {0.0, 0.0, 0.0}; // crossing, collision, scoring kernels (sum of launch durations)
  int64_t n_launch[3] = {0, 0, 0};
  // scoring runs two sweeps behind the event kernels on its own stream: its event pairs (ring
  // slot = launch number % RING) are read only once the main stream is known to have waited for them
  int64_t score_issued = 0, score_read = 0;
  int score_pairs = 2; // event pairs per scoring launch number (paired mode: one per list)
};
// adds the elapsed times of ring slot `ev` (kernels that did not launch recorded no pair)
static void drain_score_times(SweepStats& st, int64_t upto)
{
  for (; st.score_read < upto && st.score_read < st.score_issued; ++st.score_read) {
    const int ev = (int)(st.score_read % RING);
    for (int k = 2; k < 4; ++k) {
      float ms = 0.f;
      if (k == 3 && (st.score_read == 0 || st.score_pairs < 2))
        continue; // the first flight (and the merged sweep) has one scoring launch
      if (cudaEventElapsedTime(&ms, G.evk0[k][ev], G.evk1[k][ev]) == cudaSuccess) {
        st.ms_kernel[2] += ms;
        ++st.n_launch[2];
      } else {
        cudaGetLastError();
      }
    }
  }
}
static void add_kernel_times(SweepStats& st, int ev, const bool* used)
{
  for (int k = 0; k < 3; ++k)
    if (used[k]) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, G.evk0[k][ev], G.evk1[k][ev]) == cudaSuccess) {
        st.ms_kernel[k] += ms;
        ++st.n_launch[k];
      }
    }
}

// Order-preserving list build: the histories flagged 1 in ps.next_ev go to `coll`,
// those flagged 2 to `surf`, both in ascending particle index, and the two counts
// to counts[0..1].  Appending with atomics instead leaves every list in the order
// the warps happened to finish: after a few sweeps a warp's 32 particles are 32
// unrelated sectors of every column.  Index order keeps a warp's gathers within a
// few neighbouring lines and DRAM pages.
#if OMC_CUDA
struct FlagIs {
  const uint8_t* f;
  uint8_t v;
  __device__ __forceinline__ bool operator()(int i) const { return f[i] == v; }
};
#endif
static int build_lists(int64_t n, int32_t* coll, int32_t* surf, int32_t* counts, cudaStream_t stream,
  uint8_t v1 = 1, uint8_t v2 = 2)
{
#if OMC_CUDA
  FlagIs is_coll{G.ps.next_ev, v1}, is_surf{G.ps.next_ev, v2};
  thrust::counting_iterator<int32_t> ids(0);
  thrust::discard_iterator<> none;
  size_t bytes = G.part_tmp_bytes;
  CUDA_OK(cub::DevicePartition::If(G.d_part_tmp, bytes, ids, coll, surf, none, counts, (int)n, is_coll,
    is_surf, stream));
  g_launches += 2;
#else
  (void)stream;
  counts[0] = counts[1] = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (G.ps.next_ev[i] == v1)
      coll[counts[0]++] = (int32_t)i;
    else if (G.ps.next_ev[i] == v2)
      surf[counts[1]++] = (int32_t)i;
  }
#endif
  return 0;
}

// Keyed list build: one stable radix pass over the class bytes of all n histories (see
// event_flag) leaves the collision list in list[0, counts[0]) and the crossing list behind it,
// each sorted by class with ascending particle index inside a class; the dead sort to the end.
static int build_lists_keyed(int64_t n, int32_t* list, int32_t* counts, cudaStream_t stream)
{
#if OMC_CUDA
  size_t bytes = G.ksort_tmp_bytes;
  CUDA_OK(cub::DeviceRadixSort::SortPairs(G.d_ksort_tmp, bytes, G.ps.next_ev, G.d_key_out, G.d_iota, list, (int)n,
    0, 8, stream));
  g_launches += 3;
#else
  (void)stream;
  std::vector<int32_t> order((size_t)n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [](int32_t x, int32_t y) { return G.ps.next_ev[x] < G.ps.next_ev[y]; });
  for (int64_t k = 0; k < n; ++k) {
    list[k] = order[k];
    G.d_key_out[k] = G.ps.next_ev[order[k]];
  }
#endif
  BoundsArgs b;
  b.keys = G.d_key_out;
  b.n = n;
  b.counts = counts;
  OMC_LAUNCH(k_bounds, 2, stream, b);
  return 0;
}

// work list `q` re-ordered by (material, energy) into G.d_vals_out (see body_xs_keys)
static int sort_by_energy(const Queue& q, int64_t n_ub, cudaStream_t stream)
{
  KeyArgs ka;
  ka.ps = G.ps;
  ka.q_in = q;
  ka.keys = G.d_keys;
  ka.vals = G.d_vals;
  OMC_LAUNCH(k_xs_keys, n_ub, stream, ka);
#if OMC_CUDA
  size_t bytes = G.sort_tmp_bytes;
  CUDA_OK(cub::DeviceRadixSort::SortPairs(G.d_sort_tmp, bytes, G.d_keys, G.d_keys_out, G.d_vals, G.d_vals_out,
    (int)n_ub, 0, 32, stream));
  g_launches += 4;
#else
  std::vector<std::pair<uint32_t, int32_t>> kv((size_t)n_ub);
  for (int64_t k = 0; k < n_ub; ++k)
    kv[k] = {G.d_keys[k], G.d_vals[k]};
  std::stable_sort(kv.begin(), kv.end(), [](auto& x, auto& y) { return x.first < y.first; });
  for (int64_t k = 0; k < n_ub; ++k)
    G.d_vals_out[k] = kv[k].second;
#endif
  return 0;
}

// ---- sweep mode 1: one merged alive list; xs+advance, then crossing / collision --
static int run_sweeps_merged(EvArgs& a, const omcgpu_gen_in* in, int64_t n, SweepStats& st)
{
  int rc = 0;
  // Work lists: q[0] / q[1] alternate as "alive" list (input of k_xs / k_advance)
  // and "next" list; q[2] collisions, q[3] crossings.  Their COUNTERS come from a
  // ring with one slot {collision, crossing, next} per sweep, so that the host
  // can let several sweeps run ahead without reading anything back: a kernel
  // always finds its input count in device memory, launched over the last count
  // the host knows (counts never grow within a generation).
  int cur = 0;
  st.score_pairs = 1;
  a.q_a = G.q[cur];
  a.q_a.count = G.d_counts + 0;
  OMC_LAUNCH(k_init, n, G.stream, a);
  rc = read_counts();
  if (rc)
    return rc;
  int64_t n_alive = G.h_counts[0]; // last count read back: upper bound of the lists
  int32_t* alive_count = G.d_counts + 0;
  a.ordered = G.ordered ? 1 : 0;
  if (G.ordered)
    CUDA_OK(cudaMemsetAsync(G.ps.next_ev, 0, (size_t)n, G.stream));
  int64_t n_xs = 0, n_coll = 0, n_cross = 0;
  double ms_xs = 0.0;
  const bool score = in->tallies_active && G.hm.n_tallies > 0;
  // Streams: the three consumers of a flight (tally scoring, surface crossing,
  // collision) touch disjoint data and run side by side; k_score reads only the
  // flight records, which are double-buffered, so it also overlaps the next
  // sweep's k_xs / k_advance.
  cudaStream_t st_score = G.overlap ? G.s_score : G.stream;
  cudaStream_t st_surf = G.overlap ? G.s_surf : G.stream;
  cudaStream_t st_coll = G.overlap ? G.s_coll : G.stream;
  bool score_pending[2] = {false, false};
  int64_t sweep = 0, synced = 0; // sweeps issued / sweeps whose counters the host has read
  int64_t n_in_first = n_alive;  // input count of sweep `synced`
  while (n_alive > 0 && n_alive > G.tail_threshold) {
    const int nxt = 1 - cur;
    const int buf = (int)(sweep & 1);
    int32_t* slot = G.d_counts + RING_OFS + SLOT * (sweep % RING);
    CUDA_OK(cudaMemsetAsync(slot, 0, SLOT * sizeof(int32_t), G.stream));
    a.q_in = G.q[cur];
    a.q_in.count = alive_count;
    const int ev = (int)(sweep % RING);
    a.q_a = G.q[2];
    a.q_a.count = slot + 0;
    a.q_b = G.q[3];
    a.q_b.count = slot + 1;
    a.fl = G.fl[buf];
    a.fl_buf = buf;
    a.n_flights = alive_count;
    if (score && score_pending[buf] && G.overlap) // records of sweep-2 must have been consumed
      CUDA_OK(cudaStreamWaitEvent(G.stream, G.e_score[buf], 0));
    if (G.fuse_xs) {
      CUDA_OK(cudaEventRecord(G.evx0[ev], G.stream));
      LAUNCH_XS(k_xs_advance, n_alive, G.stream, a);
      CUDA_OK(cudaEventRecord(G.evx1[ev], G.stream));
    } else {
      if (G.sort_xs) {
        rc = sort_by_energy(a.q_in, n_alive, G.stream);
        if (rc)
          return rc;
        EvArgs ax = a; // the same histories, in (material, energy) order
        ax.q_in.idx = G.d_vals_out;
        CUDA_OK(cudaEventRecord(G.evx0[ev], G.stream)); // ms_xs: the lookup kernel alone
        LAUNCH_XS(k_xs, n_alive, G.stream, ax);
      } else {
        CUDA_OK(cudaEventRecord(G.evx0[ev], G.stream));
        LAUNCH_XS(k_xs, n_alive, G.stream, a);
      }
      CUDA_OK(cudaEventRecord(G.evx1[ev], G.stream));
      OMC_LAUNCH(k_advance, n_alive, G.stream, a);
    }
    if (G.ordered) { // collision / crossing lists in particle-index order (see build_lists)
      rc = build_lists(n, G.q[2].idx, G.q[3].idx, slot, G.stream);
      if (rc)
        return rc;
    }
    // cell / universe filters read the live coordinate levels: score before the
    // crossing and collision kernels may change them
    const bool score_in_order = score && G.tally_needs_cells;
    a.cursor = nullptr;
    if (score_in_order)
      LAUNCH_SCORE(n_alive, G.stream, a);
    if (G.overlap)
      CUDA_OK(cudaEventRecord(G.e_adv, G.stream));
    if (score && !score_in_order) {
      if (G.overlap)
        CUDA_OK(cudaStreamWaitEvent(st_score, G.e_adv, 0));
      CUDA_OK(cudaEventRecord(G.evk0[2][ev], st_score));
      a.cursor = nullptr;
      LAUNCH_SCORE(n_alive, st_score, a);
      CUDA_OK(cudaEventRecord(G.evk1[2][ev], st_score));
      st.score_issued = sweep + 1;
      if (G.overlap) {
        CUDA_OK(cudaEventRecord(G.e_score[buf], st_score));
        score_pending[buf] = true;
      }
    }
    a.q_in = G.q[3];
    a.q_in.count = slot + 1;
    a.q_a = G.q[nxt];
    a.q_a.count = slot + 2;
    if (G.overlap)
      CUDA_OK(cudaStreamWaitEvent(st_surf, G.e_adv, 0));
    CUDA_OK(cudaEventRecord(G.evk0[0][ev], st_surf));
    OMC_LAUNCH(k_surface, n_alive, st_surf, a);
    CUDA_OK(cudaEventRecord(G.evk1[0][ev], st_surf));
    a.q_in = G.q[2];
    a.q_in.count = slot + 0;
    if (G.overlap) {
      CUDA_OK(cudaEventRecord(G.e_surf, st_surf));
      CUDA_OK(cudaStreamWaitEvent(st_coll, G.e_adv, 0));
    }
    if (G.sort_xs && G.sort_coll) {
      // the nuclide-sampling walk of a many-nuclide material is a run of lookups: same
      // ordering, same coalescing (the sort buffers are free again: the lookup is done)
      rc = sort_by_energy(a.q_in, n_alive, st_coll);
      if (rc)
        return rc;
      EvArgs ac = a;
      ac.q_in.idx = G.d_vals_out;
      CUDA_OK(cudaEventRecord(G.evk0[1][ev], st_coll));
      LAUNCH_COLLISION_SCORE(n_alive, st_coll, ac);
      LAUNCH_XS(k_collide, n_alive, st_coll, ac);
      LAUNCH_AFTER_COLLISION_SCORE(n_alive, st_coll, ac);
    } else {
      CUDA_OK(cudaEventRecord(G.evk0[1][ev], st_coll));
      LAUNCH_COLLISION_SCORE(n_alive, st_coll, a);
      LAUNCH_XS(k_collide, n_alive, st_coll, a);
      LAUNCH_AFTER_COLLISION_SCORE(n_alive, st_coll, a);
    }
    CUDA_OK(cudaEventRecord(G.evk1[1][ev], st_coll));
    if (G.overlap) {
      CUDA_OK(cudaEventRecord(G.e_coll, st_coll));
      CUDA_OK(cudaStreamWaitEvent(G.stream, G.e_surf, 0));
      CUDA_OK(cudaStreamWaitEvent(G.stream, G.e_coll, 0));
    }
    if (G.ordered) { // the next alive list: survivors of both kernels, in index order (counts: slot[2], slot[3] = 0)
      rc = build_lists(n, G.q[nxt].idx, G.d_gather, slot + 2, G.stream, 3, 255);
      if (rc)
        return rc;
    }
    alive_count = slot + 2;
    cur = nxt;
    ++sweep;
    // read the counters back every SYNC_EVERY sweeps (and account for the sweeps since)
    if (sweep - synced >= G.sync_every) {
      rc = read_counts();
      if (rc)
        return rc;
      for (; synced < sweep; ++synced) {
        const int32_t* h = G.h_counts + RING_OFS + SLOT * (synced % RING);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, G.evx0[synced % RING], G.evx1[synced % RING]);
        ms_xs += ms;
        {
          const bool used[3] = {true, true, false};
          add_kernel_times(st, (int)(synced % RING), used);
        }
        n_xs += n_in_first;
        n_coll += h[0];
        n_cross += h[1];
        n_in_first = h[2];
      }
      n_alive = n_in_first;
      drain_score_times(st, G.overlap ? sweep - 2 : sweep);
    }
  }
  if (synced < sweep) { // cannot happen with the loop above, kept for safety
    rc = read_counts();
    if (rc)
      return rc;
  }
  for (int b = 0; b < 2; ++b)
    if (score_pending[b] && G.overlap)
      CUDA_OK(cudaStreamWaitEvent(G.stream, G.e_score[b], 0));
  if (n_alive > 0) { // the tail: one lane per remaining history
    a.q_in = G.q[cur];
    a.q_in.count = alive_count;
    a.fl = G.fl[0];
    a.fl_buf = 0;
    LAUNCH_TAIL(n_alive, G.stream, a);
    if (score && !G.tally_needs_cells) {
      rc = read_counts(); // flights the tail logged
      if (rc)
        return rc;
      a.n_flights = G.fl[0].count;
      a.cursor = nullptr;
      LAUNCH_SCORE(std::min<int64_t>(G.h_counts[6], G.fl[0].cap), G.stream, a);
    }
  }
  st.n_xs = n_xs;
  st.n_coll = n_coll;
  st.n_cross = n_cross;
  st.ms_xs = ms_xs;
  return 0;
}

// ---- sweep mode 2: paired lists ----------------------------------------------------
// Lists of "generation" g (g = 0 after the first flight): C[g & 1] = histories
// waiting for a collision, S[g & 1] = for a crossing; their counters live in ring
// slot g.  Sweep s turns the lists of generation s into those of s + 1 with
// k_collide_xa and k_surface_xa side by side; the flights they record go to
// fl[kind * 2 + parity] and are scored by k_score on its own stream while sweep
// s + 1 runs.
static int run_sweeps_paired(EvArgs& a, const omcgpu_gen_in* in, int64_t n, SweepStats& st)
{
  auto slot_of = [](int64_t g) { return G.d_counts + RING_OFS + SLOT * (g % RING); };
  const bool score = in->tallies_active && G.hm.n_tallies > 0;
  cudaStream_t st_score = G.overlap ? G.s_score : G.stream;
  cudaStream_t st_surf = G.overlap ? G.s_surf : G.stream;
  cudaStream_t st_coll = G.overlap ? G.s_coll : G.stream;
  bool score_pending[2] = {false, false};
  a.q_a = G.q_alive;
  a.q_a.count = G.d_counts + 0;
  OMC_LAUNCH(k_init, n, G.stream, a);
  int rc = read_counts();
  if (rc)
    return rc;
  const int64_t n_alive0 = G.h_counts[0];
  st.n_xs = n_alive0;
  // first flight: generation-0 lists
  CUDA_OK(cudaMemsetAsync(slot_of(0), 0, SLOT * sizeof(int32_t), G.stream));
  a.q_in = G.q_alive;
  a.q_in.count = G.d_counts + 0;
  a.q_a = G.q[0];
  a.q_a.count = slot_of(0) + 0;
  a.q_b = G.q[2];
  a.q_b.count = slot_of(0) + 1;
  a.fl = G.fl[0];
  a.fl_buf = 0;
  a.n_flights = G.d_counts + 0;
  a.ordered = G.ordered ? (G.keyed ? 2 : 1) : 0;
  if (G.ordered) // histories that never fly (zero weight, lost at birth) carry no flag
    CUDA_OK(cudaMemsetAsync(G.ps.next_ev, G.keyed ? FLAG_DEAD_KEYED : 0, (size_t)n, G.stream));
  CUDA_OK(cudaEventRecord(G.evx0[0], G.stream));
  LAUNCH_XS(k_xs_advance, n_alive0, G.stream, a);
  if (G.ordered) {
    rc = G.keyed ? build_lists_keyed(n, G.q[0].idx, slot_of(0), G.stream)
                 : build_lists(n, G.q[0].idx, G.q[2].idx, slot_of(0), G.stream);
    if (rc)
      return rc;
  }
  CUDA_OK(cudaEventRecord(G.evx1[0], G.stream));
  if (score) {
    if (G.overlap) {
      CUDA_OK(cudaEventRecord(G.e_adv, G.stream));
      CUDA_OK(cudaStreamWaitEvent(st_score, G.e_adv, 0));
    }
    CUDA_OK(cudaEventRecord(G.evk0[2][0], st_score));
    a.cursor = nullptr;
    LAUNCH_SCORE(n_alive0, st_score, a);
    CUDA_OK(cudaEventRecord(G.evk1[2][0], st_score));
    st.score_issued = 1;
    if (G.overlap) {
      CUDA_OK(cudaEventRecord(G.e_score[0], st_score));
      score_pending[0] = true;
      if (G.tally_needs_cells)
        CUDA_OK(cudaStreamWaitEvent(G.stream, G.e_score[0], 0));
    }
  }
  rc = read_counts();
  if (rc)
    return rc;
  {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, G.evx0[0], G.evx1[0]);
    st.ms_xs += ms;
  }
  int64_t sweep = 0;  // next sweep to issue = generation of its input lists
  int64_t known = 0;  // newest generation whose counters the host has read
  int64_t n_alive = (int64_t)G.h_counts[RING_OFS + 0] + G.h_counts[RING_OFS + 1];
  while (n_alive > 0 && n_alive > G.tail_threshold) {
    const int pi = (int)(sweep & 1), po = 1 - pi;
    const int ev = (int)((sweep + 1) % RING);
    int32_t* in_slot = slot_of(sweep);
    int32_t* out_slot = slot_of(sweep + 1);
    CUDA_OK(cudaMemsetAsync(out_slot, 0, SLOT * sizeof(int32_t), G.stream));
    if (score && score_pending[po] && G.overlap) // the records of sweep - 2 must have been consumed
      CUDA_OK(cudaStreamWaitEvent(G.stream, G.e_score[po], 0));
    CUDA_OK(cudaEventRecord(G.evx0[ev], G.stream));
    if (G.overlap)
      CUDA_OK(cudaEventRecord(G.e_adv, G.stream));
    a.q_a = G.q[po];
    a.q_a.count = out_slot + 0;
    a.q_b = G.q[2 + po];
    a.q_b.count = out_slot + 1;
    // crossings (keyed lists: behind the collisions in the same sorted array)
    a.q_in = G.q[G.keyed ? pi : 2 + pi];
    a.q_in.count = in_slot + 1;
    a.q_in.base = G.keyed ? in_slot + 0 : nullptr;
    a.fl = G.fl[0 * 2 + po];
    a.fl_buf = 0 * 2 + po;
    if (G.overlap)
      CUDA_OK(cudaStreamWaitEvent(st_surf, G.e_adv, 0));
    CUDA_OK(cudaEventRecord(G.evk0[0][ev], st_surf));
    a.cursor = in_slot + 4;
    LAUNCH_XA(k_surface_xa, n_alive, st_surf, a);
    CUDA_OK(cudaEventRecord(G.evk1[0][ev], st_surf));
    if (G.overlap)
      CUDA_OK(cudaEventRecord(G.e_surf, st_surf));
    // collisions
    a.q_in = G.q[pi];
    a.q_in.count = in_slot + 0;
    a.q_in.base = nullptr;
    a.fl = G.fl[1 * 2 + po];
    a.fl_buf = 1 * 2 + po;
    if (G.overlap)
      CUDA_OK(cudaStreamWaitEvent(st_coll, G.e_adv, 0));
    CUDA_OK(cudaEventRecord(G.evk0[1][ev], st_coll));
    a.cursor = in_slot + 5;
    LAUNCH_COLLISION_SCORE(n_alive, st_coll, a);
    LAUNCH_XA(k_collide_xa, n_alive, st_coll, a);
    LAUNCH_AFTER_COLLISION_SCORE(n_alive, st_coll, a);
    CUDA_OK(cudaEventRecord(G.evk1[1][ev], st_coll));
    if (G.overlap) {
      CUDA_OK(cudaEventRecord(G.e_coll, st_coll));
      CUDA_OK(cudaStreamWaitEvent(G.stream, G.e_surf, 0));
      CUDA_OK(cudaStreamWaitEvent(G.stream, G.e_coll, 0));
    }
    if (G.ordered) {
      rc = G.keyed ? build_lists_keyed(n, G.q[po].idx, out_slot, G.stream)
                   : build_lists(n, G.q[po].idx, G.q[2 + po].idx, out_slot, G.stream);
      if (rc)
        return rc;
    }
    CUDA_OK(cudaEventRecord(G.evx1[ev], G.stream));
    if (score) {
      if (G.overlap) {
        CUDA_OK(cudaStreamWaitEvent(st_score, G.e_surf, 0));
      }
      a.fl = G.fl[0 * 2 + po];
      a.fl_buf = 0 * 2 + po;
      a.n_flights = in_slot + 1;
      CUDA_OK(cudaEventRecord(G.evk0[2][ev], st_score));
      a.cursor = G.dynamic_score ? in_slot + 6 : nullptr;
      LAUNCH_SCORE(n_alive, st_score, a);
      CUDA_OK(cudaEventRecord(G.evk1[2][ev], st_score));
      if (G.overlap)
        CUDA_OK(cudaStreamWaitEvent(st_score, G.e_coll, 0));
      a.fl = G.fl[1 * 2 + po];
      a.fl_buf = 1 * 2 + po;
      a.n_flights = in_slot + 0;
      CUDA_OK(cudaEventRecord(G.evk0[3][ev], st_score));
      a.cursor = G.dynamic_score ? in_slot + 7 : nullptr;
      LAUNCH_SCORE(n_alive, st_score, a);
      CUDA_OK(cudaEventRecord(G.evk1[3][ev], st_score));
      st.score_issued = sweep + 2; // launch numbers 0 (first flight) .. sweep + 1
      if (G.overlap) {
        CUDA_OK(cudaEventRecord(G.e_score[po], st_score));
        score_pending[po] = true;
        if (G.tally_needs_cells) // cell / universe filters read the live coordinate levels
          CUDA_OK(cudaStreamWaitEvent(G.stream, G.e_score[po], 0));
      }
    }
    ++sweep;
    if (sweep - known >= G.sync_every) {
      rc = read_counts();
      if (rc)
        return rc;
      for (; known < sweep; ++known) { // account for sweep `known`: consumed generation known, made known + 1
        const int32_t* hin = G.h_counts + RING_OFS + SLOT * (known % RING);
        const int32_t* hout = G.h_counts + RING_OFS + SLOT * ((known + 1) % RING);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, G.evx0[(known + 1) % RING], G.evx1[(known + 1) % RING]);
        st.ms_xs += ms;
        {
          const bool used[3] = {true, true, false};
          add_kernel_times(st, (int)((known + 1) % RING), used);
        }
        st.n_coll += hin[0];
        st.n_cross += hin[1];
        if (G.trace)
          std::fprintf(G.trace, "%lld %lld %d %d\n", (long long)G.generation_no, (long long)known, hin[0], hin[1]);
        st.n_xs += (int64_t)hout[0] + hout[1];
        n_alive = (int64_t)hout[0] + hout[1];
      }
      // the main stream waited for scoring launch (sweep - 2) before it issued sweep - 1
      drain_score_times(st, G.overlap ? sweep - 2 : sweep + 1);
    }
  }
  for (int b = 0; b < 2; ++b)
    if (score_pending[b] && G.overlap)
      CUDA_OK(cudaStreamWaitEvent(G.stream, G.e_score[b], 0));
  if (n_alive > 0) { // the tail: one lane per remaining history, both lists into one flight log
    const int pi = (int)(sweep & 1);
    int32_t* in_slot = slot_of(sweep);
    a.fl = G.fl[0];
    a.fl_buf = 0;
    a.q_in = G.q[pi];
    a.q_in.count = in_slot + 0;
    a.q_in.base = nullptr;
    a.first_event = NEXT_COLLISION;
    LAUNCH_TAIL(n_alive, G.stream, a);
    a.q_in = G.q[G.keyed ? pi : 2 + pi];
    a.q_in.count = in_slot + 1;
    a.q_in.base = G.keyed ? in_slot + 0 : nullptr;
    a.first_event = NEXT_SURFACE;
    LAUNCH_TAIL(n_alive, G.stream, a);
    a.first_event = NEXT_NONE;
    if (score && !G.tally_needs_cells) {
      rc = read_counts(); // flights the tail logged
      if (rc)
        return rc;
      a.n_flights = G.fl[0].count;
      a.cursor = nullptr;
      LAUNCH_SCORE(std::min<int64_t>(G.h_counts[6], G.fl[0].cap), G.stream, a);
    }
  }
  return 0;
}

static int run_generation(const omcgpu_gen_in* in, omcgpu_gen_out* out)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  const int64_t n = in->n_particles;
  if (n > G.cap || n > G.n_src)
    return fail(OMCGPU_E_OUT_OF_BOUNDS, "n_particles exceeds the source bank / capacity%s");
  std::memset(out, 0, sizeof(*out));
  CUDA_OK(cudaMemsetAsync(G.d_counts, 0, N_COUNTS * sizeof(int32_t), G.stream));
  CUDA_OK(cudaMemsetAsync(G.d_counters, 0, N_COUNTERS * sizeof(unsigned long long), G.stream));
  if (G.d_sec_count)
    CUDA_OK(cudaMemsetAsync(G.d_sec_count, 0, 2 * sizeof(int32_t), G.stream));
  CUDA_OK(cudaEventRecord(G.ev0, G.stream));
  EvArgs a = make_args(in);
  SweepStats st;
  ++G.generation_no;
  int rc = G.sweep_mode == 2 ? run_sweeps_paired(a, in, n, st) : run_sweeps_merged(a, in, n, st);
  if (rc)
    return rc;
  const int64_t n_xs = st.n_xs, n_coll = st.n_coll, n_cross = st.n_cross;
  const double ms_xs = st.ms_xs;
  // death: deterministic reduction of the per-particle accumulators
  ReduceArgs r;
  r.ps = G.ps;
  r.src = G.src;
  r.n = n;
  r.partial = G.d_partial;
  r.chunk = (n + G.n_chunks - 1) / G.n_chunks;
  const int64_t used_chunks = (n + r.chunk - 1) / r.chunk;
  OMC_LAUNCH(k_reduce, used_chunks, G.stream, r);
  CUDA_OK(cudaMemcpyAsync(G.h_partial, G.d_partial, used_chunks * 6 * sizeof(double),
    cudaMemcpyDeviceToHost, G.stream));
  CUDA_OK(cudaMemcpyAsync(G.h_counters, G.d_counters, N_COUNTERS * sizeof(unsigned long long),
    cudaMemcpyDeviceToHost, G.stream));
  rc = read_counts();
  if (rc)
    return rc;
  double tot[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t c = 0; c < used_chunks; ++c)
    for (int k = 0; k < 6; ++k)
      tot[k] += G.h_partial[c * 6 + k];
  for (int k = 0; k < 4; ++k)
    out->global_tally[k] = tot[k];
  out->total_weight = tot[4];
  out->n_events = (int64_t)tot[5]; // (+ the events of particles that went on as a secondary, below)
  G.n_fis = std::min<int64_t>(G.h_counts[4], G.fis_cap);
  out->n_fission = G.n_fis;
  out->n_lost = G.h_counts[5];
  G.n_lost_last = out->n_lost;
  out->n_bank_full = G.h_counts[7];
  if (G.d_sec_count) { // a full secondary pool is an error, not a silent loss of neutrons
    int32_t sc[2] = {0, 0};
#if OMC_CUDA
    CUDA_OK(cudaMemcpy(sc, G.d_sec_count, sizeof(sc), cudaMemcpyDeviceToHost));
#else
    std::memcpy(sc, G.d_sec_count, sizeof(sc));
#endif
    if (sc[1] > 0)
      return fail(OMCGPU_E_OUT_OF_BOUNDS, "the secondary-neutron pool is full (set OMCGPU_SECONDARY_POOL to a larger number of sites)%s");
  }
  drain_score_times(st, st.score_issued); // the stream sync above covers every scoring launch
  out->ms_surface = st.ms_kernel[0];
  out->ms_collide = st.ms_kernel[1];
  out->ms_score = st.ms_kernel[2];
  out->n_launch_surface = st.n_launch[0];
  out->n_launch_collide = st.n_launch[1];
  out->n_launch_score = st.n_launch[2];
  for (int k = 0; k < 3; ++k) {
    out->lookups_done[k] = (int64_t)G.h_counters[CNT_SURFACE + 2 * k];
    out->nuclide_lookups[k] = (int64_t)G.h_counters[CNT_SURFACE + 2 * k + 1];
  }
  out->n_red = (int64_t)G.h_counters[CNT_RED];
  out->n_items_surface = st.n_cross;
  out->n_items_collide = st.n_coll;
  out->n_events += (int64_t)G.h_counters[0];
  out->n_xs_lookups = n_xs + (int64_t)G.h_counters[1];
  out->n_collisions = n_coll + (int64_t)G.h_counters[2];
  out->n_crossings = n_cross + (int64_t)G.h_counters[3];
  rc = sort_fission_bank(n, in->id_offset);
  if (rc)
    return rc;
  CUDA_OK(cudaEventRecord(G.ev1, G.stream));
  CUDA_OK(cudaStreamSynchronize(G.stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, G.ev0, G.ev1);
  out->ms_transport = ms;
  out->ms_xs = ms_xs;
  G.id_offset_last = in->id_offset;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(OMCGPU_E_CUDA, "kernel failure: %s", cudaGetErrorString(e));
  return 0;
}

} // namespace omc

// =============================================================================
// C ABI
// =============================================================================
using namespace omc;

extern "C" {

const char* omcgpu_last_error(void)
{
  return g_err;
}

int64_t omcgpu_kernel_launches(void)
{
  return g_launches;
}

int omcgpu_update(const omcgpu_model* model, uint32_t what)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  if (!model)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_update: no model%s");
  const omcgpu_model& h = G.hm; // scalars + device pointers of the model in use
  if (model->n_cells != h.n_cells || model->n_materials != G.n_base_materials)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_update: the model's shape differs from the one given to omcgpu_init%s");
  CUDA_OK(cudaStreamSynchronize(G.stream));
  if (!G.mat_base.empty()) {
    // several temperatures per nuclide: the device materials are (material, cell temperature) pairs
    if ((what & OMCGPU_UPDATE_CELL_TEMPERATURES) && G.temps_expanded)
      return fail(OMCGPU_E_UNSUPPORTED, "omcgpu_update: cell temperatures of a model with several temperatures per "
                                        "nuclide select other tables: call omcgpu_init again%s");
    if (what & OMCGPU_UPDATE_CELL_TEMPERATURES) {
      const int n = h.n_cells ? model->cell_sqrtkT_ofs[h.n_cells] : 0;
      if (n != G.n_cell_sqrtkT)
        return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_update: number of cell temperatures changed%s");
      CUDA_OK(cudaMemcpy(const_cast<double*>(h.cell_sqrtkT), model->cell_sqrtkT, n * sizeof(double), cudaMemcpyHostToDevice));
    }
    if (what & OMCGPU_UPDATE_MATERIAL_DENSITIES) {
      std::vector<double> dens;
      for (size_t d = 0; d < G.mat_base.size(); ++d) {
        const int b = G.mat_base[d];
        for (int k = model->mat_nuc_ofs[b]; k < model->mat_nuc_ofs[b + 1]; ++k)
          dens.push_back(model->mat_atom_density[k] * (d < G.mat_mult.size() ? G.mat_mult[d] : 1.0));
      }
      if ((int)dens.size() != G.n_mat_nuclides)
        return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_update: number of material nuclides changed%s");
      CUDA_OK(cudaMemcpy(const_cast<double*>(h.mat_atom_density), dens.data(), dens.size() * sizeof(double),
        cudaMemcpyHostToDevice));
    }
    if (what & ~(OMCGPU_UPDATE_CELL_TEMPERATURES | OMCGPU_UPDATE_MATERIAL_DENSITIES))
      return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_update: unknown property%s");
    return 0;
  }
  if (what & OMCGPU_UPDATE_CELL_TEMPERATURES) {
    const int n = h.n_cells ? model->cell_sqrtkT_ofs[h.n_cells] : 0;
    if (n != G.n_cell_sqrtkT)
      return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_update: number of cell temperatures changed%s");
    CUDA_OK(cudaMemcpy(const_cast<double*>(h.cell_sqrtkT), model->cell_sqrtkT, n * sizeof(double), cudaMemcpyHostToDevice));
  }
  if (what & OMCGPU_UPDATE_MATERIAL_DENSITIES) {
    const int n = G.n_base_materials ? model->mat_nuc_ofs[G.n_base_materials] : 0;
    if (n != G.n_mat_nuclides)
      return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_update: number of material nuclides changed%s");
    CUDA_OK(cudaMemcpy(const_cast<double*>(h.mat_atom_density), model->mat_atom_density, n * sizeof(double),
      cudaMemcpyHostToDevice));
  }
  if (what & ~(OMCGPU_UPDATE_CELL_TEMPERATURES | OMCGPU_UPDATE_MATERIAL_DENSITIES))
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_update: unknown property%s");
  return 0;
}

int omcgpu_sweep_kind(void)
{
  if (!G.ready)
    return -1;
  return G.sweep_mode == 2 ? 2 : (G.fuse_xs ? 1 : 0);
}

int omcgpu_count_fission_sites(int32_t n_dim, const int32_t* shape, const double* lower_left, const double* width,
  double* counts, int64_t n_bins, int32_t* sites_outside)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  if (n_dim < 1 || n_dim > 3 || !shape || !lower_left || !width || !counts || !sites_outside)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_count_fission_sites: bad arguments%s");
  int64_t bins = 1;
  for (int k = 0; k < n_dim; ++k)
    bins *= shape[k];
  if (bins != n_bins || bins < 1)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_count_fission_sites: n_bins does not match the mesh shape%s");
  CountArgs a;
  std::memset(&a, 0, sizeof(a));
  a.bank = G.fis;
  a.n = G.n_fis;
  a.nd = n_dim;
  for (int k = 0; k < n_dim; ++k) {
    a.shape[k] = shape[k];
    a.ll[k] = lower_left[k];
    a.w[k] = width[k];
  }
  double* d_counts = nullptr;
  int32_t* d_out = nullptr;
#if OMC_CUDA
  if (cudaMalloc((void**)&d_counts, bins * sizeof(double)) != cudaSuccess ||
      cudaMalloc((void**)&d_out, sizeof(int32_t)) != cudaSuccess) {
    cudaFree(d_counts);
    return fail(OMCGPU_E_ALLOCATE, "omcgpu_count_fission_sites: cudaMalloc failed%s");
  }
  cudaMemsetAsync(d_counts, 0, bins * sizeof(double), G.stream);
  cudaMemsetAsync(d_out, 0, sizeof(int32_t), G.stream);
#else
  d_counts = counts;
  std::memset(counts, 0, bins * sizeof(double));
  *sites_outside = 0;
  d_out = sites_outside;
#endif
  a.counts = d_counts;
  a.outside = d_out;
  OMC_LAUNCH(k_count_sites, a.n, G.stream, a);
#if OMC_CUDA
  cudaError_t e = cudaMemcpyAsync(counts, d_counts, bins * sizeof(double), cudaMemcpyDeviceToHost, G.stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(sites_outside, d_out, sizeof(int32_t), cudaMemcpyDeviceToHost, G.stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(G.stream);
  cudaFree(d_counts);
  cudaFree(d_out);
  if (e != cudaSuccess)
    return fail(OMCGPU_E_CUDA, "omcgpu_count_fission_sites: %s", cudaGetErrorString(e));
#endif
  return 0;
}

int64_t omcgpu_lost_particles(double* records, int64_t capacity)
{
  if (!G.ready)
    return -fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_lost_particles: library not initialised%s");
  if (!records && capacity > 0)
    return -fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_lost_particles: no buffer%s");
  const int64_t n = std::max<int64_t>(0, std::min<int64_t>(std::min<int64_t>(G.n_lost_last, OMCGPU_MAX_LOST_RECORDS), capacity));
  if (n > 0) {
#if OMC_CUDA
    if (cudaMemcpy(records, G.d_lost_rec, (size_t)n * OMCGPU_LOST_RECORD * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
      return -fail(OMCGPU_E_CUDA, "omcgpu_lost_particles: copy failed%s");
#else
    std::memcpy(records, G.d_lost_rec, (size_t)n * OMCGPU_LOST_RECORD * sizeof(double));
#endif
  }
  return n;
}

int omcgpu_set_option(const char* name, int64_t value)
{
  if (!name)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_set_option: no name%s");
  const std::string k(name);
  if (k == "count_lookups")
    G.count_lookups = value != 0;
  else if (k == "tail_threshold")
    G.tail_threshold = value;
  else if (k == "sync_every")
    G.sync_every = std::min<int64_t>(std::max<int64_t>(1, value), RING / 2);
  else if (k == "score_part")
    G.score_part = (int)std::max<int64_t>(1, value);
  else if (k == "score_version")
    g_score_version = value == 2 ? 2 : 1;
  else if (k == "probe_kT_neV") // cell temperature of the stand-alone lookup / collision probes, in 1e-9 eV
    G.probe_kT = 1e-9 * (double)value;
  else
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_set_option: unknown option %s", name);
  return 0;
}

int omcgpu_comm_finalize(void);

int omcgpu_finalize(void)
{
  omcgpu_comm_finalize(); // its staging buffers live in this engine's allocation list
  for (void* p : G.dev_allocs)
    cudaFree(p);
  G.dev_allocs.clear();
  if (G.h_counts)
    cudaFreeHost(G.h_counts);
  if (G.h_counters)
    cudaFreeHost(G.h_counters);
  if (G.h_partial)
    cudaFreeHost(G.h_partial);
  if (G.h_aos)
    cudaFreeHost(G.h_aos);
  {
    auto drop = [](cudaEvent_t e) {
      if (e)
        cudaEventDestroy(e);
    };
    drop(G.ev0);
    drop(G.ev1);
    for (int k = 0; k < RING; ++k) {
      drop(G.evx0[k]);
      drop(G.evx1[k]);
      for (int j = 0; j < 4; ++j) {
        drop(G.evk0[j][k]);
        drop(G.evk1[j][k]);
      }
    }
    drop(G.evp0);
    drop(G.evp1);
    drop(G.e_adv);
    drop(G.e_surf);
    drop(G.e_coll);
    drop(G.e_score[0]);
    drop(G.e_score[1]);
  }
  // streams and events are tracked one by one (not through G.ready): an omcgpu_init that
  // failed half way must not leak what it had created
  for (cudaStream_t st : {G.stream, G.s_score, G.s_surf, G.s_coll})
    if (st)
      cudaStreamDestroy(st);
  if (G.trace)
    std::fclose(G.trace);
  G = Engine();
  return 0;
}

static int init_impl(const omcgpu_model* model, int device, int64_t max_particles);

int omcgpu_init(const omcgpu_model* model, int device, int64_t max_particles)
{
  if (!model || max_particles <= 0)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_init: bad arguments%s");
  if (G.ready)
    omcgpu_finalize();
  const int rc = init_impl(model, device, max_particles);
  if (rc) { // release whatever the failed attempt had allocated; the error text survives
    char keep[sizeof(g_err)];
    std::memcpy(keep, g_err, sizeof(keep));
    omcgpu_finalize();
    std::memcpy(g_err, keep, sizeof(keep));
  }
  return rc;
}

static int init_impl(const omcgpu_model* model, int device, int64_t max_particles)
{
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(OMCGPU_E_CUDA, "no CUDA device available: the device path has no CPU fallback%s");
  if (device < 0 || device >= n_dev)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_init: device index out of range%s");
  CUDA_OK(cudaSetDevice(device));
  int rc = check_supported(model);
  if (rc)
    return rc;
  if (max_particles > 2000000000LL / 3)
    return fail(OMCGPU_E_OUT_OF_BOUNDS, "max_particles too large for 32-bit work lists%s");
#if OMC_CUDA
  {
    // the event kernels are the critical path of a sweep; k_score only has to be done
    // two sweeps later, so its blocks take the SM slots the event kernels leave
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    const char* t = std::getenv("OMCGPU_PRIORITY");
    const bool use = t && std::atoi(t) != 0; // measured: no gain on C2 (profiles/README.md r10), default off
    CUDA_OK(cudaStreamCreateWithPriority(&G.stream, cudaStreamDefault, use ? prio_hi : prio_lo));
    CUDA_OK(cudaStreamCreateWithPriority(&G.s_surf, cudaStreamDefault, use ? prio_hi : prio_lo));
    CUDA_OK(cudaStreamCreateWithPriority(&G.s_coll, cudaStreamDefault, use ? prio_hi : prio_lo));
    CUDA_OK(cudaStreamCreateWithPriority(&G.s_score, cudaStreamDefault, prio_lo));
  }
#else
  CUDA_OK(cudaStreamCreate(&G.stream));
  CUDA_OK(cudaStreamCreate(&G.s_score));
  CUDA_OK(cudaStreamCreate(&G.s_surf));
  CUDA_OK(cudaStreamCreate(&G.s_coll));
#endif
  if (const char* t = std::getenv("OMCGPU_FUSE_XS"))
    G.fuse_xs = std::atoi(t) != 0;
  if (const char* t = std::getenv("OMCGPU_OVERLAP"))
    G.overlap = std::atoi(t) != 0;
  if (const char* t = std::getenv("OMCGPU_SWEEP"))
    G.sweep_mode = std::atoi(t) == 1 ? 1 : 2;
  G.batch_xs = false;
  if (model->run_CE)
    for (int m = 0; m < model->n_materials; ++m)
      if (model->mat_nuc_ofs[m + 1] - model->mat_nuc_ofs[m] >= OMC_XS_BATCH_MIN)
        G.batch_xs = true;
  // many-nuclide materials (tables far larger than L2): the lookup runs as its own kernel
  // over the work list sorted by (material, energy) -- measured on C5: lookup kernel 2x,
  // whole step +35 % (profiles/r25_matrix.log); with 1-5 nuclides per material (C2, C3)
  // the fused, index-ordered sweeps are faster
  G.sort_xs = G.batch_xs;
  if (const char* t = std::getenv("OMCGPU_SORT_XS"))
    G.sort_xs = std::atoi(t) != 0;
  if (const char* t = std::getenv("OMCGPU_SORT_COLL"))
    G.sort_coll = std::atoi(t) != 0;
  if (G.sort_xs) { // needs the stand-alone lookup kernel over a merged alive list
    G.sweep_mode = 1;
    G.fuse_xs = false;
  }
  if (const char* t = std::getenv("OMCGPU_LOW_COUNT"))
    G.low_count = std::atoll(t);
  if (const char* t = std::getenv("OMCGPU_BATCH_XS"))
    G.batch_xs = std::atoi(t) != 0;
  if (const char* t = std::getenv("OMCGPU_ORDERED"))
    G.ordered = std::atoi(t) != 0;
  if (const char* t = std::getenv("OMCGPU_SCORE_CTAS"))
    G.score_ctas = std::max(0, std::atoi(t));
#if OMC_CUDA
  {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess)
      G.n_sm = prop.multiProcessorCount;
    g_score_grid_cap = G.score_ctas * G.n_sm;
  }
#endif
  // particle state is gathered in 32-byte sectors (PState): do not let L2 fetch 64
  {
    int gran = 32;
    if (const char* t = std::getenv("OMCGPU_L2_FETCH"))
      gran = std::atoi(t);
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
  }
  g_arena = nullptr;
  g_arena_size = g_arena_used = 0;
  rc = upload_model(model);
  if (rc)
    return rc;
#if OMC_CUDA
  if (const char* t = std::getenv("OMCGPU_L2_PERSIST")) {
    if (std::atoi(t) != 0 && g_arena && g_arena_used > 0) {
      cudaDeviceProp prop;
      cudaGetDeviceProperties(&prop, device);
      const size_t want = std::min<size_t>(g_arena_used, (size_t)prop.persistingL2CacheMaxSize);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      cudaStreamAttrValue attr;
      std::memset(&attr, 0, sizeof(attr));
      attr.accessPolicyWindow.base_ptr = g_arena;
      attr.accessPolicyWindow.num_bytes = std::min<size_t>(g_arena_used, (size_t)prop.accessPolicyMaxWindowSize);
      attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
      attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      for (cudaStream_t st : {G.stream, G.s_surf, G.s_coll})
        cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
      std::fprintf(stderr, "omcgpu: L2 persisting window %zu MB of %zu MB tables (max %d MB)\n", want >> 20,
        g_arena_used >> 20, prop.persistingL2CacheMaxSize >> 20);
    }
  }
#endif
  if (const char* t = std::getenv("OMCGPU_TAIL_THRESHOLD"))
    G.tail_threshold = std::atoll(t);
  if (const char* t = std::getenv("OMCGPU_TRACE"))
    G.trace = std::fopen(t, "a");
  if (const char* t = std::getenv("OMCGPU_SCORE_VERSION"))
    g_score_version = std::atoi(t) == 2 ? 2 : 1;
  G.score_part = g_score_version == 1 ? 16 : SCORE_PART_V2;
  if (const char* t = std::getenv("OMCGPU_SCORE_PART"))
    G.score_part = std::max(1, std::atoi(t));
  const int64_t cap = max_particles;
  G.cap = cap;
  G.fis_cap = 3 * cap; // init_fission_bank(3 * work_per_rank), src/simulation.cpp:340
  const int L = std::max(1, std::min<int>(model->n_coord_levels, OMCGPU_MAX_COORD));
  PState& ps = G.ps;
  ps.cap = cap;
  ps.id_shift = 0;
  ps.is_local = 0;
  ps.n_prefix = 0;
  ps.xs_prefix = nullptr;
  if (model->run_CE && G.batch_xs) {
    int most = 0;
    for (int m = 0; m < model->n_materials; ++m)
      most = std::max(most, model->mat_nuc_ofs[m + 1] - model->mat_nuc_ofs[m]);
    if (most >= 2 * XS_PREFIX_STRIDE) {
      bool on = true;
      if (const char* t = std::getenv("OMCGPU_XS_PREFIX"))
        on = std::atoi(t) != 0;
      if (on) {
        ps.n_prefix = most / XS_PREFIX_STRIDE;
        rc = rc ? rc : dev_alloc(&ps.xs_prefix, (size_t)cap * ps.n_prefix);
      }
    }
  }
  {
    // sector groups (see PState): one allocation per group, fields at fixed slots
    char *gA = nullptr, *gB = nullptr, *gC = nullptr, *p1 = nullptr, *p2 = nullptr, *p3 = nullptr,
         *p4 = nullptr, *p5 = nullptr;
    rc = rc ? rc : dev_alloc(&gA, (size_t)cap * L * 32);
    rc = rc ? rc : dev_alloc(&gB, (size_t)cap * L * 32);
    rc = rc ? rc : dev_alloc(&gC, (size_t)cap * L * 16);
    rc = rc ? rc : dev_alloc(&p1, (size_t)cap * 32);
    rc = rc ? rc : dev_alloc(&p2, (size_t)cap * 32);
    rc = rc ? rc : dev_alloc(&p3, (size_t)cap * 32);
    rc = rc ? rc : dev_alloc(&p4, (size_t)cap * 32);
    rc = rc ? rc : dev_alloc(&p5, (size_t)cap * 32);
    rc = rc ? rc : dev_alloc(&ps.n_progeny, (size_t)cap);
    rc = rc ? rc : dev_alloc(&ps.next_ev, (size_t)cap);
    if (rc)
      return rc;
#define D(col, base, slot) ps.col.p = reinterpret_cast<double*>(base) + (slot)
#define I(col, base, slot) ps.col.p = reinterpret_cast<int32_t*>(base) + (slot)
    D(rx, gA, 0); D(ry, gA, 1); D(rz, gA, 2); I(cell, gA, 6); I(lattice, gA, 7);
    D(ux, gB, 0); D(uy, gB, 1); D(uz, gB, 2); I(universe, gB, 6);
    I(lat_x, gC, 0); I(lat_y, gC, 1); I(lat_z, gC, 2);
    D(E, p1, 0); D(wgt, p1, 1); D(time, p1, 2);
    ps.seed.p = reinterpret_cast<uint64_t*>(p1) + 3;
    D(xs_total, p2, 0); D(xs_absorption, p2, 1); D(xs_fission, p2, 2); D(xs_nu_fission, p2, 3);
    D(k_abs, p3, 0); D(k_col, p3, 1); D(k_trk, p3, 2); D(k_leak, p3, 3);
    I(n_coord, p4, 0); I(surface, p4, 1); I(material, p4, 2); I(material_last, p4, 3);
    I(cell_instance, p4, 4); I(cell_last0, p4, 5); I(g, p4, 6); I(n_event, p4, 7);
    D(sqrtkT, p5, 0); D(sqrtkT_last, p5, 1);
    I(b_surface, p5, 4); I(b_level_trans, p5, 5); I(g_last, p5, 6);
#undef D
#undef I
  }
  // A rank's comb may select more sites than its share of the next generation (they
  // travel to their owners afterwards, src/eigenvalue.cpp:186-348): the source bank
  // and the gather list hold 3 x max_particles sites, like the reference's temp_sites
  // (src/eigenvalue.cpp:135).
  const int64_t src_cap = 3 * cap;
  rc = rc ? rc : alloc_bank(G.src, src_cap);
  rc = rc ? rc : alloc_bank(G.fis, G.fis_cap);
  rc = rc ? rc : alloc_bank(G.fis_sorted, G.fis_cap);
  rc = rc ? rc : dev_alloc(&G.d_counts, N_COUNTS);
  rc = rc ? rc : dev_alloc(&G.d_counters, N_COUNTERS);
  for (int k = 0; k < 4; ++k) {
    rc = rc ? rc : dev_alloc(&G.q[k].idx, cap);
    G.q[k].count = G.d_counts + k;
  }
  rc = rc ? rc : dev_alloc(&G.q_alive.idx, cap);
  G.q_alive.count = G.d_counts + 0;
  if (G.sort_xs) {
    rc = rc ? rc : dev_alloc(&G.d_keys, cap);
    rc = rc ? rc : dev_alloc(&G.d_keys_out, cap);
    rc = rc ? rc : dev_alloc(&G.d_vals, cap);
    rc = rc ? rc : dev_alloc(&G.d_vals_out, cap);
#if OMC_CUDA
    G.sort_tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, G.sort_tmp_bytes, G.d_keys, G.d_keys_out, G.d_vals, G.d_vals_out,
      (int)cap, 0, 32);
    char* stmp = nullptr;
    rc = rc ? rc : dev_alloc(&stmp, G.sort_tmp_bytes + 16);
    G.d_sort_tmp = stmp;
#endif
  }
  if (const char* t = std::getenv("OMCGPU_DYNAMIC"))
    G.dynamic = std::atoi(t) != 0;
  if (const char* t = std::getenv("OMCGPU_DYNAMIC_SCORE"))
    G.dynamic_score = std::atoi(t) != 0;
  if (const char* t = std::getenv("OMCGPU_KEYED"))
    G.keyed = std::atoi(t) != 0;
  if (G.keyed && (G.sweep_mode != 2 || !G.ordered))
    G.keyed = false;
  if (G.keyed) {
    rc = rc ? rc : dev_alloc(&G.d_key_out, cap);
    rc = rc ? rc : dev_alloc(&G.d_iota, cap);
#if OMC_CUDA
    G.ksort_tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, G.ksort_tmp_bytes, G.ps.next_ev, G.d_key_out, G.d_iota, G.d_iota, (int)cap, 0, 8);
    char* ktmp = nullptr;
    rc = rc ? rc : dev_alloc(&ktmp, G.ksort_tmp_bytes + 16);
    G.d_ksort_tmp = ktmp;
#endif
    if (!rc) {
      IotaArgs ia;
      ia.v = G.d_iota;
      ia.n = cap;
      OMC_LAUNCH(k_iota, cap, G.stream, ia);
    }
  }
  rc = rc ? rc : dev_alloc(&G.d_offs, cap + 1);
  rc = rc ? rc : dev_alloc(&G.d_gather, src_cap);
  rc = rc ? rc : dev_alloc(&G.d_aos, G.fis_cap);
  {
    // replicate while all copies together stay L2-resident (<= 4 M doubles)
    int r = 8;
    if (const char* t = std::getenv("OMCGPU_TALLY_REPLICAS"))
      r = std::max(1, std::atoi(t));
    while (r > 1 && (int64_t)r * G.n_tally > (4 << 20))
      r >>= 1;
    G.tally_replicas = G.n_tally > 0 ? r : 1;
    G.tally_stride = (G.n_tally + 15) / 16 * 16;
  }
  rc = rc ? rc : dev_alloc(&G.d_tally, (size_t)G.tally_stride * G.tally_replicas);
  rc = rc ? rc : dev_alloc(&G.d_lost_rec, (size_t)OMCGPU_MAX_LOST_RECORDS * OMCGPU_LOST_RECORD);
  G.d_sec_site = nullptr;
  G.d_sec_next = G.d_sec_head = G.d_sec_count = nullptr;
  G.sec_cap = 0;
  if (model_creates_secondaries(model)) {
    G.sec_cap = (int32_t)std::min<int64_t>(std::max<int64_t>(1 << 16, cap / 2), 1 << 30);
    if (const char* t = std::getenv("OMCGPU_SECONDARY_POOL"))
      G.sec_cap = std::max(1, std::atoi(t));
    rc = rc ? rc : dev_alloc(&G.d_sec_site, (size_t)G.sec_cap * SEC_ND);
    rc = rc ? rc : dev_alloc(&G.d_sec_next, (size_t)G.sec_cap);
    rc = rc ? rc : dev_alloc(&G.d_sec_head, (size_t)cap);
    rc = rc ? rc : dev_alloc(&G.d_sec_count, 2);
  }
  G.d_an_d = nullptr;
  G.d_an_i = nullptr;
  if (G.n_after_collision_tallies > 0) { // collision records for the tallies scored after a collision
    rc = rc ? rc : dev_alloc(&G.d_an_d, (size_t)AN_ND * cap);
    rc = rc ? rc : dev_alloc(&G.d_an_i, (size_t)AN_NI * cap);
  }
  G.n_chunks = 4096;
  rc = rc ? rc : dev_alloc(&G.d_partial, G.n_chunks * 6);
  if (rc)
    return rc;
#if OMC_CUDA
  G.scan_tmp_bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, G.scan_tmp_bytes, G.ps.n_progeny, G.d_offs, (int)cap);
  char* tmp = nullptr;
  rc = dev_alloc(&tmp, G.scan_tmp_bytes);
  if (rc)
    return rc;
  G.d_scan_tmp = tmp;
  {
    FlagIs f1{nullptr, 1}, f2{nullptr, 2};
    thrust::counting_iterator<int32_t> ids(0);
    thrust::discard_iterator<> none;
    G.part_tmp_bytes = 0;
    cub::DevicePartition::If(nullptr, G.part_tmp_bytes, ids, (int32_t*)nullptr, (int32_t*)nullptr, none,
      (int32_t*)nullptr, (int)cap, f1, f2);
    char* ptmp = nullptr;
    rc = dev_alloc(&ptmp, G.part_tmp_bytes + 16);
    if (rc)
      return rc;
    G.d_part_tmp = ptmp;
  }
#endif
  {
    // flight records: one per work-list position in a sweep; the tail kernel logs
    // up to 64 flights per history it takes over (beyond that it scores inline)
    for (int b = 0; b < 4; ++b) {
      Flights& fl = G.fl[b];
      fl.cap = std::max<int64_t>(cap, b == 0 ? 64 * std::min<int64_t>(cap, G.tail_threshold) : 0);
#define F(col) rc = rc ? rc : dev_alloc(&fl.col, (size_t)fl.cap)
      F(x); F(y); F(z); F(ux); F(uy); F(uz); F(total); F(flux); F(E);
      F(xs_t); F(xs_a); F(xs_f); F(xs_nf); F(material); F(g); F(mg_ba); F(pidx);
#undef F
      fl.count = G.d_counts + 6;
    }
    if (rc)
      return rc;
    G.tally_needs_cells = false;
    for (int f = 0; f < model->n_filters; ++f)
      if (model->filter_type[f] == OMCGPU_FILTER_CELL || model->filter_type[f] == OMCGPU_FILTER_UNIVERSE ||
          model->filter_type[f] == OMCGPU_FILTER_DISTRIBCELL)
        G.tally_needs_cells = true;
  }
#if OMC_CUDA
  {
    int pf = 0; // measured: no gain on C2 (profiles/README.md r10), default off
    if (const char* t = std::getenv("OMCGPU_PREFETCH"))
      pf = std::atoi(t) != 0;
    CUDA_OK(cudaMemcpyToSymbol(c_prefetch, &pf, sizeof(int)));
  }
  CUDA_OK(cudaMemcpyToSymbol(c_ps, &G.ps, sizeof(PState)));
  CUDA_OK(cudaMemcpyToSymbol(c_fl, G.fl, 4 * sizeof(Flights)));
#endif
  CUDA_OK(cudaMallocHost((void**)&G.h_counts, N_COUNTS * sizeof(int32_t)));
  CUDA_OK(cudaMallocHost((void**)&G.h_counters, N_COUNTERS * sizeof(unsigned long long)));
  CUDA_OK(cudaMallocHost((void**)&G.h_partial, G.n_chunks * 6 * sizeof(double)));
  CUDA_OK(cudaMallocHost((void**)&G.h_aos, G.fis_cap * sizeof(omcgpu_site)));
  CUDA_OK(cudaEventCreate(&G.ev0));
  CUDA_OK(cudaEventCreate(&G.ev1));
  for (int k = 0; k < RING; ++k) {
    CUDA_OK(cudaEventCreate(&G.evx0[k]));
    CUDA_OK(cudaEventCreate(&G.evx1[k]));
    for (int j = 0; j < 4; ++j) {
      CUDA_OK(cudaEventCreate(&G.evk0[j][k]));
      CUDA_OK(cudaEventCreate(&G.evk1[j][k]));
    }
  }
  CUDA_OK(cudaEventCreate(&G.evp0));
  CUDA_OK(cudaEventCreate(&G.evp1));
  if (const char* t = std::getenv("OMCGPU_SYNC_EVERY"))
    G.sync_every = std::min<int64_t>(std::max<int64_t>(1, std::atoll(t)), RING / 2);
  CUDA_OK(cudaEventCreateWithFlags(&G.e_adv, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&G.e_surf, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&G.e_coll, cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&G.e_score[0], cudaEventDisableTiming));
  CUDA_OK(cudaEventCreateWithFlags(&G.e_score[1], cudaEventDisableTiming));
  G.ready = true;
  g_launches = 0;
  return 0;
}

// a host buffer the CUDA driver knows (cudaMallocHost / omcgpu_host_register) is
// copied straight from / to; anything else goes through the pinned staging area
static bool is_pinned(const void* p)
{
#if OMC_CUDA
  cudaPointerAttributes attr;
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeHost;
#else
  (void)p;
  return false;
#endif
}

int omcgpu_host_register(void* p, int64_t bytes)
{
#if OMC_CUDA
  if (!p || bytes <= 0)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_host_register: bad arguments%s");
  cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterDefault);
  if (e != cudaSuccess)
    return fail(OMCGPU_E_CUDA, "cudaHostRegister failed: %s", cudaGetErrorString(e));
#else
  (void)p;
  (void)bytes;
#endif
  return 0;
}

int omcgpu_host_unregister(void* p)
{
#if OMC_CUDA
  cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess)
    return fail(OMCGPU_E_CUDA, "cudaHostUnregister failed: %s", cudaGetErrorString(e));
#else
  (void)p;
#endif
  return 0;
}

// H2D of a SourceSite[] + unpack into the SoA source bank; asynchronous on
// G.stream when the buffer is pinned (the caller synchronises before returning)
static int set_source_bank_async(const omcgpu_site* source_bank, int64_t n)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  if (n > G.cap)
    return fail(OMCGPU_E_OUT_OF_BOUNDS, "source bank larger than max_particles%s");
  const omcgpu_site* from = source_bank;
  if (!is_pinned(source_bank)) {
    std::memcpy(G.h_aos, source_bank, n * sizeof(omcgpu_site));
    from = G.h_aos;
  }
  CUDA_OK(cudaMemcpyAsync(G.d_aos, from, n * sizeof(omcgpu_site), cudaMemcpyHostToDevice, G.stream));
  BankArgs b;
  std::memset(&b, 0, sizeof(b));
  b.a = G.src;
  b.aos = G.d_aos;
  b.n = n;
  OMC_LAUNCH(k_unpack, n, G.stream, b);
  G.n_src = n;
  return 0;
}

int omcgpu_set_source_bank(const omcgpu_site* source_bank, int64_t n)
{
  int rc = set_source_bank_async(source_bank, n);
  if (rc)
    return rc;
  CUDA_OK(cudaStreamSynchronize(G.stream));
  return 0;
}

static int fetch_bank(const Bank& bank, omcgpu_site* out, int64_t n)
{
  BankArgs b;
  std::memset(&b, 0, sizeof(b));
  b.a = bank;
  b.aos = G.d_aos;
  b.n = n;
  OMC_LAUNCH(k_pack, n, G.stream, b);
  const bool direct = is_pinned(out);
  CUDA_OK(cudaMemcpyAsync(direct ? out : G.h_aos, G.d_aos, n * sizeof(omcgpu_site), cudaMemcpyDeviceToHost,
    G.stream));
  CUDA_OK(cudaStreamSynchronize(G.stream));
  if (!direct)
    std::memcpy(out, G.h_aos, n * sizeof(omcgpu_site));
  return 0;
}

int omcgpu_get_source_bank(omcgpu_site* source_bank, int64_t n)
{
  if (!G.ready || n > G.n_src)
    return fail(OMCGPU_E_OUT_OF_BOUNDS, "omcgpu_get_source_bank: bad size%s");
  return fetch_bank(G.src, source_bank, n);
}

int omcgpu_get_fission_bank(omcgpu_site* fission_bank, int64_t capacity, int64_t* n)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  const int64_t m = std::min<int64_t>(G.n_fis, capacity);
  if (n)
    *n = m;
  return fetch_bank(G.fis_sorted, fission_bank, m);
}

int omcgpu_run_generation(const omcgpu_gen_in* in, omcgpu_gen_out* out)
{
  return run_generation(in, out);
}

// host bank -> device unless the caller vouches that the device copy is current
static int upload_source(const omcgpu_gen_in* in, const omcgpu_site* source_bank)
{
  if (in->source_bank_clean) {
    if (in->n_particles > G.n_src)
      return fail(OMCGPU_E_OUT_OF_BOUNDS, "source_bank_clean: the device bank holds fewer sites than n_particles%s");
    return 0;
  }
  return set_source_bank_async(source_bank, in->n_particles);
}

int omcgpu_transport_generation(const omcgpu_gen_in* in, const omcgpu_site* source_bank,
  omcgpu_site* fission_bank, int64_t fission_capacity, omcgpu_gen_out* out)
{
  int rc = upload_source(in, source_bank);
  if (rc)
    return rc;
  rc = run_generation(in, out);
  if (rc)
    return rc;
  int64_t n = 0;
  rc = omcgpu_get_fission_bank(fission_bank, fission_capacity, &n);
  out->n_fission = n;
  return rc;
}

int omcgpu_comb_fission_bank(int64_t start, int64_t total, int64_t n_particles, int64_t seed_id,
  int64_t master_seed, uint64_t stride, int64_t* n_sampled);

int omcgpu_next_generation(const omcgpu_gen_in* in, omcgpu_site* source_bank, int64_t seed_id,
  omcgpu_gen_out* out)
{
  int rc = upload_source(in, source_bank);
  if (rc)
    return rc;
  rc = run_generation(in, out);
  if (rc)
    return rc;
  if (out->n_fission == 0)
    return fail(OMCGPU_E_PHYSICS, "no fission sites banked%s");
  int64_t n_sampled = 0;
  rc = omcgpu_comb_fission_bank(0, out->n_fission, in->n_particles, seed_id, in->master_seed, in->stride,
    &n_sampled);
  if (rc)
    return rc;
  if (n_sampled != in->n_particles)
    return fail(OMCGPU_E_PHYSICS, "comb did not select n_particles sites%s");
  return fetch_bank(G.src, source_bank, n_sampled);
}

// The comb's tooth positions are a floating-point recurrence in the reference,
// tooth += teeth_distance (src/eigenvalue.cpp:166-177), and floor(tooth) picks the site, so
// the very same roundings must be reproduced -- but not one dependent add per tooth (10 ms
// per generation at 4e6 sites).  While tooth stays inside one binade [2^k, 2^(k+1)) every
// tooth is a multiple of u = 2^(k-52) and fl(tooth + d) = tooth + D u exactly, D = d / u
// rounded to nearest (no tie, exact sum below 2^(k+1)): the teeth of a binade are an integer
// progression, filled by a loop without a dependency chain; one real addition crosses into
// the next binade.  Anything irregular (tie, d >= tooth, zero) takes the plain recurrence.
static void comb_teeth(double tooth, double d, int64_t count, int64_t start, int32_t* idx)
{
  const int64_t LIMIT = (int64_t)1 << 53;
  int64_t i = 0;
  while (i < count) {
    int e = 0;
    if (tooth > 0.0)
      std::frexp(tooth, &e);
    const int k = e - 1, sh = 52 - k; // tooth in [2^k, 2^(k+1)), floor = mantissa integer >> sh
    bool block = tooth > 0.0 && d > 0.0 && sh >= 1 && sh <= 62 && d < std::ldexp(1.0, k);
    int64_t D = 0, T = 0;
    if (block) {
      const double q = std::ldexp(d, sh); // d / u, exact
      const double qr = std::nearbyint(q);
      block = std::fabs(q - std::floor(q) - 0.5) != 0.0 && qr >= 1.0 && qr < 4503599627370496.0;
      D = (int64_t)qr;
      T = (int64_t)std::ldexp(tooth, sh);
    }
    if (!block) {
      idx[i++] = (int32_t)((int64_t)std::floor(tooth) - start);
      tooth += d;
      continue;
    }
    // teeth j = 0..m of this block: T + j D; the exact sum of the next add stays below 2^(k+1)
    int64_t m = (LIMIT - 1 - T) / D;
    if (m > count - i - 1)
      m = count - i - 1;
    int32_t* out = idx + i;
    int64_t Tj = T; // (integer add: a one-cycle chain instead of the 4-cycle floating-point one)
    for (int64_t j = 0; j <= m; ++j) {
      out[j] = (int32_t)((Tj >> sh) - start);
      Tj += D;
    }
    i += m + 1;
    tooth = std::ldexp((double)(T + m * D), -sh); // the last tooth filled, then one real addition
    tooth += d;
  }
}

// Host half of the comb (needs no device): which of this rank's n_local sorted sites the
// uniform comb selects, as indices relative to `start`.
int omcgpu_comb_indices(int64_t start, int64_t n_local, int64_t total, int64_t n_particles, int64_t seed_id,
  int64_t master_seed, uint64_t stride, int32_t* idx, int64_t capacity, int64_t* n_sampled)
{
  uint64_t seed = future_seed((uint64_t)seed_id * stride, (uint64_t)master_seed); // init_seed(id, STREAM_TRACKING)
  const double teeth_distance = static_cast<double>(total) / n_particles;
  const double teeth_offset = prn(seed) * teeth_distance;
  const int64_t end = start + n_local;
  const int64_t tooth_start = (int64_t)std::ceil((start - teeth_offset) / teeth_distance);
  const int64_t tooth_end = (int64_t)std::floor((end - teeth_offset) / teeth_distance) + 1;
  const int64_t count = tooth_end - tooth_start > 0 ? tooth_end - tooth_start : 0;
  if (n_sampled)
    *n_sampled = count;
  if (count > capacity)
    return fail(OMCGPU_E_OUT_OF_BOUNDS, "comb selects more sites than the caller's buffer holds%s");
  comb_teeth(tooth_start * teeth_distance + teeth_offset, teeth_distance, count, start, idx);
  return 0;
}

int omcgpu_comb_fission_bank(int64_t start, int64_t total, int64_t n_particles, int64_t seed_id,
  int64_t master_seed, uint64_t stride, int64_t* n_sampled)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  // the host finds the teeth (comb_teeth: the reference's recurrence, binade by binade),
  // the device does the gather
  int32_t* idx = reinterpret_cast<int32_t*>(G.h_aos);
  int64_t count = 0;
  if (omcgpu_comb_indices(start, G.n_fis, total, n_particles, seed_id, master_seed, stride, idx, G.src.cap, &count))
    return fail(OMCGPU_E_OUT_OF_BOUNDS, "comb selects more sites than the source bank holds (3 max_particles)%s");
  if (count > 0) {
    CUDA_OK(cudaMemcpyAsync(G.d_gather, idx, count * sizeof(int32_t), cudaMemcpyHostToDevice, G.stream));
    BankArgs b;
    std::memset(&b, 0, sizeof(b));
    b.a = G.fis_sorted;
    b.b = G.src;
    b.gather = G.d_gather;
    b.n = count;
    OMC_LAUNCH(k_gather_bank, count, G.stream, b);
    CUDA_OK(cudaStreamSynchronize(G.stream));
  }
  G.n_src = count;
  if (n_sampled)
    *n_sampled = count;
  return 0;
}

int omcgpu_bank_columns_device(int which, void** cols, int64_t* n, int64_t* capacity)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  const Bank& b = which == 0 ? G.src : G.fis_sorted;
  cols[0] = b.x; cols[1] = b.y; cols[2] = b.z; cols[3] = b.u; cols[4] = b.v; cols[5] = b.w;
  cols[6] = b.E; cols[7] = b.time; cols[8] = b.wgt; cols[9] = b.delayed_group;
  cols[10] = b.parent_id; cols[11] = b.progeny_id;
  if (n)
    *n = which == 0 ? G.n_src : G.n_fis;
  if (capacity)
    *capacity = b.cap;
  return 0;
}

int omcgpu_set_source_count(int64_t n)
{
  if (!G.ready || n > G.cap)
    return fail(OMCGPU_E_OUT_OF_BOUNDS, "omcgpu_set_source_count: bad size%s");
  G.n_src = n;
  return 0;
}

int omcgpu_tally_results(int32_t tally, double* value, int64_t n)
{
  if (!G.ready || tally < 0 || tally >= G.hm.n_tallies)
    return fail(OMCGPU_E_OUT_OF_BOUNDS, "omcgpu_tally_results: bad tally index%s");
  const std::vector<int64_t>& ofs = G.tally_ofs;
  const int64_t len = ofs[tally + 1] - ofs[tally];
  if (n != len)
    return fail(OMCGPU_E_INVALID_ARGUMENT, "omcgpu_tally_results: size mismatch%s");
  if (G.tally_replicas > 1) {
    FoldArgs f;
    f.v = G.d_tally + ofs[tally];
    f.n = len;
    f.stride = G.tally_stride;
    f.replicas = G.tally_replicas;
    OMC_LAUNCH(k_fold_tally, len, G.stream, f);
  }
  // through the page-locked staging area, in pieces, all on the engine's stream
  double* stage = reinterpret_cast<double*>(G.h_aos);
  const int64_t piece = std::max<int64_t>(1, (int64_t)(G.fis_cap * sizeof(omcgpu_site) / sizeof(double)));
  for (int64_t done = 0; done < len; done += piece) {
    const int64_t m = std::min<int64_t>(piece, len - done);
    CUDA_OK(cudaMemcpyAsync(stage, G.d_tally + ofs[tally] + done, m * sizeof(double), cudaMemcpyDeviceToHost, G.stream));
    CUDA_OK(cudaMemsetAsync(G.d_tally + ofs[tally] + done, 0, m * sizeof(double), G.stream));
    CUDA_OK(cudaStreamSynchronize(G.stream));
    for (int64_t i = 0; i < m; ++i)
      value[done + i] += stage[i];
  }
  return 0;
}

int omcgpu_tally_values_device(double** values, int64_t* n)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  if (G.tally_replicas > 1 && G.n_tally > 0) { // hand out one array: fold the replicas first
    FoldArgs f;
    f.v = G.d_tally;
    f.n = G.n_tally;
    f.stride = G.tally_stride;
    f.replicas = G.tally_replicas;
    OMC_LAUNCH(k_fold_tally, G.n_tally, G.stream, f);
    CUDA_OK(cudaStreamSynchronize(G.stream));
  }
  *values = G.d_tally;
  *n = G.n_tally;
  return 0;
}

// The probes take MODEL material indices; with several temperatures per nuclide the device
// materials are (material, cell temperature) pairs: take the pair of that material whose
// temperature is nearest to the probe temperature ("probe_kT_neV")
static std::vector<int32_t> device_materials(const int32_t* material, int64_t n)
{
  std::vector<int32_t> out(material, material + n);
  if (G.mat_base.empty())
    return out;
  for (int64_t k = 0; k < n; ++k) {
    int best = -1;
    double max_diff = INFINITY;
    for (size_t d = 0; d < G.mat_base.size(); ++d) {
      if (G.mat_base[d] != material[k])
        continue;
      const double diff = std::fabs(G.mat_sqrtkT[d] * G.mat_sqrtkT[d] - G.probe_kT);
      if (diff < max_diff) {
        best = (int)d;
        max_diff = diff;
      }
    }
    out[k] = best;
  }
  return out;
}

int omcgpu_calculate_xs(const double* E, const int32_t* material, int64_t n, double* out,
  int32_t* i_grid, int32_t max_nuc)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  const std::vector<int32_t> dmat = device_materials(material, n);
  for (int64_t k = 0; k < n; ++k)
    if (dmat[k] < 0 || dmat[k] >= G.hm.n_materials)
      return fail(OMCGPU_E_OUT_OF_BOUNDS, "omcgpu_calculate_xs: material index out of range%s");
  material = dmat.data();
  double *dE = nullptr, *dout = nullptr;
  int32_t *dm = nullptr, *dig = nullptr;
  CUDA_OK(cudaMalloc((void**)&dE, n * sizeof(double)));
  CUDA_OK(cudaMalloc((void**)&dm, n * sizeof(int32_t)));
  CUDA_OK(cudaMalloc((void**)&dout, n * 4 * sizeof(double)));
  if (i_grid) { // entries beyond the material's nuclide count stay 0
    CUDA_OK(cudaMalloc((void**)&dig, n * max_nuc * sizeof(int32_t)));
    CUDA_OK(cudaMemset(dig, 0, n * max_nuc * sizeof(int32_t)));
  }
  CUDA_OK(cudaMemcpy(dE, E, n * sizeof(double), cudaMemcpyHostToDevice));
  CUDA_OK(cudaMemcpy(dm, material, n * sizeof(int32_t), cudaMemcpyHostToDevice));
  XsArgs a;
  a.E = dE; a.material = dm; a.n = n; a.out = dout; a.i_grid = dig; a.max_nuc = max_nuc;
  LAUNCH_XS(k_probe_xs, n, G.stream, a);
  CUDA_OK(cudaStreamSynchronize(G.stream));
  CUDA_OK(cudaMemcpy(out, dout, n * 4 * sizeof(double), cudaMemcpyDeviceToHost));
  if (i_grid)
    CUDA_OK(cudaMemcpy(i_grid, dig, n * max_nuc * sizeof(int32_t), cudaMemcpyDeviceToHost));
  cudaFree(dE); cudaFree(dm); cudaFree(dout);
  if (dig)
    cudaFree(dig);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(OMCGPU_E_CUDA, "kernel failure: %s", cudaGetErrorString(e));
  return 0;
}

int omcgpu_calculate_xs_device(const double* E, const int32_t* material, int64_t n, double* out,
  float* ms)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  XsArgs a;
  a.E = E; a.material = material; a.n = n; a.out = out; a.i_grid = nullptr; a.max_nuc = 0;
  CUDA_OK(cudaEventRecord(G.evp0, G.stream));
  LAUNCH_XS(k_probe_xs, n, G.stream, a);
  CUDA_OK(cudaEventRecord(G.evp1, G.stream));
  CUDA_OK(cudaStreamSynchronize(G.stream));
  if (ms)
    cudaEventElapsedTime(ms, G.evp0, G.evp1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(OMCGPU_E_CUDA, "kernel failure: %s", cudaGetErrorString(e));
  return 0;
}

int omcgpu_collide(int64_t n, const int32_t* material, const double* E, const double* uvw,
  const uint64_t* seed, double keff, double wgt, double* E_out, double* uvw_out, double* wgt_out,
  double* k_abs_out, uint64_t* seed_out, int32_t* n_sites, omcgpu_site* sites, int64_t site_capacity,
  int64_t* n_sites_total)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  if (n <= 0 || n > G.cap)
    return fail(OMCGPU_E_OUT_OF_BOUNDS, "omcgpu_collide: n exceeds max_particles%s");
  std::vector<void*> tmp;
  int rc = 0;
  auto dalloc = [&](size_t bytes, const void* from) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      rc = fail(OMCGPU_E_ALLOCATE, "omcgpu_collide: cudaMalloc failed%s");
      return nullptr;
    }
    tmp.push_back(p);
    if (from && cudaMemcpy(p, from, bytes, cudaMemcpyHostToDevice) != cudaSuccess)
      rc = fail(OMCGPU_E_CUDA, "omcgpu_collide: H2D copy failed%s");
    return p;
  };
  CollideArgs a;
  std::memset(&a, 0, sizeof(a));
  omcgpu_gen_in in;
  std::memset(&in, 0, sizeof(in));
  in.n_particles = n;
  in.keff = keff;
  const EvArgs ev = make_args(&in);
  a.ps = G.ps;
  a.gc = ev.gc;
  a.gc.sec_site = nullptr; // the probe reports the colliding neutron only
  a.gc.an_d = nullptr;
  a.fb = G.fis;
  a.n = n;
  a.wgt = wgt;
  const std::vector<int32_t> dmat = device_materials(material, n);
  a.material = (const int32_t*)dalloc(n * 4, dmat.data());
  a.E = (const double*)dalloc(n * 8, E);
  a.uvw = (const double*)dalloc(n * 24, uvw);
  a.seed = (const uint64_t*)dalloc(n * 8, seed);
  a.E_out = (double*)dalloc(n * 8, nullptr);
  a.uvw_out = (double*)dalloc(n * 24, nullptr);
  a.wgt_out = (double*)dalloc(n * 8, nullptr);
  a.k_abs_out = (double*)dalloc(n * 8, nullptr);
  a.seed_out = (uint64_t*)dalloc(n * 8, nullptr);
  auto done = [&](int code) {
    for (void* p : tmp)
      cudaFree(p);
    return code;
  };
  if (rc)
    return done(rc);
  cudaMemsetAsync(G.d_counts, 0, N_COUNTS * sizeof(int32_t), G.stream);
  OMC_LAUNCH(k_probe_collide, n, G.stream, a);
  if (read_counts())
    return done(OMCGPU_E_CUDA);
  if (G.h_counts[5] > 0)
    return done(fail(OMCGPU_E_PHYSICS, "omcgpu_collide: a neutron found no nuclide to collide with%s"));
  if (G.h_counts[4] > G.fis_cap || G.h_counts[4] > site_capacity)
    return done(fail(OMCGPU_E_OUT_OF_BOUNDS, "omcgpu_collide: more fission sites than the bank / caller's buffer holds%s"));
  G.n_fis = G.h_counts[4];
  if (sort_fission_bank(n, 0))
    return done(OMCGPU_E_CUDA);
  cudaMemcpyAsync(E_out, a.E_out, n * 8, cudaMemcpyDeviceToHost, G.stream);
  cudaMemcpyAsync(uvw_out, a.uvw_out, n * 24, cudaMemcpyDeviceToHost, G.stream);
  cudaMemcpyAsync(wgt_out, a.wgt_out, n * 8, cudaMemcpyDeviceToHost, G.stream);
  cudaMemcpyAsync(k_abs_out, a.k_abs_out, n * 8, cudaMemcpyDeviceToHost, G.stream);
  cudaMemcpyAsync(seed_out, a.seed_out, n * 8, cudaMemcpyDeviceToHost, G.stream);
  cudaMemcpyAsync(n_sites, G.ps.n_progeny, n * 4, cudaMemcpyDeviceToHost, G.stream);
  if (n_sites_total)
    *n_sites_total = G.n_fis;
  rc = G.n_fis > 0 ? fetch_bank(G.fis_sorted, sites, G.n_fis) : 0;
  cudaStreamSynchronize(G.stream);
  cudaError_t e = cudaGetLastError();
  if (!rc && e != cudaSuccess)
    rc = fail(OMCGPU_E_CUDA, "kernel failure: %s", cudaGetErrorString(e));
  return done(rc);
}

static int geom_probe(const double* xyz, const double* uvw, int64_t n, int32_t* cell,
  int32_t* instance, int32_t* material, int32_t* n_coord, double* distance, int32_t* surface,
  int32_t* level, int32_t* lattice_trans)
{
  if (!G.ready)
    return fail(OMCGPU_E_UNASSIGNED, "omcgpu_init has not been called%s");
  GeoArgs a;
  std::memset(&a, 0, sizeof(a));
  a.ps = G.ps;
  a.n = n;
  std::vector<void*> tmp;
  bool alloc_failed = false;
  auto dalloc = [&](size_t bytes) {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess)
      alloc_failed = true;
    else
      tmp.push_back(p);
    return p;
  };
  auto release = [&]() {
    for (void* p : tmp)
      cudaFree(p);
    tmp.clear();
  };
  for (int64_t done = 0; done < n; done += G.cap) {
    const int64_t m = std::min<int64_t>(G.cap, n - done);
    double* dx = (double*)dalloc(m * 3 * sizeof(double));
    cudaMemcpy(dx, xyz + 3 * done, m * 3 * sizeof(double), cudaMemcpyHostToDevice);
    a.xyz = dx;
    a.uvw = nullptr;
    if (uvw) {
      double* du = (double*)dalloc(m * 3 * sizeof(double));
      cudaMemcpy(du, uvw + 3 * done, m * 3 * sizeof(double), cudaMemcpyHostToDevice);
      a.uvw = du;
    }
    a.n = m;
    if (cell) {
      a.cell = (int32_t*)dalloc(m * 4);
      a.instance = (int32_t*)dalloc(m * 4);
      a.material = (int32_t*)dalloc(m * 4);
      a.n_coord = (int32_t*)dalloc(m * 4);
    }
    if (distance) {
      a.distance = (double*)dalloc(m * 8);
      a.surface = (int32_t*)dalloc(m * 4);
      a.level = (int32_t*)dalloc(m * 4);
      a.lat_trans = (int32_t*)dalloc(m * 12);
    }
    if (alloc_failed) {
      release();
      cudaGetLastError();
      return fail(OMCGPU_E_ALLOCATE, "geometry probe: cudaMalloc failed%s");
    }
    OMC_LAUNCH(k_probe_geom, m, G.stream, a);
    cudaStreamSynchronize(G.stream);
    if (cell) {
      cudaMemcpy(cell + done, a.cell, m * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(instance + done, a.instance, m * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(material + done, a.material, m * 4, cudaMemcpyDeviceToHost);
      if (!G.mat_base.empty()) // device materials -> the model's materials
        for (int64_t k = done; k < done + m; ++k)
          if (material[k] >= 0)
            material[k] = G.mat_base[material[k]];
      cudaMemcpy(n_coord + done, a.n_coord, m * 4, cudaMemcpyDeviceToHost);
    }
    if (distance) {
      cudaMemcpy(distance + done, a.distance, m * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(surface + done, a.surface, m * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(level + done, a.level, m * 4, cudaMemcpyDeviceToHost);
      cudaMemcpy(lattice_trans + 3 * done, a.lat_trans, m * 12, cudaMemcpyDeviceToHost);
    }
    release();
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return fail(OMCGPU_E_CUDA, "kernel failure: %s", cudaGetErrorString(e));
  return 0;
}

int omcgpu_find_cell(const double* xyz, const double* uvw, int64_t n, int32_t* cell,
  int32_t* instance, int32_t* material, int32_t* n_coord)
{
  return geom_probe(xyz, uvw, n, cell, instance, material, n_coord, nullptr, nullptr, nullptr, nullptr);
}

int omcgpu_distance_to_boundary(const double* xyz, const double* uvw, int64_t n, double* distance,
  int32_t* surface, int32_t* level, int32_t* lattice_trans)
{
  return geom_probe(xyz, uvw, n, nullptr, nullptr, nullptr, nullptr, distance, surface, level,
    lattice_trans);
}

static int rng_probe(const uint64_t* n, const uint64_t* seed, int64_t count, int32_t per,
  uint64_t* out_u, double* out_d)
{
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
    return fail(OMCGPU_E_CUDA, "no CUDA device available: the device path has no CPU fallback%s");
  RngArgs a;
  std::memset(&a, 0, sizeof(a));
  uint64_t *dn = nullptr, *ds = nullptr, *du = nullptr;
  double* dd = nullptr;
  CUDA_OK(cudaMalloc((void**)&ds, count * 8));
  CUDA_OK(cudaMemcpy(ds, seed, count * 8, cudaMemcpyHostToDevice));
  if (n) {
    CUDA_OK(cudaMalloc((void**)&dn, count * 8));
    CUDA_OK(cudaMemcpy(dn, n, count * 8, cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc((void**)&du, count * 8));
  } else {
    CUDA_OK(cudaMalloc((void**)&dd, count * per * 8));
  }
  a.n = dn; a.seed = ds; a.count = count; a.per = per; a.out_u = du; a.out_d = dd;
  OMC_LAUNCH(k_probe_rng, count, 0, a);
  CUDA_OK(cudaDeviceSynchronize());
  if (out_u)
    CUDA_OK(cudaMemcpy(out_u, du, count * 8, cudaMemcpyDeviceToHost));
  else
    CUDA_OK(cudaMemcpy(out_d, dd, count * per * 8, cudaMemcpyDeviceToHost));
  cudaFree(ds);
  if (dn) cudaFree(dn);
  if (du) cudaFree(du);
  if (dd) cudaFree(dd);
  return 0;
}

int omcgpu_future_seed(const uint64_t* n, const uint64_t* seed, int64_t count, uint64_t* out)
{
  return rng_probe(n, seed, count, 0, out, nullptr);
}

int omcgpu_prn(const uint64_t* seed, int64_t n_seeds, int32_t count, double* out)
{
  return rng_probe(nullptr, seed, n_seeds, count, nullptr, out);
}

int omcgpu_comm_allgather_i64(const int64_t* send, int32_t n, int64_t* recv);

} // extern "C"

#include "comm.h"
