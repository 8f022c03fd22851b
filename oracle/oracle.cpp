// =============================================================================
// oracle.cpp -- CPU RESTATEMENT of OpenMC's transport path.  TEST INFRASTRUCTURE.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load this library.  The product (libomcgpu.so) never does.
//
// This is a plain, scalar, history-based restatement (one Particle struct at a
// time, like transport_history_based, src/simulation.cpp:804-831) of the
// algorithms the device engine implements event-based.  Every function cites
// the reference file:line it follows.  It consumes the same flattened POD model
// (include/omcgpu.h) as the device engine, so both see identical inputs.
//
// Pinning (tests/test_oracle_golden.py): the eigenvalue driver below reproduces
// the reference's own regression goldens for the multigroup path digit for
// digit (tests/regression_tests/mg_legendre, mg_max_order, mg_survival_biasing,
// mg_temperature_multi, mg_basic results_true.dat), and the helpers are checked
// against the reference's own translation units compiled verbatim into
// oracle/_ref (random_lcg.cpp, math_functions.cpp, random_dist.cpp).
// =============================================================================
#include "../include/omcgpu.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <string>
#include <vector>

namespace {

// ---- constants: include/openmc/constants.h:49-54,61,81-83,102,112,337,377 ----
constexpr double TINY_BIT = 1e-8;
constexpr double FP_PRECISION = 1e-14;
constexpr double FP_REL_PRECISION = 1e-5;
constexpr double FP_COINCIDENT = 1e-12;
constexpr int MAX_SAMPLE = 100000;
constexpr double PI = 3.141592653589793238462643383279502884L;
constexpr double INFTY = std::numeric_limits<double>::max();
constexpr double MASS_NEUTRON_EV = 939.56542052e6;
constexpr double C_LIGHT = 2.99792458e10;
constexpr int C_NONE = -1;
constexpr int MATERIAL_VOID = -1;
constexpr int SURFACE_NONE = 0;
constexpr int STREAM_TRACKING = 0;
constexpr int STREAM_SOURCE = 1;

const omcgpu_model* M = nullptr;
char errbuf[512] = "";

struct Vec3 {
  double x, y, z;
  double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
  double& ref(int i) { return i == 0 ? x : (i == 1 ? y : z); }
};
inline Vec3 operator+(Vec3 a, Vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, Vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 operator*(Vec3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline Vec3 operator/(Vec3 a, double s) { return {a.x / s, a.y / s, a.z / s}; }
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm(Vec3 a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

// =============================================================================
// RNG -- src/random_lcg.cpp:11-13,32-44,70-76,91-120
// =============================================================================
constexpr uint64_t PRN_MULT = 6364136223846793005ULL;
constexpr uint64_t PRN_ADD = 1442695040888963407ULL;

double prn(uint64_t* seed)
{
  *seed = PRN_MULT * (*seed) + PRN_ADD;
  uint64_t word = ((*seed >> ((*seed >> 59u) + 5u)) ^ *seed) * 12605985483714917081ull;
  uint64_t result = (word >> 43u) ^ word;
  return std::ldexp(static_cast<double>(result), -64);
}

uint64_t future_seed(uint64_t n, uint64_t seed)
{
  uint64_t g = PRN_MULT, c = PRN_ADD, g_new = 1, c_new = 0;
  while (n > 0) {
    if (n & 1) {
      g_new *= g;
      c_new = c_new * g + c;
    }
    c *= (g + 1);
    g *= g;
    n >>= 1;
  }
  return g_new * seed + c_new;
}

uint64_t init_seed(int64_t id, int offset, int64_t master_seed, uint64_t stride)
{
  return future_seed(static_cast<uint64_t>(id) * stride, master_seed + offset);
}

// include/openmc/search.h:14-20
template<class T>
int64_t lower_bound_index(const T* first, const T* last, T value)
{
  if (*first == value)
    return 0;
  return std::lower_bound(first, last, value) - first - 1;
}

// src/random_dist.cpp:20-39
double maxwell_spectrum(double T, uint64_t* seed)
{
  double r1 = prn(seed);
  double r2 = prn(seed);
  double r3 = prn(seed);
  double c = std::cos(PI / 2. * r3);
  return -T * (std::log(r1) + std::log(r2) * c * c);
}
double uniform_distribution(double a, double b, uint64_t* seed)
{
  return a + (b - a) * prn(seed);
}
double watt_spectrum(double a, double b, uint64_t* seed)
{
  double w = maxwell_spectrum(a, seed);
  return w + 0.25 * a * a * b + uniform_distribution(-1., 1., seed) * std::sqrt(a * a * b * w);
}

// src/math_functions.cpp:98-121
double evaluate_legendre(int n, const double* data, double x)
{
  double pnx[32];
  pnx[0] = 1.;
  if (n >= 1)
    pnx[1] = x;
  for (int l = 1; l < n; l++)
    pnx[l + 1] = ((2 * l + 1) * x * pnx[l] - l * pnx[l - 1]) / (l + 1);
  double val = 0.0;
  for (int l = 0; l <= n; l++)
    val += (l + 0.5) * data[l] * pnx[l];
  return val;
}

// src/math_functions.cpp:763-792
Vec3 rotate_angle(Vec3 u, double mu, const double* phi, uint64_t* seed)
{
  double phi_ = phi ? *phi : 2.0 * PI * prn(seed);
  double sinphi = std::sin(phi_);
  double cosphi = std::cos(phi_);
  double a = std::sqrt(std::fmax(0., 1. - mu * mu));
  double b = std::sqrt(std::fmax(0., 1. - u.z * u.z));
  if (b > 1e-10) {
    return {mu * u.x + a * (u.x * u.z * cosphi - u.y * sinphi) / b,
      mu * u.y + a * (u.y * u.z * cosphi + u.x * sinphi) / b, mu * u.z - a * b * cosphi};
  } else {
    b = std::sqrt(1. - u.y * u.y);
    return {mu * u.x + a * (-u.x * u.y * sinphi + u.z * cosphi) / b, mu * u.y + a * b * sinphi,
      mu * u.z - a * (u.y * u.z * sinphi + u.x * cosphi) / b};
  }
}

// src/distribution_multi.cpp:106-112
Vec3 isotropic_direction(uint64_t* seed)
{
  double phi = uniform_distribution(0., 2.0 * PI, seed);
  double mu = uniform_distribution(-1., 1., seed);
  return {mu, std::sqrt(1.0 - mu * mu) * std::cos(phi), std::sqrt(1.0 - mu * mu) * std::sin(phi)};
}

// =============================================================================
// Particle state -- include/openmc/particle_data.h:103-141,203-215,418-566
// =============================================================================
struct LocalCoord {
  Vec3 r {0, 0, 0};
  Vec3 u {0, 0, 0};
  int cell {C_NONE};
  int universe {C_NONE};
  int lattice {C_NONE};
  int lat_idx[3] {0, 0, 0};
  void reset()
  {
    cell = C_NONE;
    universe = C_NONE;
    lattice = C_NONE;
    lat_idx[0] = lat_idx[1] = lat_idx[2] = 0;
  }
};

struct BoundaryInfo {
  double distance {INFINITY};
  int surface {SURFACE_NONE};
  int coord_level {0};
  int lat_trans[3] {0, 0, 0};
};

struct MicroXS { // include/openmc/particle_data.h:148-180 (fields the path uses)
  double total, absorption, fission, nu_fission, elastic;
  int index_grid;
  double interp_factor;
};

struct Particle {
  LocalCoord coord[OMCGPU_MAX_COORD];
  int n_coord {1};
  int cell_last[OMCGPU_MAX_COORD];
  int n_coord_last {1};
  int cell_instance {0};
  int surface {SURFACE_NONE};
  int material {C_NONE}, material_last {C_NONE};
  double sqrtkT {-1.0}, sqrtkT_last {0.0};
  double density_mult {1.0}, density_mult_last {1.0}; // Cell::density_mult of the cell instance, src/geometry.cpp:180-181
  double E {0}, E_last {0};
  int g {0}, g_last {0};
  double wgt {1}, wgt_last {1}, wgt_born {1};
  double mu {0};
  double time {0};
  Vec3 r_last {0, 0, 0}, u_last {0, 0, 0};
  Vec3 u_last_local {0, 0, 0}; // direction at the lowest level before the collision (MG angle index of a collision-estimator score)
  uint64_t seeds[4];
  int stream {STREAM_TRACKING};
  int64_t id {0};
  int n_progeny {0};
  int64_t n_event {0};
  int n_collision {0};
  std::vector<omcgpu_site> secondary_bank; // Particle::create_secondary, src/particle.cpp:75-93
  std::vector<omcgpu_site> nu_bank;        // the fission sites banked in THIS collision (p.nu_bank(), score_fission_eout)
  bool fission {false};
  // what the analog estimator reads after a collision (src/tallies/tally_scoring.cpp:1116-1330, 1622-1990)
  bool event_scatter {false};  // p.event() == TallyEvent::SCATTER
  double wgt_bank {0};         // fission sites stored / weight (src/physics.cpp:280-281)
  MicroXS event_micro {};      // p.neutron_xs(p.event_nuclide()) at the collision energy
  int delayed_group {0};
  int cell_born {C_NONE};
  double xs_total {0}, xs_absorption {0}, xs_fission {0}, xs_nu_fission {0};
  int mg_material {C_NONE}, mg_t {0}, mg_a {0};
  double mg_sqrtkT {0};
  Vec3 mg_u {0, 0, 0};
  BoundaryInfo boundary;
  double collision_distance {0};
  double k_abs {0}, k_col {0}, k_trk {0}, k_leak {0};
  bool lost {false};

  uint64_t* current_seed() { return &seeds[stream]; }
  Vec3& r() { return coord[0].r; }
  Vec3& u() { return coord[0].u; }
  Vec3& r_local() { return coord[n_coord - 1].r; }
  Vec3& u_local() { return coord[n_coord - 1].u; }
  LocalCoord& lowest_coord() { return coord[n_coord - 1]; }
  bool alive() const { return wgt != 0.0; }
};

// run-time state the physics needs (simulation:: globals)
struct RunState {
  double keff {1.0};
  int64_t work_index {0};
  bool tallies_active {false};
  int n_collision_tallies {0};
  int n_analog_tallies {0};
  std::vector<omcgpu_site> fission_bank;
  int64_t fission_capacity {0};
  std::vector<int64_t> progeny_per_particle;
  std::vector<double> tally_values;
  double global_tally[4] {0, 0, 0, 0};
  double total_weight {0};
  int64_t n_lost {0};
  int64_t n_events {0}, n_xs {0}, n_coll {0}, n_cross {0};
} RS;

void mark_as_lost(Particle& p, const char* msg)
{
  // src/particle.cpp:762-785 (no restart file; counts only)
  (void)msg;
  p.wgt = 0.0;
  p.lost = true;
  RS.n_lost += 1;
}

// =============================================================================
// Surfaces -- src/surface.cpp:116-130 (sense), 207-220 (planes), 354-375,
// 399-456 (cylinders), 598-648 (sphere)
// =============================================================================
double surf_evaluate(int s, Vec3 r)
{
  const double* c = M->surf_coef + OMCGPU_SURF_NCOEF * s;
  switch (M->surf_type[s]) {
  case OMCGPU_SURF_XPLANE:
    return r.x - c[0];
  case OMCGPU_SURF_YPLANE:
    return r.y - c[0];
  case OMCGPU_SURF_ZPLANE:
    return r.z - c[0];
  case OMCGPU_SURF_PLANE:
    return c[0] * r.x + c[1] * r.y + c[2] * r.z - c[3];
  case OMCGPU_SURF_XCYLINDER: {
    const double r1 = r.y - c[0], r2 = r.z - c[1];
    return r1 * r1 + r2 * r2 - c[2] * c[2];
  }
  case OMCGPU_SURF_YCYLINDER: {
    const double r1 = r.x - c[0], r2 = r.z - c[1];
    return r1 * r1 + r2 * r2 - c[2] * c[2];
  }
  case OMCGPU_SURF_ZCYLINDER: {
    const double r1 = r.x - c[0], r2 = r.y - c[1];
    return r1 * r1 + r2 * r2 - c[2] * c[2];
  }
  case OMCGPU_SURF_SPHERE: {
    const double x = r.x - c[0], y = r.y - c[1], z = r.z - c[2];
    return x * x + y * y + z * z - c[3] * c[3];
  }
  // axis_aligned_cone_evaluate, src/surface.cpp:688-696 (r1 along the cone's axis)
  case OMCGPU_SURF_XCONE: {
    const double r1 = r.x - c[0], r2 = r.y - c[1], r3 = r.z - c[2];
    return r2 * r2 + r3 * r3 - c[3] * r1 * r1;
  }
  case OMCGPU_SURF_YCONE: {
    const double r1 = r.y - c[1], r2 = r.x - c[0], r3 = r.z - c[2];
    return r2 * r2 + r3 * r3 - c[3] * r1 * r1;
  }
  case OMCGPU_SURF_ZCONE: {
    const double r1 = r.z - c[2], r2 = r.x - c[0], r3 = r.y - c[1];
    return r2 * r2 + r3 * r3 - c[3] * r1 * r1;
  }
  case OMCGPU_SURF_QUADRIC: { // SurfaceQuadric::evaluate, src/surface.cpp:861-868
    c = M->quad_coef + 10 * (int)c[0];
    const double x = r.x, y = r.y, z = r.z;
    return x * (c[0] * x + c[3] * y + c[6]) + y * (c[1] * y + c[4] * z + c[7]) + z * (c[2] * z + c[5] * x + c[8]) + c[9];
  }
  }
  return 0.0;
}

Vec3 surf_normal(int s, Vec3 r)
{
  const double* c = M->surf_coef + OMCGPU_SURF_NCOEF * s;
  switch (M->surf_type[s]) {
  case OMCGPU_SURF_XPLANE:
    return {1., 0., 0.};
  case OMCGPU_SURF_YPLANE:
    return {0., 1., 0.};
  case OMCGPU_SURF_ZPLANE:
    return {0., 0., 1.};
  case OMCGPU_SURF_PLANE:
    return {c[0], c[1], c[2]};
  case OMCGPU_SURF_XCYLINDER:
    return {0.0, 2.0 * (r.y - c[0]), 2.0 * (r.z - c[1])};
  case OMCGPU_SURF_YCYLINDER:
    return {2.0 * (r.x - c[0]), 0.0, 2.0 * (r.z - c[1])};
  case OMCGPU_SURF_ZCYLINDER:
    return {2.0 * (r.x - c[0]), 2.0 * (r.y - c[1]), 0.0};
  case OMCGPU_SURF_SPHERE:
    return {2.0 * (r.x - c[0]), 2.0 * (r.y - c[1]), 2.0 * (r.z - c[2])};
  // axis_aligned_cone_normal, src/surface.cpp:747-756
  case OMCGPU_SURF_XCONE:
    return {-2.0 * c[3] * (r.x - c[0]), 2.0 * (r.y - c[1]), 2.0 * (r.z - c[2])};
  case OMCGPU_SURF_YCONE:
    return {2.0 * (r.x - c[0]), -2.0 * c[3] * (r.y - c[1]), 2.0 * (r.z - c[2])};
  case OMCGPU_SURF_ZCONE:
    return {2.0 * (r.x - c[0]), 2.0 * (r.y - c[1]), -2.0 * c[3] * (r.z - c[2])};
  case OMCGPU_SURF_QUADRIC: // SurfaceQuadric::normal, src/surface.cpp:941-949
    c = M->quad_coef + 10 * (int)c[0];
    return {2.0 * c[0] * r.x + c[3] * r.y + c[5] * r.z + c[6], 2.0 * c[1] * r.y + c[3] * r.x + c[4] * r.z + c[7],
      2.0 * c[2] * r.z + c[4] * r.y + c[5] * r.x + c[8]};
  }
  return {0, 0, 0};
}

bool surf_sense(int s, Vec3 r, Vec3 u)
{
  const double f = surf_evaluate(s, r);
  if (std::abs(f) < FP_COINCIDENT)
    return dot(u, surf_normal(s, r)) > 0.0;
  return f > 0.0;
}

double plane_axis_distance(double ri, double ui, bool coincident, double offset)
{
  const double f = offset - ri;
  if (coincident || std::abs(f) < FP_COINCIDENT || ui == 0.0)
    return INFTY;
  const double d = f / ui;
  if (d < 0.0)
    return INFTY;
  return d;
}

// axis = cylinder axis component; (r2,u2),(r3,u3) are the perpendicular ones
double cylinder_axis_distance(double u1, double r2, double r3, double u2, double u3,
  bool coincident, double off1, double off2, double radius)
{
  const double a = 1.0 - u1 * u1;
  if (a == 0.0)
    return INFTY;
  const double q2 = r2 - off1;
  const double q3 = r3 - off2;
  const double k = q2 * u2 + q3 * u3;
  const double c = q2 * q2 + q3 * q3 - radius * radius;
  const double quad = k * k - a * c;
  if (quad < 0.0) {
    return INFTY;
  } else if (coincident || std::abs(c) < FP_COINCIDENT) {
    if (k >= 0.0)
      return INFTY;
    return (-k + std::sqrt(quad)) / a;
  } else if (c < 0.0) {
    return (-k + std::sqrt(quad)) / a;
  } else {
    const double d = (-k - std::sqrt(quad)) / a;
    if (d < 0.0)
      return INFTY;
    return d;
  }
}

// axis_aligned_cone_distance, src/surface.cpp:701-742 (index 1 = the cone's axis)
double cone_axis_distance(double r1, double r2, double r3, double u1, double u2, double u3, double radius_sq,
  bool coincident)
{
  const double a = u2 * u2 + u3 * u3 - radius_sq * u1 * u1;
  const double k = r2 * u2 + r3 * u3 - radius_sq * r1 * u1;
  const double c = r2 * r2 + r3 * r3 - radius_sq * r1 * r1;
  double quad = k * k - a * c;
  double d;
  if (quad < 0.0) {
    return INFTY;
  } else if (coincident || std::abs(c) < FP_COINCIDENT) {
    if (k >= 0.0) {
      d = (-k - std::sqrt(quad)) / a;
    } else {
      d = (-k + std::sqrt(quad)) / a;
    }
  } else {
    quad = std::sqrt(quad);
    d = (-k - quad) / a;
    const double b = (-k + quad) / a;
    if (d < 0.0) {
      if (b > 0.0)
        d = b;
    } else {
      if (b > 0.0) {
        if (b < d)
          d = b;
      }
    }
  }
  if (d <= 0.0)
    return INFTY;
  return d;
}

double surf_distance(int s, Vec3 r, Vec3 u, bool coincident)
{
  const double* c = M->surf_coef + OMCGPU_SURF_NCOEF * s;
  switch (M->surf_type[s]) {
  case OMCGPU_SURF_XPLANE:
    return plane_axis_distance(r.x, u.x, coincident, c[0]);
  case OMCGPU_SURF_YPLANE:
    return plane_axis_distance(r.y, u.y, coincident, c[0]);
  case OMCGPU_SURF_ZPLANE:
    return plane_axis_distance(r.z, u.z, coincident, c[0]);
  case OMCGPU_SURF_PLANE: {
    const double f = c[0] * r.x + c[1] * r.y + c[2] * r.z - c[3];
    const double projection = c[0] * u.x + c[1] * u.y + c[2] * u.z;
    if (coincident || std::abs(f) < FP_COINCIDENT || projection == 0.0)
      return INFTY;
    const double d = -f / projection;
    if (d < 0.0)
      return INFTY;
    return d;
  }
  case OMCGPU_SURF_XCYLINDER:
    return cylinder_axis_distance(u.x, r.y, r.z, u.y, u.z, coincident, c[0], c[1], c[2]);
  case OMCGPU_SURF_YCYLINDER:
    return cylinder_axis_distance(u.y, r.x, r.z, u.x, u.z, coincident, c[0], c[1], c[2]);
  case OMCGPU_SURF_ZCYLINDER:
    return cylinder_axis_distance(u.z, r.x, r.y, u.x, u.y, coincident, c[0], c[1], c[2]);
  case OMCGPU_SURF_SPHERE: {
    const double x = r.x - c[0], y = r.y - c[1], z = r.z - c[2];
    const double k = x * u.x + y * u.y + z * u.z;
    const double cc = x * x + y * y + z * z - c[3] * c[3];
    const double quad = k * k - cc;
    if (quad < 0.0) {
      return INFTY;
    } else if (coincident || std::abs(cc) < FP_COINCIDENT) {
      if (k >= 0.0)
        return INFTY;
      return -k + std::sqrt(quad);
    } else if (cc < 0.0) {
      return -k + std::sqrt(quad);
    } else {
      const double d = -k - std::sqrt(quad);
      if (d < 0.0)
        return INFTY;
      return d;
    }
  }
  case OMCGPU_SURF_XCONE:
    return cone_axis_distance(r.x - c[0], r.y - c[1], r.z - c[2], u.x, u.y, u.z, c[3], coincident);
  case OMCGPU_SURF_YCONE:
    return cone_axis_distance(r.y - c[1], r.x - c[0], r.z - c[2], u.y, u.x, u.z, c[3], coincident);
  case OMCGPU_SURF_ZCONE:
    return cone_axis_distance(r.z - c[2], r.x - c[0], r.y - c[1], u.z, u.x, u.y, c[3], coincident);
  case OMCGPU_SURF_QUADRIC: { // SurfaceQuadric::distance, src/surface.cpp:870-939
    c = M->quad_coef + 10 * (int)c[0];
    const double A = c[0], B = c[1], C = c[2], D = c[3], E = c[4], F = c[5], G = c[6], H = c[7], J = c[8], K = c[9];
    const double x = r.x, y = r.y, z = r.z, uu = u.x, v = u.y, w = u.z;
    const double a = A * uu * uu + B * v * v + C * w * w + D * uu * v + E * v * w + F * uu * w;
    const double k = A * uu * x + B * v * y + C * w * z +
                     0.5 * (D * (uu * y + v * x) + E * (v * z + w * y) + F * (w * x + uu * z) + G * uu + H * v + J * w);
    const double cc = A * x * x + B * y * y + C * z * z + D * x * y + E * y * z + F * x * z + G * x + H * y + J * z + K;
    double quad = k * k - a * cc;
    double d;
    if (quad < 0.0) {
      return INFTY;
    } else if (coincident || std::abs(cc) < FP_COINCIDENT) {
      if (a == 0.0)
        d = INFTY;
      else if (k >= 0.0)
        d = (-k - std::sqrt(quad)) / a;
      else
        d = (-k + std::sqrt(quad)) / a;
    } else if (a == 0.0) {
      d = -0.5 * cc / k;
    } else {
      quad = std::sqrt(quad);
      d = (-k - quad) / a;
      const double b = (-k + quad) / a;
      if (d < 0.0) {
        if (b > 0.0)
          d = b;
      } else if (b > 0.0 && b < d) {
        d = b;
      }
    }
    return d <= 0.0 ? INFTY : d;
  }
  }
  return INFTY;
}

// =============================================================================
// Regions -- src/cell.cpp:910-936 (distance), 940-970 (simple), 974-1029 (complex)
// =============================================================================
bool cell_contains(int c, Vec3 r, Vec3 u, int on_surface)
{
  const int32_t* tok = M->region_tokens + M->cell_region_ofs[c];
  const int n = M->cell_region_ofs[c + 1] - M->cell_region_ofs[c];
  if (M->cell_simple[c]) {
    for (int i = 0; i < n; ++i) {
      int32_t token = tok[i];
      if (token == on_surface) {
      } else if (-token == on_surface) {
        return false;
      } else {
        bool sense = surf_sense(std::abs(token) - 1, r, u);
        if (sense != (token > 0))
          return false;
      }
    }
    return true;
  }
  bool in_cell = true;
  int total_depth = 0;
  for (int it = 0; it < n; ++it) {
    int32_t token = tok[it];
    if (token < OMCGPU_OP_UNION) {
      if (token == on_surface) {
        in_cell = true;
      } else if (-token == on_surface) {
        in_cell = false;
      } else {
        bool sense = surf_sense(std::abs(token) - 1, r, u);
        in_cell = (sense == (token > 0));
      }
    } else if ((token == OMCGPU_OP_UNION && in_cell) ||
               (token == OMCGPU_OP_INTERSECTION && !in_cell)) {
      if (total_depth == 0)
        return in_cell;
      total_depth--;
      int depth = 1;
      do {
        it++;
        int32_t next_token = tok[it];
        if (next_token > OMCGPU_OP_COMPLEMENT) {
          if (next_token == OMCGPU_OP_RIGHT_PAREN)
            depth--;
          else
            depth++;
        }
      } while (depth > 0);
    } else if (token == OMCGPU_OP_LEFT_PAREN) {
      total_depth++;
    } else if (token == OMCGPU_OP_RIGHT_PAREN) {
      total_depth--;
    }
  }
  return in_cell;
}

void cell_distance(int c, Vec3 r, Vec3 u, int on_surface, double& min_dist, int& i_surf)
{
  min_dist = INFTY;
  i_surf = std::numeric_limits<int32_t>::max();
  const int32_t* tok = M->region_tokens + M->cell_region_ofs[c];
  const int n = M->cell_region_ofs[c + 1] - M->cell_region_ofs[c];
  for (int i = 0; i < n; ++i) {
    int32_t token = tok[i];
    if (token >= OMCGPU_OP_UNION)
      continue;
    bool coincident = std::abs(token) == std::abs(on_surface);
    double d = surf_distance(std::abs(token) - 1, r, u, coincident);
    if (d < min_dist) {
      if (min_dist - d >= FP_PRECISION * min_dist) {
        min_dist = d;
        i_surf = -token;
      }
    }
  }
}

// =============================================================================
// Rectangular lattices -- src/lattice.cpp:243-248,252-306,310-342,352-377
// =============================================================================
bool lat_is_hex(int l)
{
  return M->lat_type[l] != OMCGPU_LATTICE_RECT;
}
bool lat_valid(int l, const int* i)
{
  const int32_t* n = M->lat_shape + 3 * l;
  if (lat_is_hex(l)) { // HexLattice::are_valid_indices, src/lattice.cpp:727-735
    const int n_rings = (n[0] + 1) / 2;
    return (i[0] >= 0) && (i[1] >= 0) && (i[2] >= 0) && (i[0] < 2 * n_rings - 1) && (i[1] < 2 * n_rings - 1) &&
           (i[0] + i[1] > n_rings - 2) && (i[0] + i[1] < 3 * n_rings - 2) && (i[2] < n[2]);
  }
  return i[0] >= 0 && i[0] < n[0] && i[1] >= 0 && i[1] < n[1] && i[2] >= 0 && i[2] < n[2];
}
int lat_flat(int l, const int* i)
{
  const int32_t* n = M->lat_shape + 3 * l;
  return n[0] * n[1] * i[2] + n[0] * i[1] + i[0];
}
int lat_universe(int l, const int* i)
{
  return M->lat_universes[M->lat_univ_ofs[l] + lat_flat(l, i)];
}
int lat_offset(int l, int map, const int* i)
{
  const int32_t* n = M->lat_shape + 3 * l;
  return M->lat_offsets[M->lat_offset_ofs[l] + n[0] * n[1] * n[2] * map + lat_flat(l, i)];
}
bool coincident(double d1, double d2)
{
  return std::abs(d1 - d2) < FP_COINCIDENT;
}
Vec3 lat_local_position(int l, Vec3 r, const int* i);
// HexLattice::get_indices, src/lattice.cpp:879-973
void hex_get_indices(int l, Vec3 r, Vec3 u, int* result)
{
  const double* center = M->lat_lower_left + 3 * l;
  const double* pitch = M->lat_pitch + 3 * l;
  const int n_rings = (M->lat_shape[3 * l] + 1) / 2;
  const int n_axial = M->lat_shape[3 * l + 2];
  const bool is_3d = M->lat_is3d[l];
  Vec3 r_o {r.x - center[0], r.y - center[1], r.z};
  if (is_3d)
    r_o.z -= center[2];
  result[2] = 0;
  if (is_3d) {
    double iz_ {r_o.z / pitch[1] + 0.5 * n_axial};
    long iz_close {std::lround(iz_)};
    if (coincident(iz_, iz_close))
      result[2] = (u.z > 0) ? iz_close : iz_close - 1;
    else
      result[2] = std::floor(iz_);
  }
  if (M->lat_type[l] == OMCGPU_LATTICE_HEX_Y) {
    double alpha = r_o.y - r_o.x / std::sqrt(3.0);
    result[0] = std::floor(r_o.x / (0.5 * std::sqrt(3.0) * pitch[0]));
    result[1] = std::floor(alpha / pitch[0]);
  } else {
    double alpha = r_o.y - r_o.x * std::sqrt(3.0);
    result[0] = std::floor(-alpha / (std::sqrt(3.0) * pitch[0]));
    result[1] = std::floor(r_o.y / (0.5 * std::sqrt(3.0) * pitch[0]));
  }
  result[0] += n_rings - 1;
  result[1] += n_rings - 1;
  int i1_chg = 0;
  int i2_chg = 0;
  double d_min {INFTY};
  double dp_min {INFTY};
  for (int i = 0; i < 2; i++) {
    for (int j = 0; j < 2; j++) {
      const int i_xyz[3] = {result[0] + j, result[1] + i, 0};
      Vec3 r_t = lat_local_position(l, r, i_xyz);
      double d = r_t.x * r_t.x + r_t.y * r_t.y;
      bool on_boundary = coincident(1.0, d_min / d);
      if (d < d_min || on_boundary) {
        const double nrm = std::sqrt(d);
        r_t = Vec3 {r_t.x / nrm, r_t.y / nrm, r_t.z / nrm};
        double dp = u.x * r_t.x + u.y * r_t.y;
        if (on_boundary && dp > dp_min)
          continue;
        d_min = d;
        i1_chg = j;
        i2_chg = i;
        dp_min = dp;
      }
    }
  }
  result[0] += i1_chg;
  result[1] += i2_chg;
}
void lat_get_indices(int l, Vec3 r, Vec3 u, int* result)
{
  if (lat_is_hex(l)) {
    hex_get_indices(l, r, u, result);
    return;
  }
  const double* ll = M->lat_lower_left + 3 * l;
  const double* pitch = M->lat_pitch + 3 * l;
  double ix_ = (r.x - ll[0]) / pitch[0];
  long ix_close = std::lround(ix_);
  if (coincident(ix_, ix_close))
    result[0] = (u.x > 0) ? ix_close : ix_close - 1;
  else
    result[0] = std::floor(ix_);
  double iy_ = (r.y - ll[1]) / pitch[1];
  long iy_close = std::lround(iy_);
  if (coincident(iy_, iy_close))
    result[1] = (u.y > 0) ? iy_close : iy_close - 1;
  else
    result[1] = std::floor(iy_);
  result[2] = 0;
  if (M->lat_is3d[l]) {
    double iz_ = (r.z - ll[2]) / pitch[2];
    long iz_close = std::lround(iz_);
    if (coincident(iz_, iz_close))
      result[2] = (u.z > 0) ? iz_close : iz_close - 1;
    else
      result[2] = std::floor(iz_);
  }
}
Vec3 lat_local_position(int l, Vec3 r, const int* i)
{
  const double* ll = M->lat_lower_left + 3 * l;
  const double* pitch = M->lat_pitch + 3 * l;
  if (lat_is_hex(l)) { // HexLattice::get_local_position, src/lattice.cpp:983-1007 (ll = the centre)
    const int n_rings = (M->lat_shape[3 * l] + 1) / 2;
    const int n_axial = M->lat_shape[3 * l + 2];
    if (M->lat_type[l] == OMCGPU_LATTICE_HEX_Y) {
      r.x -= ll[0] + std::sqrt(3.0) / 2.0 * (i[0] - n_rings + 1) * pitch[0];
      r.y -= (ll[1] + (i[1] - n_rings + 1) * pitch[0] + (i[0] - n_rings + 1) * pitch[0] / 2.0);
    } else {
      r.x -= (ll[0] + (i[0] - n_rings + 1) * pitch[0] + (i[1] - n_rings + 1) * pitch[0] / 2.0);
      r.y -= ll[1] + std::sqrt(3.0) / 2.0 * (i[1] - n_rings + 1) * pitch[0];
    }
    if (M->lat_is3d[l])
      r.z -= ll[2] - (0.5 * n_axial - i[2] - 0.5) * pitch[1];
    return r;
  }
  r.x -= (ll[0] + (i[0] + 0.5) * pitch[0]);
  r.y -= (ll[1] + (i[1] + 0.5) * pitch[1]);
  if (M->lat_is3d[l])
    r.z -= (ll[2] + (i[2] + 0.5) * pitch[2]);
  return r;
}
// HexLattice::distance, src/lattice.cpp:738-875: r is the GLOBAL (lattice-level) position and
// i_xyz the current indices -- the distances are taken relative to the NEIGHBOUR tiles
double hex_distance(int l, Vec3 r, Vec3 u, const int* i_xyz, int* trans)
{
  const double* pitch = M->lat_pitch + 3 * l;
  const bool oy = M->lat_type[l] == OMCGPU_LATTICE_HEX_Y;
  double beta_dir, gamma_dir, delta_dir;
  if (oy) {
    beta_dir = u.x * std::sqrt(3.0) / 2.0 + u.y / 2.0;
    gamma_dir = u.x * std::sqrt(3.0) / 2.0 - u.y / 2.0;
    delta_dir = u.y;
  } else {
    beta_dir = u.x;
    gamma_dir = u.x / 2.0 - u.y * std::sqrt(3.0) / 2.0;
    delta_dir = u.x / 2.0 + u.y * std::sqrt(3.0) / 2.0;
  }
  double d {INFTY};
  trans[0] = trans[1] = trans[2] = 0;
  // beta direction
  double edge = -std::copysign(0.5 * pitch[0], beta_dir);
  Vec3 r_t;
  {
    const int s = beta_dir > 0 ? 1 : -1;
    const int t[3] = {i_xyz[0] + s, i_xyz[1], i_xyz[2]};
    r_t = lat_local_position(l, r, t);
  }
  double beta = oy ? r_t.x * std::sqrt(3.0) / 2.0 + r_t.y / 2.0 : r_t.x;
  if ((std::abs(beta - edge) > FP_PRECISION) && beta_dir != 0) {
    d = (edge - beta) / beta_dir;
    trans[0] = beta_dir > 0 ? 1 : -1;
    trans[1] = 0;
    trans[2] = 0;
  }
  // gamma direction
  edge = -std::copysign(0.5 * pitch[0], gamma_dir);
  {
    const int s = gamma_dir > 0 ? 1 : -1;
    const int t[3] = {i_xyz[0] + s, i_xyz[1] - s, i_xyz[2]};
    r_t = lat_local_position(l, r, t);
  }
  double gamma = oy ? r_t.x * std::sqrt(3.0) / 2.0 - r_t.y / 2.0 : r_t.x / 2.0 - r_t.y * std::sqrt(3.0) / 2.0;
  if ((std::abs(gamma - edge) > FP_PRECISION) && gamma_dir != 0) {
    double this_d = (edge - gamma) / gamma_dir;
    if (this_d < d) {
      trans[0] = gamma_dir > 0 ? 1 : -1;
      trans[1] = gamma_dir > 0 ? -1 : 1;
      trans[2] = 0;
      d = this_d;
    }
  }
  // delta direction
  edge = -std::copysign(0.5 * pitch[0], delta_dir);
  {
    const int s = delta_dir > 0 ? 1 : -1;
    const int t[3] = {i_xyz[0], i_xyz[1] + s, i_xyz[2]};
    r_t = lat_local_position(l, r, t);
  }
  double delta = oy ? r_t.y : r_t.x / 2.0 + r_t.y * std::sqrt(3.0) / 2.0;
  if ((std::abs(delta - edge) > FP_PRECISION) && delta_dir != 0) {
    double this_d = (edge - delta) / delta_dir;
    if (this_d < d) {
      trans[0] = 0;
      trans[1] = delta_dir > 0 ? 1 : -1;
      trans[2] = 0;
      d = this_d;
    }
  }
  if (M->lat_is3d[l]) { // top and bottom; z is LOCAL to the tile here as in the reference (r.z of the argument)
    double z = r.z;
    double z0 {std::copysign(0.5 * pitch[1], u.z)};
    if ((std::abs(z - z0) > FP_PRECISION) && u.z != 0) {
      double this_d = (z0 - z) / u.z;
      if (this_d < d) {
        d = this_d;
        trans[0] = 0;
        trans[1] = 0;
        trans[2] = u.z > 0 ? 1 : -1;
      }
    }
  }
  return d;
}
double lat_distance(int l, Vec3 r, Vec3 u, int* trans)
{
  const double* pitch = M->lat_pitch + 3 * l;
  double x0 = std::copysign(0.5 * pitch[0], u.x);
  double y0 = std::copysign(0.5 * pitch[1], u.y);
  double d = INFTY;
  trans[0] = trans[1] = trans[2] = 0;
  if ((std::abs(r.x - x0) > FP_PRECISION) && u.x != 0) {
    d = (x0 - r.x) / u.x;
    trans[0] = (u.x > 0) ? 1 : -1;
    trans[1] = 0;
    trans[2] = 0;
  }
  if ((std::abs(r.y - y0) > FP_PRECISION) && u.y != 0) {
    double this_d = (y0 - r.y) / u.y;
    if (this_d < d) {
      d = this_d;
      trans[0] = 0;
      trans[1] = (u.y > 0) ? 1 : -1;
      trans[2] = 0;
    }
  }
  if (M->lat_is3d[l]) {
    double z0 = std::copysign(0.5 * pitch[2], u.z);
    if ((std::abs(r.z - z0) > FP_PRECISION) && u.z != 0) {
      double this_d = (z0 - r.z) / u.z;
      if (this_d < d) {
        d = this_d;
        trans[0] = 0;
        trans[1] = 0;
        trans[2] = (u.z > 0) ? 1 : -1;
      }
    }
  }
  return d;
}

// =============================================================================
// Locate -- src/geometry.cpp:66-98 (instance), 102-252 (find_cell_inner),
// 256-297 (neighbor/exhaustive), 301-357 (cross_lattice); src/universe.cpp:40-60
// Neighbour lists only reorder the search of a universe's cells, so the universe
// list is searched directly (identical result for non-overlapping geometry).
// =============================================================================
int cell_instance_at_level(const Particle& p, int level)
{
  int c = p.coord[level].cell;
  int map = M->cell_distribcell_index[c];
  if (map == C_NONE)
    return C_NONE;
  int instance = 0;
  for (int i = 0; i < level; i++) {
    int c_i = p.coord[i].cell;
    if (M->cell_type[c_i] == OMCGPU_FILL_UNIVERSE) {
      instance += M->cell_offset[M->cell_offset_ofs[c_i] + map];
    } else if (M->cell_type[c_i] == OMCGPU_FILL_LATTICE) {
      instance += M->cell_offset[M->cell_offset_ofs[c_i] + map];
      int lat = p.coord[i + 1].lattice;
      if (lat_valid(lat, p.coord[i + 1].lat_idx))
        instance += lat_offset(lat, map, p.coord[i + 1].lat_idx);
    }
  }
  return instance;
}

bool universe_find_cell(Particle& p)
{
  int i_univ = p.lowest_coord().universe;
  Vec3 r = p.r_local();
  Vec3 u = p.u_local();
  for (int k = M->univ_cell_ofs[i_univ]; k < M->univ_cell_ofs[i_univ + 1]; ++k) {
    int i_cell = M->univ_cells[k];
    if (M->cell_universe[i_cell] != i_univ)
      continue;
    if (cell_contains(i_cell, r, u, p.surface)) {
      p.lowest_coord().cell = i_cell;
      return true;
    }
  }
  return false;
}

// Position::rotate with a fill cell's matrix, include/openmc/position.h:97-104
Vec3 cell_rotate(int c, Vec3 v)
{
  const double* m = M->cell_rotation + 9 * c;
  return {v.x * m[0] + v.y * m[1] + v.z * m[2], v.x * m[3] + v.y * m[4] + v.z * m[5],
    v.x * m[6] + v.y * m[7] + v.z * m[8]};
}

bool find_cell_inner(Particle& p)
{
  for (;; ++p.n_coord) {
    if (!universe_find_cell(p))
      return false;
    int i_cell = p.lowest_coord().cell;
    if (M->cell_type[i_cell] == OMCGPU_FILL_MATERIAL) {
      p.cell_instance = 0;
      if (M->cell_distribcell_index[i_cell] >= 0)
        p.cell_instance = cell_instance_at_level(p, p.n_coord - 1);
      int nmat = M->cell_mat_ofs[i_cell + 1] - M->cell_mat_ofs[i_cell];
      int nkt = M->cell_sqrtkT_ofs[i_cell + 1] - M->cell_sqrtkT_ofs[i_cell];
      p.material_last = p.material;
      p.material = M->cell_material[M->cell_mat_ofs[i_cell] + (nmat > 1 ? p.cell_instance : 0)];
      p.sqrtkT_last = p.sqrtkT;
      p.sqrtkT = M->cell_sqrtkT[M->cell_sqrtkT_ofs[i_cell] + (nkt > 1 ? p.cell_instance : 0)];
      // (a model description without the arrays -- older flattened fixtures -- has no multipliers)
      const int nd = M->cell_density_ofs ? M->cell_density_ofs[i_cell + 1] - M->cell_density_ofs[i_cell] : 0;
      p.density_mult_last = p.density_mult; // src/geometry.cpp:180-181, Cell::density_mult src/cell.cpp:104-111
      p.density_mult = nd > 0 ? M->cell_density_mult[M->cell_density_ofs[i_cell] + (nd > 1 ? p.cell_instance : 0)] : 1.0;
      return true;
    }
    if (p.n_coord >= OMCGPU_MAX_COORD) {
      mark_as_lost(p, "too many coordinate levels");
      return false;
    }
    LocalCoord& coord = p.coord[p.n_coord];
    const double* t = M->cell_translation + 3 * i_cell;
    if (M->cell_type[i_cell] == OMCGPU_FILL_UNIVERSE) {
      coord.universe = M->cell_fill[i_cell];
      coord.r = p.r_local();
      coord.u = p.u_local();
      coord.r.x -= t[0];
      coord.r.y -= t[1];
      coord.r.z -= t[2];
      if (M->cell_rotated[i_cell]) { // LocalCoord::rotate, src/geometry.cpp:200-203
        coord.r = cell_rotate(i_cell, coord.r);
        coord.u = cell_rotate(i_cell, coord.u);
      }
    } else {
      int lat = M->cell_fill[i_cell];
      coord.r = p.r_local();
      coord.u = p.u_local();
      coord.r.x -= t[0];
      coord.r.y -= t[1];
      coord.r.z -= t[2];
      if (M->cell_rotated[i_cell]) { // src/geometry.cpp:219-222
        coord.r = cell_rotate(i_cell, coord.r);
        coord.u = cell_rotate(i_cell, coord.u);
      }
      lat_get_indices(lat, coord.r, coord.u, coord.lat_idx);
      coord.r = lat_local_position(lat, coord.r, coord.lat_idx);
      coord.lattice = lat;
      if (lat_valid(lat, coord.lat_idx)) {
        coord.universe = lat_universe(lat, coord.lat_idx);
      } else if (M->lat_outer[lat] != C_NONE) {
        coord.universe = M->lat_outer[lat];
      } else {
        mark_as_lost(p, "left lattice with no outer universe");
        return false;
      }
    }
  }
}

void reset_deeper(Particle& p)
{
  for (int i = p.n_coord; i < M->n_coord_levels && i < OMCGPU_MAX_COORD; i++)
    p.coord[i].reset();
}

bool neighbor_list_find_cell(Particle& p)
{
  reset_deeper(p);
  return find_cell_inner(p);
}

bool exhaustive_find_cell(Particle& p)
{
  if (p.lowest_coord().universe == C_NONE) {
    p.coord[0].universe = M->root_universe;
    p.n_coord = 1;
  }
  reset_deeper(p);
  return find_cell_inner(p);
}

void cross_lattice(Particle& p, const BoundaryInfo& b)
{
  LocalCoord& coord = p.lowest_coord();
  int lat = coord.lattice;
  coord.lat_idx[0] += b.lat_trans[0];
  coord.lat_idx[1] += b.lat_trans[1];
  coord.lat_idx[2] += b.lat_trans[2];
  const LocalCoord& upper = p.coord[p.n_coord - 2];
  const double* t = M->cell_translation + 3 * upper.cell;
  Vec3 r = upper.r;
  r.x -= t[0];
  r.y -= t[1];
  r.z -= t[2];
  if (M->cell_rotated[upper.cell]) // src/geometry.cpp:324-326
    r = cell_rotate(upper.cell, r);
  p.r_local() = lat_local_position(lat, r, coord.lat_idx);
  if (!lat_valid(lat, coord.lat_idx)) {
    p.n_coord = 1;
    if (!exhaustive_find_cell(p))
      mark_as_lost(p, "could not be located after crossing a lattice boundary");
  } else {
    p.lowest_coord().universe = lat_universe(lat, coord.lat_idx);
    if (!exhaustive_find_cell(p)) {
      p.n_coord = 1;
      if (!exhaustive_find_cell(p))
        mark_as_lost(p, "could not be located after crossing a lattice boundary");
    }
  }
}

// src/geometry.cpp:361-453
BoundaryInfo distance_to_boundary(Particle& p)
{
  BoundaryInfo info;
  double d_lat = INFINITY;
  double d_surf = INFINITY;
  int level_surf_cross = 0;
  int level_lat_trans[3] = {0, 0, 0};
  for (int i = 0; i < p.n_coord; i++) {
    const LocalCoord& coord = p.coord[i];
    Vec3 r = coord.r, u = coord.u;
    int c = coord.cell;
    cell_distance(c, r, u, p.surface, d_surf, level_surf_cross);
    if (coord.lattice != C_NONE) {
      if (lat_is_hex(coord.lattice)) { // src/geometry.cpp:392-400: the position at the lattice's own level
        const int c_above = p.coord[i - 1].cell;
        const double* t = M->cell_translation + 3 * c_above;
        Vec3 r_hex = p.coord[i - 1].r;
        r_hex.x -= t[0];
        r_hex.y -= t[1];
        r_hex.z -= t[2];
        if (M->cell_rotated[c_above])
          r_hex = cell_rotate(c_above, r_hex);
        r_hex.z = coord.r.z;
        d_lat = hex_distance(coord.lattice, r_hex, u, coord.lat_idx, level_lat_trans);
      } else {
        d_lat = lat_distance(coord.lattice, r, u, level_lat_trans);
      }
      if (d_lat < 0)
        mark_as_lost(p, "negative distance to a lattice boundary");
    }
    double& d = info.distance;
    if (d_surf < d_lat - FP_COINCIDENT) {
      if (d == INFINITY || (d - d_surf) / d >= FP_REL_PRECISION) {
        d = d_surf;
        if (M->cell_simple[c] || d == INFTY) {
          info.surface = level_surf_cross;
        } else {
          Vec3 r_hit = r + d_surf * u;
          Vec3 nrm = surf_normal(std::abs(level_surf_cross) - 1, r_hit);
          info.surface = (dot(u, nrm) > 0) ? std::abs(level_surf_cross) : -std::abs(level_surf_cross);
        }
        info.lat_trans[0] = info.lat_trans[1] = info.lat_trans[2] = 0;
        info.coord_level = i + 1;
      }
    } else {
      if (d == INFINITY || (d - d_lat) / d >= FP_REL_PRECISION) {
        d = d_lat;
        info.surface = SURFACE_NONE;
        info.lat_trans[0] = level_lat_trans[0];
        info.lat_trans[1] = level_lat_trans[1];
        info.lat_trans[2] = level_lat_trans[2];
        info.coord_level = i + 1;
      }
    }
  }
  return info;
}

// =============================================================================
// Multigroup cross sections -- src/mgxs.cpp:599-675, 537-586
// =============================================================================
inline int mg_ba(int material, int t, int a)
{
  return M->blk_ang_ofs[M->mgm_blk_ofs[material] + t] + a;
}

int mg_temperature_index(int material, double sqrtkT)
{
  // xt::argmin(xt::abs(kTs - sqrtkT*sqrtkT)) -- first minimum
  int b0 = M->mgm_blk_ofs[material];
  int nt = M->mgm_n_temps[material];
  double kT = sqrtkT * sqrtkT;
  int best = 0;
  double bestv = std::abs(M->mgm_kTs[b0] - kT);
  for (int t = 1; t < nt; ++t) {
    double v = std::abs(M->mgm_kTs[b0 + t] - kT);
    if (v < bestv) {
      bestv = v;
      best = t;
    }
  }
  return best;
}

int mg_angle_index(int material, Vec3 u)
{
  if (M->mgm_isotropic[material])
    return 0;
  int n_pol = M->mgm_n_pol[material], n_azi = M->mgm_n_azi[material];
  double my_pol = std::acos(u.z);
  double my_azi = std::atan2(u.y, u.x);
  double delta_angle = PI / n_pol;
  int pp = std::floor(my_pol / delta_angle);
  delta_angle = 2. * PI / n_azi;
  int aa = std::floor((my_azi + PI) / delta_angle);
  return n_azi * pp + aa;
}

void mg_calculate_xs(Particle& p)
{
  int m = p.material;
  if (m != p.mg_material) {
    p.mg_t = mg_temperature_index(m, p.sqrtkT);
    p.mg_sqrtkT = p.sqrtkT;
    if (!M->mgm_isotropic[m]) {
      p.mg_a = mg_angle_index(m, p.u_local());
      p.mg_u = p.u_local();
    }
    p.mg_material = m;
  } else {
    if (p.sqrtkT != p.mg_sqrtkT) {
      p.mg_t = mg_temperature_index(m, p.sqrtkT);
      p.mg_sqrtkT = p.sqrtkT;
    }
    Vec3 ul = p.u_local();
    if (ul.x != p.mg_u.x || ul.y != p.mg_u.y || ul.z != p.mg_u.z) {
      if (!M->mgm_isotropic[m]) {
        p.mg_a = mg_angle_index(m, ul);
        p.mg_u = ul;
      }
    }
  }
  int G = M->mg_n_groups;
  int ba = mg_ba(m, p.mg_t, p.mg_a);
  p.xs_total = M->mg_total[ba * G + p.g] * p.density_mult; // src/mgxs.cpp:620-625
  p.xs_absorption = M->mg_absorption[ba * G + p.g] * p.density_mult;
  p.xs_nu_fission = M->mgm_fissionable[m] ? M->mg_nu_fission[ba * G + p.g] * p.density_mult : 0.;
}

void mg_sample_fission_energy(int m, int gin, int& dg, int& gout, uint64_t* seed, int t, int a)
{
  int G = M->mg_n_groups, DG = M->mg_n_delayed;
  int ba = mg_ba(m, t, a);
  double nu_fission = M->mg_nu_fission[ba * G + gin];
  double prob_prompt = M->mg_prompt_nu_fission[ba * G + gin];
  double xi_pd = prn(seed) * nu_fission;
  double xi_gout = prn(seed);
  if (xi_pd <= prob_prompt) {
    dg = -1;
    double prob_gout = 0.;
    for (gout = 0; gout < G; ++gout) {
      prob_gout += M->mg_chi_prompt[(ba * G + gin) * G + gout];
      if (xi_gout < prob_gout)
        break;
    }
  } else {
    for (dg = 0; dg < DG; ++dg) {
      prob_prompt += M->mg_delayed_nu_fission[(ba * DG + dg) * G + gin];
      if (xi_pd < prob_prompt)
        break;
    }
    dg = std::min(dg, DG - 1);
    double prob_gout = 0.;
    for (gout = 0; gout < G; ++gout) {
      prob_gout += M->mg_chi_delayed[((ba * DG + dg) * G + gin) * G + gout];
      if (xi_gout < prob_gout)
        break;
    }
  }
}

// src/scattdata.cpp:158-170 (sample_energy), 341-367 (Legendre), 528-558
// (histogram), 729-775 (tabular)
void mg_sample_scatter(int ba, int gin, int& gout, double& mu, double& wgt, uint64_t* seed)
{
  int G = M->mg_n_groups;
  int gmin = M->sc_gmin[ba * G + gin], gmax = M->sc_gmax[ba * G + gin];
  int e0 = M->sc_gofs[ba * G + gin];
  int order = M->sc_order[ba];
  // outgoing group
  double xi = prn(seed);
  double prob = 0.;
  int i_gout = 0;
  for (gout = gmin; gout < gmax; ++gout) {
    prob += M->sc_energy[e0 + i_gout];
    if (xi < prob)
      break;
    ++i_gout;
  }
  const double* dist = M->sc_dist + M->sc_dofs[ba * G + gin] + i_gout * order;
  const double* fmu = M->sc_fmu + M->sc_dofs[ba * G + gin] + i_gout * order;
  const double* mugrid = M->sc_mu + M->sc_mu_ofs[ba];
  int fmt = M->sc_format[ba];
  if (fmt == OMCGPU_SCATT_LEGENDRE) {
    double Mx = M->sc_maxval[e0 + i_gout];
    int samples;
    for (samples = 0; samples < MAX_SAMPLE; ++samples) {
      mu = 2. * prn(seed) - 1.;
      double f = evaluate_legendre(order - 1, dist, mu);
      if (f > 0.) {
        double uu = prn(seed) * Mx;
        if (uu <= f)
          break;
      }
    }
  } else if (fmt == OMCGPU_SCATT_HISTOGRAM) {
    double xi2 = prn(seed);
    int imu;
    if (xi2 < dist[0]) {
      imu = 0;
    } else {
      imu = std::upper_bound(dist, dist + order, xi2) - dist;
    }
    mu = prn(seed) * M->sc_dmu[ba] + mugrid[imu];
    if (mu < -1.)
      mu = -1.;
    else if (mu > 1.)
      mu = 1.;
  } else {
    int NP = order;
    double xi2 = prn(seed);
    double c_k = dist[0];
    int k;
    for (k = 0; k < NP - 1; k++) {
      double c_k1 = dist[k + 1];
      if (xi2 < c_k1)
        break;
      c_k = c_k1;
    }
    k = std::min(k, NP - 2);
    double p0 = fmu[k];
    double mu0 = mugrid[k];
    double p1 = fmu[k + 1];
    double mu1 = mugrid[k + 1];
    if (p0 == p1) {
      mu = mu0 + (xi2 - c_k) / p0;
    } else {
      double frac = (p1 - p0) / (mu1 - mu0);
      mu = mu0 + (std::sqrt(std::max(0., p0 * p0 + 2. * frac * (xi2 - c_k))) - p0) / frac;
    }
    if (mu < -1.)
      mu = -1.;
    else if (mu > 1.)
      mu = 1.;
  }
  wgt *= M->sc_mult[e0 + i_gout];
}

// =============================================================================
// fission bank -- include/openmc/shared_array.h:71-89 (append returns -1 if full)
// =============================================================================
bool bank_fission_site(const omcgpu_site& site)
{
  if ((int64_t)RS.fission_bank.size() >= RS.fission_capacity)
    return false;
  RS.fission_bank.push_back(site);
  return true;
}

void russian_roulette(Particle& p, double weight_survive)
{ // src/physics_common.cpp:12-19
  if (weight_survive * prn(p.current_seed()) < p.wgt)
    p.wgt = weight_survive;
  else
    p.wgt = 0.;
}

// =============================================================================
// Multigroup collision -- src/physics_mg.cpp:26-259
// =============================================================================
void mg_create_fission_sites(Particle& p)
{
  double weight = 1.0;
  double nu_t = p.wgt / RS.keff * weight * p.xs_nu_fission / p.xs_total;
  int nu = static_cast<int>(nu_t);
  if (prn(p.current_seed()) <= (nu_t - int(nu_t)))
    nu++;
  if (nu == 0)
    return;
  p.nu_bank.clear();
  p.fission = true;
  int n_sites_stored;
  for (n_sites_stored = 0; n_sites_stored < nu; n_sites_stored++) {
    omcgpu_site site;
    std::memset(&site, 0, sizeof(site));
    site.r[0] = p.r().x;
    site.r[1] = p.r().y;
    site.r[2] = p.r().z;
    site.particle = 0;
    site.time = p.time;
    site.wgt = 1. / weight;
    site.surf_id = SURFACE_NONE;
    site.parent_nuclide = -1;
    double mu = 2. * prn(p.current_seed()) - 1.;
    double phi = 2. * PI * prn(p.current_seed());
    site.u[0] = mu;
    site.u[1] = std::sqrt(1. - mu * mu) * std::cos(phi);
    site.u[2] = std::sqrt(1. - mu * mu) * std::sin(phi);
    int dg, gout;
    mg_sample_fission_energy(p.material, p.g, dg, gout, p.current_seed(), p.mg_t, p.mg_a);
    site.E = gout;
    site.delayed_group = dg + 1;
    if (dg != -1) {
      // get_xs(DECAY_RATE, 0, nullptr, nullptr, &dg, 0, 0): t = 0, a = 0
      double decay_rate = M->mg_decay_rate[mg_ba(p.material, 0, 0) * M->mg_n_delayed + dg];
      site.time -= std::log(prn(p.current_seed())) / decay_rate;
    }
    site.parent_id = p.id;
    site.progeny_id = p.n_progeny++;
    if (!bank_fission_site(site)) {
      p.n_progeny--;
      break;
    }
    p.nu_bank.push_back(site);
    p.delayed_group = dg + 1;
  }
  if (n_sites_stored == 0) {
    p.fission = false;
    return;
  }
  p.wgt_bank = n_sites_stored / weight; // src/physics_mg.cpp:232-233
}

void collision_mg(Particle& p)
{
  p.n_collision++;
  if (M->mat_fissionable[p.material])
    mg_create_fission_sites(p);
  if (p.xs_absorption > 0.) {
    if (M->survival_biasing) {
      double wgt_absorb = p.wgt * p.xs_absorption / p.xs_total;
      p.wgt -= wgt_absorb;
      p.k_abs += wgt_absorb * p.xs_nu_fission / p.xs_absorption;
    } else {
      if (p.xs_absorption > prn(p.current_seed()) * p.xs_total) {
        p.k_abs += p.wgt * p.xs_nu_fission / p.xs_absorption;
        p.wgt = 0.0;
        p.event_scatter = false; // TallyEvent::ABSORB, src/physics_mg.cpp:256
      }
    }
  }
  if (!p.alive())
    return;
  // scatter (src/physics_mg.cpp:83-96)
  p.event_scatter = true;
  int ba = mg_ba(p.material, p.mg_t, p.mg_a);
  mg_sample_scatter(ba, p.g_last, p.g, p.mu, p.wgt, p.current_seed());
  p.u() = rotate_angle(p.u(), p.mu, nullptr, p.current_seed());
  p.E = M->mg_energy_bin_avg[p.g];
  if (M->survival_biasing) {
    if (p.wgt < M->weight_cutoff)
      russian_roulette(p, M->weight_survive);
  }
}

// =============================================================================
// Continuous-energy cross sections -- src/material.cpp:829-901,
// src/nuclide.cpp:616-775, src/reaction.cpp:98-112
// =============================================================================
double eval_function(int f, double x)
{ // src/endf.cpp:125-134 (Polynomial), 168-226 (Tabulated1D lin-lin)
  const double* d = M->fn_data + M->fn_ofs[f];
  int n = M->fn_ofs[f + 1] - M->fn_ofs[f];
  if (M->fn_type[f] == OMCGPU_FN_POLYNOMIAL) {
    double y = 0.0;
    for (int i = n - 1; i >= 0; --i)
      y = y * x + d[i];
    return y;
  }
  int np = n / 2;
  const double* xs = d;
  const double* ys = d + np;
  if (x < xs[0])
    return ys[0];
  if (x > xs[np - 1])
    return ys[np - 1];
  int i = lower_bound_index(xs, xs + np, x);
  double x0 = xs[i], x1 = xs[i + 1], y0 = ys[i], y1 = ys[i + 1];
  double r = (x - x0) / (x1 - x0);
  return y0 + r * (y1 - y0);
}

int ce_log_index(double E)
{ // src/material.cpp:833-834.  For E within rounding of energy_max the reference's
  // index equals n_log_bins and Nuclide::calculate_xs reads grid_index[M + 1], one
  // past the end of the table (undefined behaviour that cannot be restated); both
  // the oracle and the device clamp to the last valid window [M-1, M], which holds
  // the same grid interval.  Energies below energy_min never reach the hash.
  int i = std::log(E / M->energy_min) / M->log_spacing;
  if (i > M->n_log_bins - 1)
    i = M->n_log_bins - 1;
  if (i < 0)
    i = 0;
  return i;
}

void ce_micro_xs(int i_nuclide, int i_log_union, double E, MicroXS& micro)
{ // src/nuclide.cpp:676-775 (single temperature -> i_temp = 0)
  const int64_t g0 = M->nuc_grid_ofs[i_nuclide];
  const int64_t n_grid = M->nuc_grid_ofs[i_nuclide + 1] - g0;
  const double* energy = M->nuc_energy + g0;
  const double* xs = M->nuc_xs + 4 * g0;
  const int32_t* grid_index = M->nuc_grid_index + (int64_t)i_nuclide * (M->n_log_bins + 1);
  int i_grid;
  if (E < energy[0]) {
    i_grid = 0;
  } else if (E > energy[n_grid - 1]) {
    i_grid = n_grid - 2;
  } else {
    int i_low = grid_index[i_log_union];
    int i_high = grid_index[i_log_union + 1] + 1;
    i_grid = i_low + lower_bound_index(&energy[i_low], &energy[i_high], E);
  }
  if (energy[i_grid] == energy[i_grid + 1])
    ++i_grid;
  double f = (E - energy[i_grid]) / (energy[i_grid + 1] - energy[i_grid]);
  micro.index_grid = i_grid;
  micro.interp_factor = f;
  micro.total = (1.0 - f) * xs[4 * i_grid + 0] + f * xs[4 * (i_grid + 1) + 0];
  micro.absorption = (1.0 - f) * xs[4 * i_grid + 1] + f * xs[4 * (i_grid + 1) + 1];
  if (M->nuc_fissionable[i_nuclide]) {
    micro.fission = (1.0 - f) * xs[4 * i_grid + 2] + f * xs[4 * (i_grid + 1) + 2];
    micro.nu_fission = (1.0 - f) * xs[4 * i_grid + 3] + f * xs[4 * (i_grid + 1) + 3];
  } else {
    micro.fission = 0.0;
    micro.nu_fission = 0.0;
  }
  micro.elastic = -1.0;
}

double ce_reaction_xs(int rx, const MicroXS& micro)
{ // src/reaction.cpp:98-112
  int thr = M->rx_threshold[rx];
  if (micro.index_grid < thr)
    return 0.0;
  const double* v = M->rx_xs + M->rx_xs_ofs[rx];
  return (1.0 - micro.interp_factor) * v[micro.index_grid - thr] +
         micro.interp_factor * v[micro.index_grid - thr + 1];
}

// temperature pick of Nuclide::calculate_xs, TemperatureMethod::NEAREST (src/nuclide.cpp:682-691):
// the table of material entry d (one nuclide at one loaded temperature) nearest to the cell's kT
int nuc_table(int d, double sqrtkT)
{
  if (M->n_nuclide_temps == 0)
    return d;
  const double kT = sqrtkT * sqrtkT;
  int i_temp = M->nuct_ofs[d];
  double max_diff = INFINITY;
  for (int t = M->nuct_ofs[d]; t < M->nuct_ofs[d + 1]; ++t) {
    const double diff = std::abs(M->nuc_kT[t] - kT);
    if (diff < max_diff) {
      i_temp = t;
      max_diff = diff;
    }
  }
  return i_temp;
}

void ce_calculate_xs(Particle& p, int32_t* i_grid_out = nullptr, int max_nuc = 0)
{
  p.xs_total = p.xs_absorption = p.xs_fission = p.xs_nu_fission = 0.0;
  int i_log = ce_log_index(p.E);
  int m = p.material;
  for (int k = M->mat_nuc_ofs[m]; k < M->mat_nuc_ofs[m + 1]; ++k) {
    MicroXS micro;
    ce_micro_xs(nuc_table(M->mat_nuclide[k], p.sqrtkT), i_log, p.E, micro);
    double atom_density = M->mat_atom_density[k] * p.density_mult; // Material::atom_density(i, rho_multiplier)
    p.xs_total += atom_density * micro.total;
    p.xs_absorption += atom_density * micro.absorption;
    p.xs_fission += atom_density * micro.fission;
    p.xs_nu_fission += atom_density * micro.nu_fission;
    int j = k - M->mat_nuc_ofs[m];
    if (i_grid_out && j < max_nuc)
      i_grid_out[j] = micro.index_grid;
  }
}

// =============================================================================
// Continuous-energy collision -- src/physics.cpp:93-285, 494-518, 638-818,
// 990-1167
// =============================================================================
double ce_nu(int i_nuclide, double E, int mode /*0 prompt 1 delayed 2 total*/)
{ // src/nuclide.cpp:535-576
  if (!M->nuc_fissionable[i_nuclide])
    return 0.0;
  int rx = M->nuc_rx_ofs[i_nuclide] + M->nuc_fission_rx[i_nuclide];
  int p0 = M->rx_prod_ofs[rx], p1 = M->rx_prod_ofs[rx + 1];
  switch (mode) {
  case 0:
    return eval_function(M->prod_yield_fn[p0], E);
  case 1:
    if (M->nuc_n_precursor[i_nuclide] > 0 && M->create_delayed_neutrons) {
      double nu = 0.0;
      for (int i = p0 + 1; i < p1; ++i)
        if (M->prod_emission[i] == 1)
          nu += eval_function(M->prod_yield_fn[i], E);
      return nu;
    }
    return 0.0;
  default:
    if (M->nuc_total_nu_fn[i_nuclide] >= 0 && M->create_delayed_neutrons)
      return eval_function(M->nuc_total_nu_fn[i_nuclide], E);
    return eval_function(M->prod_yield_fn[p0], E);
  }
}

void ce_sample_product(int prod, double E_in, double& E_out, double& mu, uint64_t* seed)
{ // src/reaction_product.cpp:109-130 -> src/secondary_uncorrelated.cpp:53-66
  mu = uniform_distribution(-1., 1., seed); // no angle distribution: isotropic
  const double* par = M->prod_par + 2 * prod;
  const int32_t* fn = M->prod_fn + 2 * prod;
  switch (M->prod_elaw[prod]) {
  case OMCGPU_ELAW_LEVEL: // src/distribution_energy.cpp:49-52
    E_out = par[1] * (E_in - par[0]);
    break;
  case OMCGPU_ELAW_MAXWELL: { // :284-298
    double theta = eval_function(fn[0], E_in);
    while (true) {
      E_out = maxwell_spectrum(theta, seed);
      if (E_out <= E_in - par[0])
        break;
    }
    break;
  }
  case OMCGPU_ELAW_EVAPORATION: { // :309-327
    double theta = eval_function(fn[0], E_in);
    double y = (E_in - par[0]) / theta;
    double v = 1.0 - std::exp(-y);
    double x;
    while (true) {
      x = -std::log((1.0 - v * prn(seed)) * (1.0 - v * prn(seed)));
      if (x <= y)
        break;
    }
    E_out = x * theta;
    break;
  }
  case OMCGPU_ELAW_WATT: { // :347-361
    double a = eval_function(fn[0], E_in);
    double b = eval_function(fn[1], E_in);
    while (true) {
      E_out = watt_spectrum(a, b, seed);
      if (E_out <= E_in - par[0])
        break;
    }
    break;
  }
  default:
    E_out = E_in;
  }
}

void ce_sample_fission_neutron(int i_nuclide, int rx, omcgpu_site* site, Particle& p)
{ // src/physics.cpp:1043-1117
  double E_in = p.E;
  uint64_t* seed = p.current_seed();
  double nu_t = ce_nu(i_nuclide, E_in, 2);
  double nu_d = ce_nu(i_nuclide, E_in, 1);
  double beta = nu_d / nu_t;
  int p0 = M->rx_prod_ofs[rx];
  if (prn(seed) < beta) {
    double xi = prn(seed) * nu_d;
    double prob = 0.0;
    int group;
    int n_prec = M->nuc_n_precursor[i_nuclide];
    for (group = 1; group < n_prec; ++group) {
      double yield = eval_function(M->prod_yield_fn[p0 + group], E_in);
      prob += yield;
      if (xi < prob)
        break;
    }
    group = std::min(group, n_prec);
    site->delayed_group = group;
    double decay_rate = M->prod_decay_rate[p0 + group];
    site->time -= std::log(prn(seed)) / decay_rate;
  } else {
    site->delayed_group = 0;
  }
  int n_sample = 0;
  double mu;
  while (true) {
    ce_sample_product(p0 + site->delayed_group, E_in, site->E, mu, seed);
    if (site->E < M->energy_max)
      break;
    if (++n_sample == MAX_SAMPLE)
      break;
  }
  Vec3 u = rotate_angle(p.u(), mu, nullptr, seed);
  site->u[0] = u.x;
  site->u[1] = u.y;
  site->u[2] = u.z;
}

void ce_create_fission_sites(Particle& p, int i_nuclide, int rx, const MicroXS& micro)
{ // src/physics.cpp:170-285
  double weight = 1.0;
  double nu_t = p.wgt / RS.keff * weight * micro.nu_fission / micro.total;
  int nu = static_cast<int>(nu_t);
  if (prn(p.current_seed()) <= (nu_t - nu))
    ++nu;
  if (nu == 0)
    return;
  p.nu_bank.clear();
  p.fission = true;
  int n_sites_stored;
  for (n_sites_stored = 0; n_sites_stored < nu; n_sites_stored++) {
    omcgpu_site site;
    std::memset(&site, 0, sizeof(site));
    site.r[0] = p.r().x;
    site.r[1] = p.r().y;
    site.r[2] = p.r().z;
    site.particle = 0;
    site.time = p.time;
    site.wgt = 1. / weight;
    site.surf_id = 0;
    site.parent_nuclide = -1;
    ce_sample_fission_neutron(i_nuclide, rx, &site, p);
    site.parent_id = p.id;
    site.progeny_id = p.n_progeny++;
    if (!bank_fission_site(site)) {
      p.n_progeny--;
      break;
    }
    p.nu_bank.push_back(site);
  }
  if (n_sites_stored == 0) {
    p.fission = false;
    return;
  }
  p.wgt_bank = n_sites_stored / weight; // src/physics.cpp:280-281
}

Vec3 sample_cxs_target_velocity(double awr, double E, Vec3 u, double kT, uint64_t* seed)
{ // src/physics.cpp:990-1041
  double beta_vn = std::sqrt(awr * E / kT);
  double alpha = 1.0 / (1.0 + std::sqrt(PI) * beta_vn / 2.0);
  double beta_vt_sq;
  double mu;
  while (true) {
    double r1 = prn(seed);
    double r2 = prn(seed);
    if (prn(seed) < alpha) {
      beta_vt_sq = -std::log(r1 * r2);
    } else {
      double c = std::cos(PI / 2.0 * prn(seed));
      beta_vt_sq = -std::log(r1) - std::log(r2) * c * c;
    }
    double beta_vt = std::sqrt(beta_vt_sq);
    mu = uniform_distribution(-1., 1., seed);
    double accept_prob =
      std::sqrt(beta_vn * beta_vn + beta_vt_sq - 2 * beta_vn * beta_vt * mu) / (beta_vn + beta_vt);
    if (prn(seed) < accept_prob)
      break;
  }
  double vt = std::sqrt(beta_vt_sq * kT / awr);
  return vt * rotate_angle(u, mu, nullptr, seed);
}

void ce_elastic_scatter(int i_nuclide, double kT, Particle& p)
{ // src/physics.cpp:754-818 ; target velocity :836-866 (non-resonant branch)
  double vel = std::sqrt(p.E);
  double awr = M->nuc_awr[i_nuclide];
  Vec3 v_n = vel * p.u();
  Vec3 v_t {0, 0, 0};
  if (!(p.E >= M->free_gas_threshold * kT && awr > 1.0))
    v_t = sample_cxs_target_velocity(awr, p.E, p.u(), kT, p.current_seed());
  Vec3 v_cm = (v_n + awr * v_t) / (awr + 1.0);
  v_n = v_n - v_cm;
  vel = norm(v_n);
  double mu_cm = uniform_distribution(-1., 1., p.current_seed());
  Vec3 u_cm = v_n / vel;
  v_n = vel * rotate_angle(u_cm, mu_cm, nullptr, p.current_seed());
  v_n = v_n + v_cm;
  p.E = dot(v_n, v_n);
  vel = std::sqrt(p.E);
  p.mu = dot(p.u(), v_n) / vel;
  p.u() = v_n / vel;
  if (std::abs(p.mu) > 1.0)
    p.mu = std::copysign(1.0, p.mu);
}

void ce_inelastic_scatter(int i_nuclide, int rx, Particle& p)
{ // src/physics.cpp:1119-1167
  double E_in = p.E;
  double E, mu;
  int prod = M->rx_prod_ofs[rx];
  ce_sample_product(prod, E_in, E, mu, p.current_seed());
  if (M->rx_scatter_in_cm[rx]) {
    double E_cm = E;
    double A = M->nuc_awr[i_nuclide];
    E = E_cm + (E_in + 2.0 * mu * (A + 1.0) * std::sqrt(E_in * E_cm)) / ((A + 1.0) * (A + 1.0));
    mu = mu * std::sqrt(E_cm / E) + 1.0 / (A + 1.0) * std::sqrt(E_in / E);
  }
  if (std::abs(mu) > 1.0)
    mu = std::copysign(1.0, mu);
  p.E = E;
  p.mu = mu;
  p.u() = rotate_angle(p.u(), mu, nullptr, p.current_seed());
  double yield = eval_function(M->prod_yield_fn[prod], E_in);
  if (std::floor(yield) == yield && yield > 0) {
    // an integral yield creates exactly yield - 1 secondary neutrons (:1158-1162)
    for (int i = 0; i < static_cast<int>(std::round(yield)) - 1; ++i) {
      if (p.E < M->energy_cutoff) // create_secondary, src/particle.cpp:78-82
        continue;
      omcgpu_site site;
      std::memset(&site, 0, sizeof(site));
      site.r[0] = p.r().x;
      site.r[1] = p.r().y;
      site.r[2] = p.r().z;
      site.u[0] = p.u().x;
      site.u[1] = p.u().y;
      site.u[2] = p.u().z;
      site.E = p.E;
      site.wgt = p.wgt;
      site.time = p.time;
      site.surf_id = SURFACE_NONE;
      p.secondary_bank.push_back(site);
    }
  } else {
    p.wgt *= yield;
  }
}

void collision_ce(Particle& p)
{ // src/physics.cpp:43-168
  ++p.n_collision;
  int m = p.material;
  int i_log = ce_log_index(p.E);
  // sample_nuclide (:494-518) -- the reference reads the per-nuclide cache that
  // calculate_xs filled at this energy; recomputing gives the same numbers
  double cutoff = prn(p.current_seed()) * p.xs_total;
  double prob = 0.0;
  int i_nuclide = -1;
  MicroXS micro;
  for (int k = M->mat_nuc_ofs[m]; k < M->mat_nuc_ofs[m + 1]; ++k) {
    const int table = nuc_table(M->mat_nuclide[k], p.sqrtkT);
    ce_micro_xs(table, i_log, p.E, micro);
    prob += M->mat_atom_density[k] * p.density_mult * micro.total; // src/physics.cpp:507
    if (prob >= cutoff) {
      i_nuclide = table; // every nuc_* / rx_* array is indexed by table
      break;
    }
  }
  if (i_nuclide < 0) {
    mark_as_lost(p, "Did not sample any nuclide during collision.");
    return;
  }
  p.event_micro = micro; // p.event_nuclide() = i_nuclide, :99
  int rx0 = M->nuc_rx_ofs[i_nuclide];
  if (M->nuc_fissionable[i_nuclide] && micro.fission > 0.0) {
    int rx = rx0 + M->nuc_fission_rx[i_nuclide]; // no partial fission reactions
    ce_create_fission_sites(p, i_nuclide, rx, micro);
  }
  if (micro.absorption > 0.0) { // absorption (:638-672)
    if (M->survival_biasing) {
      const double wgt_absorb = p.wgt * micro.absorption / micro.total;
      p.wgt -= wgt_absorb;
      p.k_abs += wgt_absorb * micro.nu_fission / micro.absorption;
    } else {
      if (micro.absorption > prn(p.current_seed()) * micro.total) {
        p.k_abs += p.wgt * micro.nu_fission / micro.absorption;
        p.wgt = 0.0;
        p.event_scatter = false; // TallyEvent::ABSORB, :666
      }
    }
  }
  if (!p.alive())
    return;
  // scatter (:674-752)
  p.event_scatter = true; // :740
  {
    double cut = prn(p.current_seed()) * (micro.total - micro.absorption);
    micro.elastic = ce_reaction_xs(rx0, micro); // calculate_elastic_xs, src/nuclide.cpp:578-590
    double pr = micro.elastic;
    if (pr > cut) {
      ce_elastic_scatter(i_nuclide, M->nuc_kT[i_nuclide], p);
    } else {
      int i = rx0;
      int n_rx = M->nuc_rx_ofs[i_nuclide + 1];
      for (int j = rx0; j < n_rx && pr < cut; ++j) {
        if (!M->rx_inelastic[j])
          continue;
        i = j;
        pr += ce_reaction_xs(j, micro);
      }
      ce_inelastic_scatter(i_nuclide, i, p);
    }
  }
  if (M->survival_biasing) {
    if (p.wgt < M->weight_cutoff)
      russian_roulette(p, M->weight_survive);
  }
}

// =============================================================================
// Tallies (tracklength) -- src/tallies/tally_scoring.cpp:31-141, 578-700,
// 1622-1990, 2416-2541; src/mesh.cpp:874-929, 999-1146, 1328-1369;
// src/tallies/filter_{mesh,energy,cell,material,universe}.cpp
// =============================================================================
struct Match {
  std::vector<int> bins;
  std::vector<double> weights;
};

struct MeshDistance {
  int next_index {-1};
  bool max_surface {true};
  double distance {INFTY};
};

int mesh_index_in_direction(int m, double r, int i)
{
  return std::ceil((r - M->mesh_lower_left[3 * m + i]) / M->mesh_width[3 * m + i]);
}
void mesh_get_indices(int m, Vec3 r, int* ijk, bool& in_mesh)
{
  in_mesh = true;
  int n = M->mesh_ndim[m];
  for (int i = 0; i < n; ++i) {
    ijk[i] = mesh_index_in_direction(m, r[i], i);
    if (ijk[i] < 1 || ijk[i] > M->mesh_shape[3 * m + i])
      in_mesh = false;
  }
}
int mesh_bin_from_indices(int m, const int* ijk)
{
  const int32_t* s = M->mesh_shape + 3 * m;
  switch (M->mesh_ndim[m]) {
  case 1:
    return ijk[0] - 1;
  case 2:
    return (ijk[1] - 1) * s[0] + ijk[0] - 1;
  default:
    return ((ijk[2] - 1) * s[1] + (ijk[1] - 1)) * s[0] + ijk[0] - 1;
  }
}
MeshDistance mesh_distance_to_grid_boundary(int m, const int* ijk, int i, Vec3 r0, Vec3 u, double)
{
  MeshDistance d;
  d.next_index = ijk[i];
  if (std::abs(u[i]) < FP_PRECISION)
    return d;
  d.max_surface = (u[i] > 0);
  const double ll = M->mesh_lower_left[3 * m + i], w = M->mesh_width[3 * m + i];
  if (d.max_surface && (ijk[i] <= M->mesh_shape[3 * m + i])) {
    d.next_index++;
    d.distance = ((ll + ijk[i] * w) - r0[i]) / u[i];
  } else if (!d.max_surface && (ijk[i] >= 1)) {
    d.next_index--;
    d.distance = ((ll + (ijk[i] - 1) * w) - r0[i]) / u[i];
  }
  return d;
}

void mesh_bins_crossed(int m, Vec3 r0, Vec3 r1, Vec3 u, Match& match)
{
  double total_distance = norm(r1 - r0);
  if (total_distance == 0.0)
    return;
  const int n = M->mesh_ndim[m];
  bool in_mesh;
  double traveled_distance = 0.0;
  int ijk[3] = {0, 0, 0};
  mesh_get_indices(m, r0 + TINY_BIT * u, ijk, in_mesh);
  if (total_distance < 2 * TINY_BIT) {
    if (in_mesh) {
      match.bins.push_back(mesh_bin_from_indices(m, ijk));
      match.weights.push_back(1.0);
    }
    return;
  }
  MeshDistance distances[3];
  for (int k = 0; k < n; ++k)
    distances[k] = mesh_distance_to_grid_boundary(m, ijk, k, r0, u, 0.0);
  while (true) {
    if (in_mesh) {
      // std::min_element over all 3 entries (unused dimensions stay at INFTY)
      int k = 0;
      for (int q = 1; q < 3; ++q)
        if (distances[q].distance < distances[k].distance)
          k = q;
      match.bins.push_back(mesh_bin_from_indices(m, ijk));
      match.weights.push_back(
        (std::min(distances[k].distance, total_distance) - traveled_distance) / total_distance);
      traveled_distance = distances[k].distance;
      if (traveled_distance >= total_distance)
        return;
      ijk[k] = distances[k].next_index;
      distances[k] = mesh_distance_to_grid_boundary(m, ijk, k, r0, u, traveled_distance);
      in_mesh = ((ijk[k] >= 1) && (ijk[k] <= M->mesh_shape[3 * m + k]));
    } else {
      int k_max = -1;
      for (int k = 0; k < n; ++k) {
        if ((ijk[k] < 1 || ijk[k] > M->mesh_shape[3 * m + k]) &&
            (distances[k].distance > traveled_distance)) {
          traveled_distance = distances[k].distance;
          k_max = k;
        }
      }
      if (k_max == -1)
        traveled_distance += TINY_BIT;
      if (traveled_distance >= total_distance)
        return;
      mesh_get_indices(m, r0 + (traveled_distance + TINY_BIT) * u, ijk, in_mesh);
      for (int k = 0; k < n; ++k)
        distances[k] = mesh_distance_to_grid_boundary(m, ijk, k, r0, u, traveled_distance);
    }
  }
}

// `tracklength`: the estimator of the tally being scored (MeshFilter::get_all_bins and
// EnergyFilter::get_all_bins differ for the collision estimator: src/tallies/filter_mesh.cpp:61-69,
// src/tallies/filter_energy.cpp:62-67)
void filter_get_all_bins(int f, Particle& p, Match& match, bool tracklength = true)
{
  match.bins.clear();
  match.weights.clear();
  const int32_t* idata = M->filter_idata + M->filter_data_ofs[f];
  const double* ddata = M->filter_ddata + M->filter_data_ofs[f];
  switch (M->filter_type[f]) {
  case OMCGPU_FILTER_MESH:
    if (tracklength) {
      mesh_bins_crossed(M->filter_mesh[f], p.r_last, p.r(), p.u(), match);
    } else { // StructuredMesh::get_bin(r), src/mesh.cpp:874-929
      int ijk[3] = {0, 0, 0};
      bool in_mesh;
      mesh_get_indices(M->filter_mesh[f], p.r(), ijk, in_mesh);
      if (in_mesh) {
        match.bins.push_back(mesh_bin_from_indices(M->filter_mesh[f], ijk));
        match.weights.push_back(1.0);
      }
    }
    break;
  case OMCGPU_FILTER_ENERGY: {
    int nb = M->filter_n_bins[f];
    if (!M->run_CE && M->filter_matches_groups[f]) {
      match.bins.push_back(M->mg_n_groups - (tracklength ? p.g : p.g_last) - 1);
      match.weights.push_back(1.0);
    } else {
      double E = p.E_last;
      if (E >= ddata[0] && E <= ddata[nb]) {
        match.bins.push_back(lower_bound_index(ddata, ddata + nb + 1, E));
        match.weights.push_back(1.0);
      }
    }
    break;
  }
  case OMCGPU_FILTER_ENERGYOUT: { // EnergyoutFilter::get_all_bins, src/tallies/filter_energy.cpp:100-125
    int nb = M->filter_n_bins[f];
    if (!M->run_CE && M->filter_matches_groups[f]) {
      match.bins.push_back(M->mg_n_groups - p.g - 1);
      match.weights.push_back(1.0);
    } else if (p.E >= ddata[0] && p.E <= ddata[nb]) {
      match.bins.push_back(lower_bound_index(ddata, ddata + nb + 1, p.E));
      match.weights.push_back(1.0);
    }
    break;
  }
  case OMCGPU_FILTER_CELL:
    for (int i = 0; i < p.n_coord; i++) {
      int b = idata[p.coord[i].cell];
      if (b >= 0) {
        match.bins.push_back(b);
        match.weights.push_back(1.0);
      }
    }
    break;
  case OMCGPU_FILTER_MATERIAL:
    if (p.material >= 0 && idata[p.material] >= 0) {
      match.bins.push_back(idata[p.material]);
      match.weights.push_back(1.0);
    }
    break;
  case OMCGPU_FILTER_UNIVERSE:
    for (int i = 0; i < p.n_coord; i++) {
      int b = idata[p.coord[i].universe];
      if (b >= 0) {
        match.bins.push_back(b);
        match.weights.push_back(1.0);
      }
    }
    break;
  case OMCGPU_FILTER_DISTRIBCELL: { // DistribcellFilter::get_all_bins, src/tallies/filter_distribcell.cpp:40-63
    const int cell = idata[0];
    const int map = M->cell_distribcell_index[cell];
    int offset = 0;
    for (int i = 0; i < p.n_coord; i++) {
      const int c = p.coord[i].cell;
      if (M->cell_type[c] == OMCGPU_FILL_UNIVERSE) {
        offset += M->cell_offset[M->cell_offset_ofs[c] + map];
      } else if (M->cell_type[c] == OMCGPU_FILL_LATTICE) {
        const int lat = p.coord[i + 1].lattice;
        const int* li = p.coord[i + 1].lat_idx;
        const int32_t* n = M->lat_shape + 3 * lat;
        if (li[0] >= 0 && li[0] < n[0] && li[1] >= 0 && li[1] < n[1] && li[2] >= 0 && li[2] < n[2])
          offset += M->lat_offsets[M->lat_offset_ofs[lat] + n[0] * n[1] * n[2] * map +
                                   n[0] * n[1] * li[2] + n[0] * li[1] + li[0]] +
                    M->cell_offset[M->cell_offset_ofs[c] + map];
      }
      if (cell == c) {
        match.bins.push_back(offset);
        match.weights.push_back(1.0);
        break;
      }
    }
    break;
  }
  }
}

double mg_scatter_xs(int ba, int g)
{ // ScattData::get_xs(SCATTER, gin, nullptr, ..) src/scattdata.cpp:174-225
  int G = M->mg_n_groups;
  int e0 = M->sc_gofs[ba * G + g];
  int n = M->sc_gmax[ba * G + g] - M->sc_gmin[ba * G + g] + 1;
  double ip = 0.0;
  for (int i = 0; i < n; ++i)
    ip = ip + M->sc_mult[e0 + i] * M->sc_energy[e0 + i];
  return M->sc_scattxs[ba * G + g] / ip;
}

// score_tracklength_tally_general / score_collision_tally, src/tallies/tally_scoring.cpp:2416-2490,
// 2543-2620: the same filter / score machinery for the tallies of one estimator
void score_tallies(Particle& p, double flux, int estimator);
void score_tracklength_tally(Particle& p, double distance)
{
  score_tallies(p, p.wgt * distance, OMCGPU_ESTIMATOR_TRACKLENGTH);
}
void score_collision_tally(Particle& p)
{ // flux = wgt_last / Sigma_t (:2546-2549); called after the collision, which leaves macro_xs alone
  score_tallies(p, p.wgt_last / p.xs_total, OMCGPU_ESTIMATOR_COLLISION);
}
void score_analog_tally(Particle& p)
{ // score_analog_tally_ce / _mg, :2314-2414: flux = 1 for neutrons; after the collision
  score_tallies(p, 1.0, OMCGPU_ESTIMATOR_ANALOG);
}
void score_tallies(Particle& p, double flux, int estimator)
{
  const bool tracklength = estimator == OMCGPU_ESTIMATOR_TRACKLENGTH;
  const bool analog = estimator == OMCGPU_ESTIMATOR_ANALOG;
  int G = M->mg_n_groups;
  std::vector<Match> matches(M->n_filters);
  std::vector<char> present(M->n_filters, 0);
  for (int t = 0; t < M->n_tallies; ++t) {
    if (M->tally_estimator[t] != estimator)
      continue;
    int nf = M->tally_n_filters[t];
    const int32_t* filt = M->tally_filters + OMCGPU_MAX_FILTERS * t;
    const int32_t* strides = M->tally_strides + OMCGPU_MAX_FILTERS * t;
    bool empty = false;
    for (int i = 0; i < nf; ++i) {
      int f = filt[i];
      if (!present[f]) {
        filter_get_all_bins(f, p, matches[f], tracklength);
        present[f] = 1;
      }
      if (matches[f].bins.empty()) {
        empty = true;
        break;
      }
    }
    if (empty)
      continue;
    int ns = M->tally_n_scores[t];
    const int32_t* scores = M->tally_scores + OMCGPU_MAX_SCORES * t;
    double* results = RS.tally_values.data() + M->tally_result_ofs[t];
    int eo_pos = -1; // tally.energyout_filter_
    for (int i = 0; i < nf; ++i)
      if (M->filter_type[filt[i]] == OMCGPU_FILTER_ENERGYOUT)
        eo_pos = i;
    const bool skip_all_eout = false;
    int ibin[OMCGPU_MAX_FILTERS] = {0, 0, 0};
    while (true) {
      int64_t index = 0;
      double weight = 1.;
      for (int i = 0; i < nf; ++i) {
        const Match& mt = matches[filt[i]];
        index += (int64_t)mt.bins[ibin[i]] * strides[i];
        weight *= mt.weights[ibin[i]];
      }
      for (int s = 0; s < ns; ++s) {
        double score = 0.0;
        bool skip = false;
        if (M->run_CE && analog) { // score_general_ce_analog :1116-1330
          const MicroXS& em = p.event_micro;
          const double wgt_absorb = M->survival_biasing ? p.wgt_last * em.absorption / em.total : 0.0;
          switch (scores[s]) {
          case OMCGPU_SCORE_FLUX:
            score = flux * p.wgt_last / p.xs_total;
            break;
          case OMCGPU_SCORE_TOTAL:
            score = p.wgt_last * flux;
            break;
          case OMCGPU_SCORE_SCATTER:
            if (!p.event_scatter)
              skip = true;
            score = (p.wgt_last - wgt_absorb) * flux;
            break;
          case OMCGPU_SCORE_ABSORPTION:
            if (M->survival_biasing) {
              score = wgt_absorb * flux;
            } else {
              if (p.event_scatter)
                skip = true;
              score = p.wgt_last * flux;
            }
            break;
          case OMCGPU_SCORE_FISSION:
            if (p.xs_fission == 0) {
              skip = true;
            } else if (M->survival_biasing) {
              score = em.total > 0 ? p.wgt_last * em.fission / em.total * flux : 0.;
            } else {
              if (p.event_scatter)
                skip = true;
              score = p.wgt_last * em.fission / em.absorption * flux;
            }
            break;
          case OMCGPU_SCORE_NU_FISSION:
            if (p.xs_fission == 0) {
              skip = true;
            } else if (M->survival_biasing) {
              score = em.total > 0 ? p.wgt_last * em.nu_fission / em.total * flux : 0.;
            } else {
              if (!p.fission)
                skip = true;
              score = RS.keff * p.wgt_bank * flux; // sites banked, weighted by 1 / keff
            }
            break;
          }
        } else if (M->run_CE) { // score_general_ce_nonanalog :578-700
          switch (scores[s]) {
          case OMCGPU_SCORE_FLUX:
            score = flux;
            break;
          case OMCGPU_SCORE_TOTAL:
            score = p.xs_total * flux;
            break;
          case OMCGPU_SCORE_SCATTER:
            score = (p.xs_total - p.xs_absorption) * flux;
            break;
          case OMCGPU_SCORE_ABSORPTION:
            score = p.xs_absorption * flux;
            break;
          case OMCGPU_SCORE_FISSION:
            if (p.xs_fission == 0)
              skip = true;
            score = p.xs_fission * flux;
            break;
          case OMCGPU_SCORE_NU_FISSION:
            if (p.xs_fission == 0)
              skip = true;
            score = p.xs_nu_fission * flux;
            break;
          }
        } else if (analog) { // score_general_mg :1622-1990, estimator_ == ANALOG, i_nuclide = -1
          const int m = p.material;
          const double wgt_absorb = M->survival_biasing ? p.wgt_last * p.xs_absorption / p.xs_total : 0.0;
          // :1632-1657: with survival biasing a particle killed by the roulette after its scatter is
          // scored with its NEW group; every other case resolves to the pre-collision group
          const int p_g = (M->survival_biasing && !p.alive()) ? p.g : p.g_last;
          const int ba = mg_ba(m, p.mg_t, mg_angle_index(m, p.u_last_local));
          const bool fis = M->mgm_fissionable[m];
          switch (scores[s]) {
          case OMCGPU_SCORE_FLUX:
            score = flux * p.wgt_last / p.xs_total;
            break;
          case OMCGPU_SCORE_TOTAL:
            score = flux * p.wgt_last;
            break;
          case OMCGPU_SCORE_SCATTER:
            if (!p.event_scatter)
              skip = true;
            score = (p.wgt_last - wgt_absorb) * flux;
            break;
          case OMCGPU_SCORE_ABSORPTION:
            if (M->survival_biasing) {
              score = wgt_absorb * flux;
            } else {
              if (p.event_scatter)
                skip = true;
              score = p.wgt_last * flux;
            }
            break;
          case OMCGPU_SCORE_FISSION:
            if (M->survival_biasing) {
              score = wgt_absorb * flux;
            } else {
              if (p.event_scatter)
                skip = true;
              score = p.wgt_last * flux;
            }
            score *= (fis ? M->mg_fission[ba * G + p_g] : 0.) / M->mg_absorption[ba * G + p_g];
            break;
          case OMCGPU_SCORE_NU_FISSION:
            if (M->survival_biasing) {
              score = wgt_absorb * flux;
              score *= (fis ? M->mg_nu_fission[ba * G + p_g] : 0.) / M->mg_absorption[ba * G + p_g];
            } else {
              if (!p.fission)
                skip = true;
              score = RS.keff * p.wgt_bank * flux;
            }
            break;
          }
        } else { // score_general_mg :1622-1990 (tracklength branch, i_nuclide = -1)
          int m = p.material;
          // tracklength: no collision happened, g and the local direction are current; collision
          // estimator: the pre-collision group and direction (:1632-1663 -- whether the particle
          // scattered or not, that is what p_g / p_u resolve to)
          // (:1632-1657: with survival biasing a particle the roulette killed after its scatter is
          // scored with its NEW group)
          const int p_g = tracklength ? p.g : ((M->survival_biasing && !p.alive()) ? p.g : p.g_last);
          int ba = (m >= 0) ? mg_ba(m, p.mg_t, mg_angle_index(m, tracklength ? p.u_local() : p.u_last_local)) : 0;
          switch (scores[s]) {
          case OMCGPU_SCORE_FLUX:
            score = flux;
            break;
          case OMCGPU_SCORE_TOTAL:
            score = p.xs_total * flux;
            break;
          case OMCGPU_SCORE_ABSORPTION:
            score = p.xs_absorption * flux;
            break;
          case OMCGPU_SCORE_SCATTER:
            score = (m >= 0) ? flux * mg_scatter_xs(ba, p_g) : 0.0;
            break;
          case OMCGPU_SCORE_FISSION:
            score = (m >= 0 && M->mgm_fissionable[m]) ? flux * M->mg_fission[ba * G + p_g] : flux * 0.;
            break;
          case OMCGPU_SCORE_NU_FISSION:
            score =
              (m >= 0 && M->mgm_fissionable[m]) ? flux * M->mg_nu_fission[ba * G + p_g] : flux * 0.;
            break;
          }
        }
        // score_fission_eout, :383-470: nu-fission of a tally with an EnergyoutFilter is scored once per
        // fission site banked in this collision, into the outgoing-energy bin of THAT site
        if (analog && scores[s] == OMCGPU_SCORE_NU_FISSION && eo_pos >= 0 && !skip_all_eout &&
            (M->survival_biasing || p.fission)) {
          if (!(M->run_CE ? p.xs_fission == 0 : false)) { // (CE: `if (p.macro_xs().fission == 0) continue;` comes first)
            const int fe = filt[eo_pos];
            const double* eb = M->filter_ddata + M->filter_data_ofs[fe];
            const int neb = M->filter_n_bins[fe];
            const int n_bank = p.fission ? (int)p.nu_bank.size() : 0; // p.n_bank(): reset after every collision
            for (int b = 0; b < n_bank; ++b) {
              const omcgpu_site& site = p.nu_bank[b];
              const double sc_site = RS.keff * site.wgt;
              int bin_out;
              if (!M->run_CE && M->filter_matches_groups[fe]) {
                bin_out = neb - static_cast<int>(site.E) - 1;
              } else {
                const double E_out = M->run_CE ? site.E : M->mg_energy_bin_avg[static_cast<int>(site.E)];
                if (E_out < eb[0] || E_out > eb[neb])
                  continue;
                bin_out = lower_bound_index(eb, eb + neb + 1, E_out);
              }
              int64_t idx2 = 0;
              double w2 = 1.;
              for (int i = 0; i < nf; ++i) {
                const Match& mt = matches[filt[i]];
                idx2 += (int64_t)(i == eo_pos ? bin_out : mt.bins[ibin[i]]) * strides[i];
                w2 *= mt.weights[ibin[i]];
              }
              results[idx2 * ns + s] += sc_site * w2;
            }
          }
          continue;
        }
        if (!skip)
          results[index * ns + s] += score * weight;
      }
      // advance the cartesian product (last filter fastest)
      bool done = true;
      for (int i = nf - 1; i >= 0; --i) {
        if (ibin[i] < (int)matches[filt[i]].bins.size() - 1) {
          ++ibin[i];
          done = false;
          break;
        } else {
          ibin[i] = 0;
        }
      }
      if (done)
        break;
    }
  }
}

// =============================================================================
// Particle events -- src/particle.cpp:114-158 (from_source), 160-230, 232-286,
// 288-329, 331-415, 417-473, 475-512, 549-760 ; src/simulation.cpp:573-649
// =============================================================================
void from_source(Particle& p, const omcgpu_site* src)
{
  for (int i = 0; i < OMCGPU_MAX_COORD; ++i)
    p.coord[i].reset();
  p.n_coord = 1;
  p.surface = SURFACE_NONE;
  p.cell_born = C_NONE;
  p.material = C_NONE;
  p.n_collision = 0;
  p.fission = false;
  p.wgt = src->wgt;
  p.wgt_last = src->wgt;
  p.r() = {src->r[0], src->r[1], src->r[2]};
  p.u() = {src->u[0], src->u[1], src->u[2]};
  p.r_last = p.r();
  p.u_last = p.u();
  if (M->run_CE) {
    p.E = src->E;
    p.g = 0;
  } else {
    p.g = static_cast<int>(src->E);
    p.g_last = static_cast<int>(src->E);
    p.E = M->mg_energy_bin_avg[p.g];
  }
  p.E_last = p.E;
  p.time = src->time;
  p.delayed_group = src->delayed_group;
}

void initialize_history(Particle& p, const omcgpu_site* site, int64_t index_source,
  const omcgpu_gen_in& in)
{
  p = Particle();
  from_source(p, site);
  p.id = in.id_offset + index_source;
  p.n_progeny = 0;
  p.n_event = 0;
  p.wgt_born = p.wgt;
  int64_t particle_seed = in.particle_seed_base + p.id;
  for (int i = 0; i < 4; i++)
    p.seeds[i] = future_seed(static_cast<uint64_t>(particle_seed) * in.stride, in.master_seed + i);
  RS.total_weight += p.wgt;
}

void event_calculate_xs(Particle& p)
{
  p.stream = STREAM_TRACKING;
  p.wgt_last = p.wgt;
  p.E_last = p.E;
  p.u_last = p.u();
  p.r_last = p.r();
  if (p.lowest_coord().cell == C_NONE) {
    if (!exhaustive_find_cell(p)) {
      mark_as_lost(p, "Could not find the cell containing particle");
      return;
    }
    if (p.cell_born == C_NONE)
      p.cell_born = p.lowest_coord().cell;
    for (int j = 0; j < p.n_coord; ++j)
      p.cell_last[j] = p.coord[j].cell;
    p.n_coord_last = p.n_coord;
  }
  RS.n_xs++;
  if (p.material != MATERIAL_VOID) {
    if (M->run_CE) {
      if (p.material != p.material_last || p.sqrtkT != p.sqrtkT_last || p.density_mult != p.density_mult_last)
        ce_calculate_xs(p);
    } else {
      mg_calculate_xs(p);
      p.g_last = p.g;
    }
  } else {
    p.xs_total = p.xs_absorption = p.xs_fission = p.xs_nu_fission = 0.0;
  }
}

double particle_speed(Particle& p)
{ // src/particle.cpp:46-73
  if (M->run_CE) {
    double mass = MASS_NEUTRON_EV;
    return C_LIGHT * std::sqrt(p.E * (p.E + 2 * mass)) / (p.E + mass);
  }
  if (p.material < 0)
    return INFINITY;
  int ba = mg_ba(p.material, p.mg_t, mg_angle_index(p.material, p.u()));
  return 1.0 / M->mg_inverse_velocity[ba * M->mg_n_groups + p.g];
}

void event_advance(Particle& p)
{
  p.boundary = distance_to_boundary(p);
  if (p.xs_total == 0.0)
    p.collision_distance = INFINITY;
  else
    p.collision_distance = -std::log(prn(p.current_seed())) / p.xs_total;
  double speed = particle_speed(p);
  double distance = std::min(p.boundary.distance, p.collision_distance);
  for (int j = 0; j < p.n_coord; ++j)
    p.coord[j].r = p.coord[j].r + distance * p.coord[j].u;
  double dt = distance / speed;
  p.time += dt;
  if (RS.tallies_active && M->n_tallies > 0)
    score_tracklength_tally(p, distance);
  p.k_trk += p.wgt * distance * p.xs_nu_fission;
}

void cross_surface(Particle& p)
{
  int s = std::abs(p.surface) - 1;
  int bc = M->surf_bc[s];
  if (bc == OMCGPU_BC_VACUUM) { // cross_vacuum_bc :631-655
    p.k_leak += p.wgt;
    p.wgt = 0.0;
    return;
  }
  if (bc == OMCGPU_BC_PERIODIC) { // TranslationalPeriodicBC::handle_particle + cross_periodic_bc :712-760
    if (p.n_coord != 1) {
      mark_as_lost(p, "Cannot transfer particle across surface in a lower universe.");
      return;
    }
    const double* t = M->surf_periodic_shift + 3 * s;
    const int axis = M->surf_periodic_axis[s];
    if (axis < 0) {
      p.r().x += t[0];
      p.r().y += t[1];
      p.r().z += t[2];
    } else { // RotationalPeriodicBC::handle_particle, src/boundary_condition.cpp:242-287
      const double c = t[0], sn = t[1];
      double r[3] = {p.r().x, p.r().y, p.r().z}, u[3] = {p.u().x, p.u().y, p.u().z};
      const int a1 = axis == 0 ? 1 : 0, a2 = axis == 2 ? 1 : 2;
      const double r1 = c * r[a1] - sn * r[a2], r2 = sn * r[a1] + c * r[a2];
      const double u1 = c * u[a1] - sn * u[a2], u2 = sn * u[a1] + c * u[a2];
      r[a1] = r1; r[a2] = r2; u[a1] = u1; u[a2] = u2;
      p.r() = {r[0], r[1], r[2]};
      p.u() = {u[0], u[1], u[2]};
    }
    const int partner = M->surf_periodic[s] + 1;
    // translational: same side of the partner (:137-143); rotational: the other side (:252-258) -- the
    // normals of a rotational pair both point into the geometry
    p.surface = (p.surface > 0) == (axis < 0) ? partner : -partner;
    p.n_coord = 1;
    if (!neighbor_list_find_cell(p))
      mark_as_lost(p, "Couldn't find particle after hitting periodic boundary");
    return;
  }
  if (bc == OMCGPU_BC_REFLECTIVE || bc == OMCGPU_BC_WHITE) { // ReflectiveBC / WhiteBC + cross_reflective_bc :657-710
    Vec3 n = surf_normal(s, p.r());
    Vec3 u = p.u();
    if (bc == OMCGPU_BC_WHITE) { // Surface::diffuse_reflect, src/surface.cpp:142-161
      n = n / norm(n);
      const double projection = dot(n, u);
      const double mu = (projection >= 0.0) ? -std::sqrt(prn(p.current_seed())) : std::sqrt(prn(p.current_seed()));
      u = rotate_angle(n, mu, nullptr, p.current_seed());
      u = u / norm(u);
    } else {
      const double projection = dot(n, u);
      const double magnitude = dot(n, n);
      n = n * (2.0 * projection / magnitude);
      u = u - n;
    }
    u = u / norm(u);
    if (p.n_coord != 1) {
      mark_as_lost(p, "Cannot reflect particle off surface in a lower universe.");
      return;
    }
    p.u() = u;
    p.coord[0].cell = p.cell_last[0];
    p.surface = -p.surface;
    p.n_coord = 1;
    if (!neighbor_list_find_cell(p)) {
      mark_as_lost(p, "Couldn't find particle after reflecting from surface");
      return;
    }
    return;
  }
  if (neighbor_list_find_cell(p))
    return;
  p.n_coord = 1;
  bool found = exhaustive_find_cell(p);
  if (!found) {
    p.surface = SURFACE_NONE;
    p.n_coord = 1;
    p.r() = p.r() + TINY_BIT * p.u();
    if (!exhaustive_find_cell(p)) {
      mark_as_lost(p, "could not be located in any cell and it did not leak");
      return;
    }
  }
}

void event_cross_surface(Particle& p)
{
  for (int j = 0; j < p.n_coord; ++j)
    p.cell_last[j] = p.coord[j].cell;
  p.n_coord_last = p.n_coord;
  p.surface = p.boundary.surface;
  p.n_coord = p.boundary.coord_level;
  if (p.boundary.lat_trans[0] != 0 || p.boundary.lat_trans[1] != 0 || p.boundary.lat_trans[2] != 0) {
    cross_lattice(p, p.boundary);
  } else {
    cross_surface(p);
  }
  RS.n_cross++;
}

void event_collide(Particle& p)
{
  p.k_col += p.wgt * p.xs_nu_fission / p.xs_total;
  p.surface = SURFACE_NONE;
  p.u_last_local = p.u_local();
  if (M->run_CE) {
    collision_ce(p);
    if (p.E < M->energy_cutoff) // src/physics.cpp:67-71
      p.wgt = 0.0;
  } else {
    collision_mg(p);
  }
  RS.n_coll++;
  if (RS.tallies_active && RS.n_collision_tallies > 0) // src/particle.cpp:362-363
    score_collision_tally(p);
  if (RS.tallies_active && RS.n_analog_tallies > 0) // :364-370
    score_analog_tally(p);
  p.wgt_bank = 0.0; // :379-385
  p.fission = false;
  p.material_last = C_NONE;
  for (int j = 0; j < p.n_coord - 1; ++j) { // src/particle.cpp:396-406
    if (M->cell_rotated[p.coord[j].cell])
      p.coord[j + 1].u = cell_rotate(p.coord[j].cell, p.coord[j].u);
    else
      p.coord[j + 1].u = p.coord[j].u;
  }
}

void event_revive(Particle& p)
{ // Particle::event_revive_from_secondary, src/particle.cpp:417-473
  ++p.n_event;
  RS.n_events++;
  if (p.n_event == M->max_particle_events)
    p.wgt = 0.0;
  if (!p.alive() && !p.secondary_bank.empty()) {
    const omcgpu_site site = p.secondary_bank.back();
    p.secondary_bank.pop_back();
    from_source(p, &site); // the random-number streams, the id and the k accumulators carry on
    p.n_event = 0;
  }
}

void event_death(Particle& p)
{
  RS.global_tally[OMCGPU_K_ABSORPTION] += p.k_abs;
  RS.global_tally[OMCGPU_K_COLLISION] += p.k_col;
  RS.global_tally[OMCGPU_K_TRACKLENGTH] += p.k_trk;
  RS.global_tally[OMCGPU_LEAKAGE] += p.k_leak;
  int64_t offset = p.id - 1 - RS.work_index;
  RS.progeny_per_particle[offset] = p.n_progeny;
}

void transport_history_based_single_particle(Particle& p)
{ // src/simulation.cpp:804-821
  while (p.alive()) {
    event_calculate_xs(p);
    if (p.alive())
      event_advance(p);
    if (p.alive()) {
      if (p.collision_distance > p.boundary.distance)
        event_cross_surface(p);
      else
        event_collide(p);
    }
    event_revive(p);
  }
  event_death(p);
}

// src/bank.cpp:76-134
void sort_fission_bank()
{
  auto& ppp = RS.progeny_per_particle;
  if (ppp.empty())
    return;
  int64_t run = 0;
  for (auto& v : ppp) {
    int64_t t = v;
    v = run;
    run += t;
  }
  std::vector<omcgpu_site> sorted(RS.fission_bank.size());
  for (size_t i = 0; i < RS.fission_bank.size(); i++) {
    const auto& site = RS.fission_bank[i];
    int64_t offset = site.parent_id - 1 - RS.work_index;
    int64_t idx = ppp[offset] + site.progeny_id;
    sorted[idx] = site;
  }
  RS.fission_bank.swap(sorted);
}

// one generation on `n` source sites (transport + sort) -- same contract as
// omcgpu_transport_generation
void transport_generation(const omcgpu_gen_in& in, const omcgpu_site* source_bank,
  int64_t fission_capacity)
{
  RS.keff = in.keff;
  RS.work_index = in.id_offset;
  RS.tallies_active = in.tallies_active != 0;
  RS.n_collision_tallies = 0;
  RS.n_analog_tallies = 0;
  for (int t = 0; t < M->n_tallies; ++t) {
    RS.n_collision_tallies += M->tally_estimator[t] == OMCGPU_ESTIMATOR_COLLISION ? 1 : 0;
    RS.n_analog_tallies += M->tally_estimator[t] == OMCGPU_ESTIMATOR_ANALOG ? 1 : 0;
  }
  RS.fission_bank.clear();
  RS.fission_capacity = fission_capacity;
  RS.progeny_per_particle.assign(in.n_particles, 0);
  for (int k = 0; k < 4; ++k)
    RS.global_tally[k] = 0.0;
  RS.total_weight = 0.0;
  RS.n_lost = RS.n_events = RS.n_xs = RS.n_coll = RS.n_cross = 0;
  if (M->n_tallies > 0 && (int64_t)RS.tally_values.size() != M->tally_result_ofs[M->n_tallies])
    RS.tally_values.assign(M->tally_result_ofs[M->n_tallies], 0.0);
  Particle p;
  for (int64_t i = 1; i <= in.n_particles; ++i) {
    initialize_history(p, &source_bank[i - 1], i, in);
    transport_history_based_single_particle(p);
  }
  sort_fission_bank();
}

// src/eigenvalue.cpp:152-178 (single rank: start = 0)
int64_t comb_indices(int64_t start, int64_t n_local, int64_t total, int64_t n_particles,
  int64_t seed_id, int64_t master_seed, uint64_t stride, std::vector<int64_t>& idx)
{
  uint64_t seed = init_seed(seed_id, STREAM_TRACKING, master_seed, stride);
  double teeth_distance = static_cast<double>(total) / n_particles;
  double teeth_offset = prn(&seed) * teeth_distance;
  int64_t end = start + n_local;
  int64_t tooth_start = std::ceil((start - teeth_offset) / teeth_distance);
  int64_t tooth_end = std::floor((end - teeth_offset) / teeth_distance) + 1;
  double tooth = tooth_start * teeth_distance + teeth_offset;
  idx.clear();
  for (int64_t i = tooth_start; i < tooth_end; i++) {
    idx.push_back((int64_t)std::floor(tooth) - start);
    tooth += teeth_distance;
  }
  return (int64_t)idx.size();
}

} // namespace

// =============================================================================
// C interface (ctypes)
// =============================================================================
extern "C" {

const char* orc_last_error()
{
  return errbuf;
}

int orc_load(const omcgpu_model* model)
{
  M = model;
  RS = RunState();
  return 0;
}

int orc_future_seed(const uint64_t* n, const uint64_t* seed, int64_t count, uint64_t* out)
{
  for (int64_t i = 0; i < count; ++i)
    out[i] = future_seed(n[i], seed[i]);
  return 0;
}

int orc_prn(const uint64_t* seed, int64_t n_seeds, int32_t count, double* out)
{
  for (int64_t i = 0; i < n_seeds; ++i) {
    uint64_t s = seed[i];
    for (int j = 0; j < count; ++j)
      out[i * count + j] = prn(&s);
  }
  return 0;
}

double orc_evaluate_legendre(int n, const double* data, double x)
{
  return evaluate_legendre(n, data, x);
}

void orc_rotate_angle(double* uvw, double mu, const double* phi, uint64_t* seed)
{
  Vec3 u = rotate_angle({uvw[0], uvw[1], uvw[2]}, mu, phi, seed);
  uvw[0] = u.x;
  uvw[1] = u.y;
  uvw[2] = u.z;
}

double orc_maxwell_spectrum(double T, uint64_t* seed)
{
  return maxwell_spectrum(T, seed);
}
double orc_watt_spectrum(double a, double b, uint64_t* seed)
{
  return watt_spectrum(a, b, seed);
}
int64_t orc_lower_bound_index(const double* a, int64_t n, double v)
{
  return lower_bound_index(a, a + n, v);
}

int orc_find_cell(const double* xyz, const double* uvw, int64_t n, int32_t* cell, int32_t* instance,
  int32_t* material, int32_t* n_coord)
{
  for (int64_t i = 0; i < n; ++i) {
    Particle p;
    p.r() = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    p.u() = uvw ? Vec3 {uvw[3 * i], uvw[3 * i + 1], uvw[3 * i + 2]} : Vec3 {0.0, 0.0, 1.0};
    bool found = exhaustive_find_cell(p);
    cell[i] = found ? p.lowest_coord().cell : -1;
    instance[i] = found ? p.cell_instance : -1;
    material[i] = found ? p.material : -2;
    n_coord[i] = found ? p.n_coord : 0;
  }
  return 0;
}

int orc_distance_to_boundary(const double* xyz, const double* uvw, int64_t n, double* distance,
  int32_t* surface, int32_t* level, int32_t* lattice_trans)
{
  for (int64_t i = 0; i < n; ++i) {
    Particle p;
    p.r() = {xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
    p.u() = {uvw[3 * i], uvw[3 * i + 1], uvw[3 * i + 2]};
    if (!exhaustive_find_cell(p)) {
      distance[i] = -1.0;
      surface[i] = 0;
      level[i] = 0;
      lattice_trans[3 * i] = lattice_trans[3 * i + 1] = lattice_trans[3 * i + 2] = 0;
      continue;
    }
    BoundaryInfo b = distance_to_boundary(p);
    distance[i] = b.distance;
    surface[i] = b.surface;
    level[i] = b.coord_level;
    for (int k = 0; k < 3; ++k)
      lattice_trans[3 * i + k] = b.lat_trans[k];
  }
  return 0;
}

// the reflected direction of ReflectiveBC (the arithmetic of cross_surface's reflective branch)
int orc_reflect(int64_t n, const int32_t* surface, const double* xyz, const double* uvw, double* new_uvw)
{
  for (int64_t k = 0; k < n; ++k) {
    const int s = std::abs(surface[k]) - 1;
    Vec3 r {xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2]};
    Vec3 u {uvw[3 * k], uvw[3 * k + 1], uvw[3 * k + 2]};
    Vec3 nrm = surf_normal(s, r);
    const double projection = dot(nrm, u);
    const double magnitude = dot(nrm, nrm);
    nrm = nrm * (2.0 * projection / magnitude);
    u = u - nrm;
    u = u / norm(u);
    new_uvw[3 * k] = u.x;
    new_uvw[3 * k + 1] = u.y;
    new_uvw[3 * k + 2] = u.z;
  }
  return 0;
}

// Straight-line tracking with the crossing logic of event_cross_surface (transmission
// surfaces and lattice boundaries; a track ends at a surface with a boundary condition).
// per step: cell id (-1 ended at a BC, -2 lost), material, depth, flight distance.
int orc_geom_track(int64_t n, const double* xyz, const double* uvw, int steps, int32_t* cell_id,
  int32_t* material, int32_t* n_coord, double* distance)
{
  for (int64_t k = 0; k < n; ++k) {
    for (int s = 0; s < steps; ++s) {
      cell_id[k * steps + s] = -1;
      material[k * steps + s] = -1;
      n_coord[k * steps + s] = 0;
      distance[k * steps + s] = 0.0;
    }
    Particle p;
    p.r() = {xyz[3 * k], xyz[3 * k + 1], xyz[3 * k + 2]};
    p.u() = {uvw[3 * k], uvw[3 * k + 1], uvw[3 * k + 2]};
    if (!exhaustive_find_cell(p))
      continue;
    for (int s = 0; s < steps; ++s) {
      p.boundary = distance_to_boundary(p);
      if (!(p.boundary.distance < 1e30))
        break;
      for (int j = 0; j < p.n_coord; ++j)
        p.coord[j].r = p.coord[j].r + p.boundary.distance * p.coord[j].u;
      const bool lattice = p.boundary.lat_trans[0] != 0 || p.boundary.lat_trans[1] != 0 || p.boundary.lat_trans[2] != 0;
      if (!lattice && M->surf_bc[std::abs(p.boundary.surface) - 1] != OMCGPU_BC_NONE)
        break; // ended at a boundary condition
      event_cross_surface(p);
      distance[k * steps + s] = p.boundary.distance;
      if (p.lost) {
        cell_id[k * steps + s] = -2;
        break;
      }
      cell_id[k * steps + s] = M->cell_id[p.coord[p.n_coord - 1].cell];
      material[k * steps + s] = p.material;
      n_coord[k * steps + s] = p.n_coord;
    }
  }
  return 0;
}

// sample_cxs_target_velocity (src/physics.cpp:990-1041) for n (E, u, seed) triples
int orc_sample_cxs_target_velocity(double awr, double kT, int64_t n, const double* E, const double* uvw,
  const uint64_t* seed_in, double* v_target, uint64_t* seed_out)
{
  for (int64_t k = 0; k < n; ++k) {
    uint64_t seed = seed_in[k];
    const Vec3 v = sample_cxs_target_velocity(awr, E[k], Vec3 {uvw[3 * k], uvw[3 * k + 1], uvw[3 * k + 2]}, kT, &seed);
    v_target[3 * k] = v.x;
    v_target[3 * k + 1] = v.y;
    v_target[3 * k + 2] = v.z;
    seed_out[k] = seed;
  }
  return 0;
}

// elastic_scatter (src/physics.cpp:754-818) off nuclide `nuc` of the loaded model
int orc_elastic_scatter(int nuc, double kT, int64_t n, const double* E, const double* uvw,
  const uint64_t* seed_in, double* E_out, double* uvw_out, double* mu_out, uint64_t* seed_out)
{
  for (int64_t k = 0; k < n; ++k) {
    Particle p;
    p.E = E[k];
    p.u() = {uvw[3 * k], uvw[3 * k + 1], uvw[3 * k + 2]};
    p.stream = 0;
    p.seeds[0] = seed_in[k];
    ce_elastic_scatter(nuc, kT, p);
    E_out[k] = p.E;
    uvw_out[3 * k] = p.u().x;
    uvw_out[3 * k + 1] = p.u().y;
    uvw_out[3 * k + 2] = p.u().z;
    mu_out[k] = p.mu;
    seed_out[k] = p.seeds[0];
  }
  return 0;
}

// collision (src/physics.cpp:43-168) of n fresh neutrons of weight wgt0 in the loaded CE model; fission sites of
// collision k are site_*[site_ofs[k] .. site_ofs[k+1])
int64_t orc_collision(int64_t n, const int32_t* material, const double* E, const double* uvw,
  const uint64_t* seed_in, double k_eff, double wgt0, double* E_out, double* uvw_out, double* wgt_out,
  double* kabs_out, uint64_t* seed_out, int64_t* site_ofs, double* site_E, double* site_u, int32_t* site_group,
  double* site_time, int64_t site_cap)
{
  RS.keff = k_eff;
  RS.fission_capacity = 64;
  int64_t total = 0;
  for (int64_t k = 0; k < n; ++k) {
    Particle p;
    p.E = E[k];
    p.wgt = wgt0;
    p.u() = {uvw[3 * k], uvw[3 * k + 1], uvw[3 * k + 2]};
    p.stream = 0;
    p.seeds[0] = seed_in[k];
    p.material = material[k];
    RS.fission_bank.clear();
    ce_calculate_xs(p);
    collision_ce(p);
    E_out[k] = p.E;
    uvw_out[3 * k] = p.u().x;
    uvw_out[3 * k + 1] = p.u().y;
    uvw_out[3 * k + 2] = p.u().z;
    wgt_out[k] = p.wgt;
    kabs_out[k] = p.k_abs;
    seed_out[k] = p.seeds[0];
    site_ofs[k] = total;
    if (total + (int64_t)RS.fission_bank.size() > site_cap)
      return -1;
    for (const auto& s : RS.fission_bank) {
      site_E[total] = s.E;
      site_u[3 * total] = s.u[0];
      site_u[3 * total + 1] = s.u[1];
      site_u[3 * total + 2] = s.u[2];
      site_group[total] = s.delayed_group;
      site_time[total] = s.time;
      ++total;
    }
  }
  site_ofs[n] = total;
  return total;
}

// inelastic_scatter (src/physics.cpp:1119-1167) through reaction `rx` (global index) of nuclide `nuc`
int orc_inelastic_scatter(int nuc, int rx, int64_t n, const double* E, const double* uvw,
  const uint64_t* seed_in, double* E_out, double* uvw_out, double* mu_out, double* wgt_out, uint64_t* seed_out)
{
  for (int64_t k = 0; k < n; ++k) {
    Particle p;
    p.E = E[k];
    p.wgt = 1.0;
    p.u() = {uvw[3 * k], uvw[3 * k + 1], uvw[3 * k + 2]};
    p.stream = 0;
    p.seeds[0] = seed_in[k];
    ce_inelastic_scatter(nuc, rx, p);
    E_out[k] = p.E;
    uvw_out[3 * k] = p.u().x;
    uvw_out[3 * k + 1] = p.u().y;
    uvw_out[3 * k + 2] = p.u().z;
    mu_out[k] = p.mu;
    wgt_out[k] = p.wgt;
    seed_out[k] = p.seeds[0];
  }
  return 0;
}

// sample_fission_neutron (src/physics.cpp:1043-1117) of nuclide `nuc`, fission reaction `rx` (global index)
int orc_fission_neutron(int nuc, int rx, int64_t n, const double* E, const double* uvw,
  const uint64_t* seed_in, double* E_out, double* uvw_out, int32_t* group_out, double* time_out,
  uint64_t* seed_out)
{
  for (int64_t k = 0; k < n; ++k) {
    Particle p;
    p.E = E[k];
    p.u() = {uvw[3 * k], uvw[3 * k + 1], uvw[3 * k + 2]};
    p.stream = 0;
    p.seeds[0] = seed_in[k];
    omcgpu_site site;
    std::memset(&site, 0, sizeof(site));
    ce_sample_fission_neutron(nuc, rx, &site, p);
    E_out[k] = site.E;
    uvw_out[3 * k] = site.u[0];
    uvw_out[3 * k + 1] = site.u[1];
    uvw_out[3 * k + 2] = site.u[2];
    group_out[k] = site.delayed_group;
    time_out[k] = site.time;
    seed_out[k] = p.seeds[0];
  }
  return 0;
}

// ReactionProduct::sample of product `prod` of the loaded model (isotropic angle + energy law)
int orc_sample_product(int prod, int64_t n, const double* E_in, const uint64_t* seed_in, double* E_out,
  double* mu_out, uint64_t* seed_out)
{
  for (int64_t k = 0; k < n; ++k) {
    uint64_t seed = seed_in[k];
    ce_sample_product(prod, E_in[k], E_out[k], mu_out[k], &seed);
    seed_out[k] = seed;
  }
  return 0;
}

// Function1D evaluation of function `fn` of the loaded model at n abscissae
int orc_eval_function(int fn, int64_t n, const double* x, double* y)
{
  for (int64_t k = 0; k < n; ++k)
    y[k] = eval_function(fn, x[k]);
  return 0;
}

// StructuredMesh::bins_crossed (src/mesh.cpp:999-1146) for n tracks r0 -> r1: track k's
// (bin, length fraction) pairs are out[ofs[k] .. ofs[k+1]); returns the number of pairs
int64_t orc_mesh_bins_crossed(int mesh, int64_t n, const double* r0, const double* r1, const double* u,
  int64_t* ofs, int32_t* bins, double* lengths, int64_t cap)
{
  Match match;
  int64_t total = 0;
  for (int64_t k = 0; k < n; ++k) {
    match.bins.clear();
    match.weights.clear();
    mesh_bins_crossed(mesh, Vec3 {r0[3 * k], r0[3 * k + 1], r0[3 * k + 2]},
      Vec3 {r1[3 * k], r1[3 * k + 1], r1[3 * k + 2]}, Vec3 {u[3 * k], u[3 * k + 1], u[3 * k + 2]}, match);
    ofs[k] = total;
    if (total + (int64_t)match.bins.size() > cap)
      return -1;
    for (size_t j = 0; j < match.bins.size(); ++j) {
      bins[total + j] = match.bins[j];
      lengths[total + j] = match.weights[j];
    }
    total += match.bins.size();
  }
  ofs[n] = total;
  return total;
}

int orc_calculate_xs(const double* E, const int32_t* material, int64_t n, double* out,
  int32_t* i_grid, int32_t max_nuc)
{
  for (int64_t i = 0; i < n; ++i) {
    Particle p;
    p.material = material[i];
    if (M->run_CE) {
      p.E = E[i];
      ce_calculate_xs(p, i_grid ? i_grid + i * max_nuc : nullptr, max_nuc);
    } else {
      p.g = (int)E[i];
      p.sqrtkT = M->cell_sqrtkT[0];
      p.coord[0].u = {0, 0, 1};
      mg_calculate_xs(p);
    }
    out[4 * i + 0] = p.xs_total;
    out[4 * i + 1] = p.xs_absorption;
    out[4 * i + 2] = p.xs_fission;
    out[4 * i + 3] = p.xs_nu_fission;
  }
  return 0;
}

int orc_transport_generation(const omcgpu_gen_in* in, const omcgpu_site* source_bank,
  omcgpu_site* fission_bank, int64_t fission_capacity, omcgpu_gen_out* out)
{
  transport_generation(*in, source_bank, fission_capacity);
  std::memset(out, 0, sizeof(*out));
  out->n_fission = RS.fission_bank.size();
  out->n_lost = RS.n_lost;
  out->n_events = RS.n_events;
  out->n_xs_lookups = RS.n_xs;
  out->n_collisions = RS.n_coll;
  out->n_crossings = RS.n_cross;
  for (int k = 0; k < 4; ++k)
    out->global_tally[k] = RS.global_tally[k];
  out->total_weight = RS.total_weight;
  if (fission_bank)
    std::copy(RS.fission_bank.begin(), RS.fission_bank.end(), fission_bank);
  return 0;
}

int orc_tally_results(int32_t t, double* value, int64_t n)
{
  if (RS.tally_values.empty())
    return 0;
  double* v = RS.tally_values.data() + M->tally_result_ofs[t];
  for (int64_t i = 0; i < n; ++i) {
    value[i] += v[i];
    v[i] = 0.0;
  }
  return 0;
}

int64_t orc_comb_indices(int64_t start, int64_t n_local, int64_t total, int64_t n_particles,
  int64_t seed_id, int64_t master_seed, uint64_t stride, int64_t* idx, int64_t cap)
{
  std::vector<int64_t> v;
  int64_t n = comb_indices(start, n_local, total, n_particles, seed_id, master_seed, stride, v);
  for (int64_t i = 0; i < n && i < cap; ++i)
    idx[i] = v[i];
  return n;
}

// ---- initial source: src/source.cpp:344-399 (IndependentSource::sample with a
// box + isotropic + Watt(0.988e6, 2.249e-6) + time 0), :601-625, :627-660 ------
int orc_sample_initial_source(int64_t n, int64_t id_offset, int64_t total_gen_particles,
  int64_t master_seed, uint64_t stride, const double* box, int only_fissionable, double watt_a,
  double watt_b, omcgpu_site* bank)
{
  for (int64_t i = 0; i < n; ++i) {
    int64_t id = total_gen_particles + id_offset + i + 1;
    uint64_t seed = init_seed(id, STREAM_SOURCE, master_seed, stride);
    omcgpu_site site;
    std::memset(&site, 0, sizeof(site));
    site.wgt = 1.0;
    site.parent_nuclide = -1;
    bool accepted = false;
    int64_t tries = 0;
    while (!accepted) {
      double xi0 = prn(&seed), xi1 = prn(&seed), xi2 = prn(&seed);
      site.r[0] = box[0] + xi0 * (box[3] - box[0]);
      site.r[1] = box[1] + xi1 * (box[4] - box[1]);
      site.r[2] = box[2] + xi2 * (box[5] - box[2]);
      // satisfies_spatial_constraints, src/source.cpp:226-275
      Particle p;
      p.r() = {site.r[0], site.r[1], site.r[2]};
      p.u() = {0.0, 0.0, 1.0};
      accepted = exhaustive_find_cell(p);
      if (accepted && only_fissionable) {
        accepted = (p.material != MATERIAL_VOID) && M->mat_fissionable[p.material];
      }
      if (++tries > 10000000) {
        std::snprintf(errbuf, sizeof(errbuf), "source rejection limit reached");
        return -1;
      }
    }
    Vec3 u = isotropic_direction(&seed);
    site.u[0] = u.x;
    site.u[1] = u.y;
    site.u[2] = u.z;
    while (true) {
      site.E = watt_spectrum(watt_a, watt_b, &seed);
      if (site.E < M->energy_max)
        break;
    }
    site.time = 0.0; // Discrete with one point consumes no random number
    if (!M->run_CE) {
      // src/source.cpp:651-656 ; rev_energy_bins_ is ascending
      int G = M->mg_n_groups;
      std::vector<double> rev(M->mg_energy_bins, M->mg_energy_bins + G + 1);
      std::reverse(rev.begin(), rev.end());
      double e = lower_bound_index(rev.data(), rev.data() + G + 1, site.E);
      site.E = G - e - 1.;
    }
    bank[i] = site;
  }
  return 0;
}

// ---- eigenvalue driver: src/simulation.cpp:232-280,359-571 ;
// src/eigenvalue.cpp:50-84,86-365,367-418,420-577 ; src/tallies/tally.cpp:821-878,
// 1059-1104 -------------------------------------------------------------------
struct orc_settings {
  int64_t n_particles;
  int32_t n_batches;
  int32_t n_inactive;
  int32_t gen_per_batch;
  int32_t reserved;
  int64_t master_seed;
  uint64_t stride;
};

struct orc_results {
  double k_combined[2];
  double keff, keff_std;
  double global_sum[4];
  double global_sum_sq[4];
  int64_t n_realizations;
  int64_t n_lost;
  int64_t n_events;
  int64_t n_xs_lookups;
  double seconds_active;
};

int orc_get_keff(int64_t n, const double* gsum, const double* gsq, double k_col_abs,
  double k_col_tra, double k_abs_tra, double keff, double keff_std, double* k_combined)
{ // src/eigenvalue.cpp:420-577
  k_combined[0] = 0.0;
  k_combined[1] = 0.0;
  if (n <= 3) {
    k_combined[0] = keff;
    k_combined[1] = keff_std;
    if (n <= 1)
      k_combined[1] = std::numeric_limits<double>::infinity();
    return 0;
  }
  double kv[3];
  double cov[3][3];
  kv[0] = gsum[OMCGPU_K_COLLISION] / n;
  kv[1] = gsum[OMCGPU_K_ABSORPTION] / n;
  kv[2] = gsum[OMCGPU_K_TRACKLENGTH] / n;
  cov[0][0] = (gsq[OMCGPU_K_COLLISION] - n * kv[0] * kv[0]) / (n - 1);
  cov[1][1] = (gsq[OMCGPU_K_ABSORPTION] - n * kv[1] * kv[1]) / (n - 1);
  cov[2][2] = (gsq[OMCGPU_K_TRACKLENGTH] - n * kv[2] * kv[2]) / (n - 1);
  cov[0][1] = (k_col_abs - n * kv[0] * kv[1]) / (n - 1);
  cov[0][2] = (k_col_tra - n * kv[0] * kv[2]) / (n - 1);
  cov[1][2] = (k_abs_tra - n * kv[1] * kv[2]) / (n - 1);
  cov[1][0] = cov[0][1];
  cov[2][0] = cov[0][2];
  cov[2][1] = cov[1][2];
  int i, j;
  bool use_three = false;
  if ((std::abs(kv[0] - kv[1]) / kv[0] < FP_REL_PRECISION) &&
      (std::abs(cov[0][0] - cov[1][1]) / cov[0][0] < FP_REL_PRECISION)) {
    i = 0;
    j = 2;
  } else if ((std::abs(kv[0] - kv[2]) / kv[0] < FP_REL_PRECISION) &&
             (std::abs(cov[0][0] - cov[2][2]) / cov[0][0] < FP_REL_PRECISION)) {
    i = 0;
    j = 1;
  } else if ((std::abs(kv[1] - kv[2]) / kv[1] < FP_REL_PRECISION) &&
             (std::abs(cov[1][1] - cov[2][2]) / cov[1][1] < FP_REL_PRECISION)) {
    i = 0;
    j = 1;
  } else {
    use_three = true;
  }
  if (use_three) {
    double g = 0.0;
    double S[3] = {0, 0, 0};
    for (int l = 0; l < 3; ++l) {
      int k;
      switch (l) {
      case 0:
        i = 0;
        j = 1;
        k = 2;
        break;
      case 1:
        i = 1;
        j = 2;
        k = 0;
        break;
      default:
        i = 2;
        j = 0;
        k = 1;
        break;
      }
      double f = cov[j][j] * (cov[k][k] - cov[i][k]) - cov[k][k] * cov[i][j] +
                 cov[j][k] * (cov[i][j] + cov[i][k] - cov[j][k]);
      S[0] += f * cov[0][l];
      S[1] += (cov[j][j] + cov[k][k] - 2.0 * cov[j][k]) * kv[l] * kv[l];
      S[2] += (cov[k][k] + cov[i][j] - cov[j][k] - cov[i][k]) * kv[l] * kv[j];
      k_combined[0] += f * kv[l];
      g += f;
    }
    for (auto& S_i : S)
      S_i *= (n - 1);
    S[0] *= (n - 1) * (n - 1);
    k_combined[0] /= g;
    g *= (n - 1) * (n - 1);
    k_combined[1] = std::sqrt(S[0] / (g * n * (n - 3)) * (1 + n * ((S[1] - 2 * S[2]) / g)));
  } else {
    double f = kv[i] - kv[j];
    double g = cov[i][i] + cov[j][j] - 2.0 * cov[i][j];
    k_combined[0] = kv[i] - (cov[i][i] - cov[i][j]) / g * f;
    k_combined[1] =
      (cov[i][i] * cov[j][j] - cov[i][j] * cov[i][j]) * (g + n * f * f) / (n * (n - 2) * g * g);
    k_combined[1] = std::sqrt(k_combined[1]);
  }
  return 0;
}

// Full single-rank eigenvalue run.  source_bank[n_particles] holds the initial
// source on entry and the final source on exit.  tally_sum / tally_sum_sq are
// dense [sum of n_filter_bins*n_scores over tallies] accumulators; k_generation
// has n_batches*gen_per_batch entries.
int orc_run_eigenvalue(const orc_settings* st, omcgpu_site* source_bank, orc_results* res,
  double* k_generation, double* tally_sum, double* tally_sum_sq)
{
  const int64_t N = st->n_particles;
  double keff = 1.0, keff_std = 0.0;
  double k_sum[2] = {0, 0};
  double gt_value[4] = {0, 0, 0, 0}, gt_sum[4] = {0, 0, 0, 0}, gt_sq[4] = {0, 0, 0, 0};
  double k_col_abs = 0, k_col_tra = 0, k_abs_tra = 0;
  int64_t n_realizations = 0;
  std::vector<double> kgen;
  int total_gen = 0;
  int64_t n_tally = M->n_tallies > 0 ? M->tally_result_ofs[M->n_tallies] : 0;
  std::vector<omcgpu_site> fission_bank(3 * N);
  std::vector<int64_t> idx;
  RS = RunState();
  int64_t ev = 0, nxs = 0, lost = 0;
  for (int batch = 1; batch <= st->n_batches; ++batch) {
    double total_weight = 0.0;
    bool active = batch > st->n_inactive;
    for (int gen = 1; gen <= st->gen_per_batch; ++gen) {
      int overall_gen = st->gen_per_batch * (batch - 1) + gen;
      double keff_generation = gt_value[OMCGPU_K_TRACKLENGTH];
      omcgpu_gen_in in;
      in.n_particles = N;
      in.id_offset = 0;
      in.particle_seed_base = (int64_t)(total_gen + overall_gen - 1) * N;
      in.master_seed = st->master_seed;
      in.stride = st->stride;
      in.keff = keff;
      in.tallies_active = active ? 1 : 0;
      in.source_bank_clean = 0;
      transport_generation(in, source_bank, 3 * N);
      total_weight += RS.total_weight;
      ev += RS.n_events;
      nxs += RS.n_xs;
      lost += RS.n_lost;
      // finalize_generation, src/simulation.cpp:522-571
      gt_value[OMCGPU_K_COLLISION] += RS.global_tally[OMCGPU_K_COLLISION];
      gt_value[OMCGPU_K_ABSORPTION] += RS.global_tally[OMCGPU_K_ABSORPTION];
      gt_value[OMCGPU_K_TRACKLENGTH] += RS.global_tally[OMCGPU_K_TRACKLENGTH];
      gt_value[OMCGPU_LEAKAGE] += RS.global_tally[OMCGPU_LEAKAGE];
      // synchronize_bank, src/eigenvalue.cpp:86-365 (one rank)
      int64_t n_bank = RS.fission_bank.size();
      if (n_bank == 0) {
        std::snprintf(errbuf, sizeof(errbuf), "No fission sites banked");
        return -1;
      }
      comb_indices(0, n_bank, n_bank, N, total_gen + overall_gen, st->master_seed, st->stride, idx);
      for (int64_t i = 0; i < N; ++i)
        source_bank[i] = RS.fission_bank[idx[i]];
      // calculate_generation_keff, src/eigenvalue.cpp:50-84
      keff_generation = gt_value[OMCGPU_K_TRACKLENGTH] - keff_generation;
      double keff_reduced = keff_generation / N;
      kgen.push_back(keff_reduced);
      // calculate_average_keff, src/eigenvalue.cpp:367-418
      int i = overall_gen - 1;
      int n = active ? st->gen_per_batch * (int)n_realizations + gen : 0;
      if (n <= 0) {
        keff = kgen[i];
      } else {
        k_sum[0] += kgen[i];
        k_sum[1] += std::pow(kgen[i], 2);
        keff = k_sum[0] / n;
        if (n > 1) {
          keff_std = 1.0 * std::sqrt((k_sum[1] / n - std::pow(keff, 2)) / (n - 1));
          if (!std::isfinite(keff_std))
            keff_std = 0.0;
        }
      }
    }
    // finalize_batch -> accumulate_tallies, src/tallies/tally.cpp:1059-1104
    n_realizations += 1;
    if (active) {
      double k_col = gt_value[OMCGPU_K_COLLISION] / total_weight;
      double k_abs = gt_value[OMCGPU_K_ABSORPTION] / total_weight;
      double k_tra = gt_value[OMCGPU_K_TRACKLENGTH] / total_weight;
      k_col_abs += k_col * k_abs;
      k_col_tra += k_col * k_tra;
      k_abs_tra += k_abs * k_tra;
    }
    for (int i = 0; i < 4; ++i) {
      double val = gt_value[i] / total_weight;
      gt_value[i] = 0.0;
      gt_sum[i] += val;
      gt_sq[i] += val * val;
    }
    if (active && n_tally > 0) { // Tally::accumulate, :821-878
      double norm = 1.0 / (static_cast<double>(N) * st->gen_per_batch);
      for (int64_t i = 0; i < n_tally; ++i) {
        double val = RS.tally_values[i] * norm;
        RS.tally_values[i] = 0.0;
        tally_sum[i] += val;
        tally_sum_sq[i] += val * val;
      }
    }
    if (batch <= st->n_inactive) { // src/simulation.cpp:416-419
      for (int i = 0; i < 4; ++i)
        gt_value[i] = gt_sum[i] = gt_sq[i] = 0.0;
      n_realizations = 0;
    }
  }
  std::copy(kgen.begin(), kgen.end(), k_generation);
  res->keff = keff;
  res->keff_std = keff_std;
  for (int i = 0; i < 4; ++i) {
    res->global_sum[i] = gt_sum[i];
    res->global_sum_sq[i] = gt_sq[i];
  }
  res->n_realizations = n_realizations;
  res->n_lost = lost;
  res->n_events = ev;
  res->n_xs_lookups = nxs;
  orc_get_keff(n_realizations, gt_sum, gt_sq, k_col_abs, k_col_tra, k_abs_tra, keff, keff_std,
    res->k_combined);
  return 0;
}

} // extern "C"
