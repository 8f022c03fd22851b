// device.h -- the per-particle event functions of the engine (host/device).
//
// Each function cites the reference lines whose behaviour it reproduces.  The
// code is organised for the GPU: particle state is read from / written to SoA
// columns (PState) one coordinate level at a time, so no per-thread arrays are
// indexed dynamically; the read-only model lives behind one constant-memory
// struct (`DM`) of device pointers.
#pragma once

#include "common.h"
#include <cstddef>
#include <cstring>

// ---- what the extended build (libomcgpu_ext.so, -DOMC_EXTENDED=1) carries and the default build does not:
// omcgpu_needs_extended_build (engine.cu) names the models that need it, omcgpu_init of the default build
// refuses them (every one of these features costs the event kernels registers / spills / instruction
// cache even when no particle ever takes the branch: profiles/r58, r66, r68-r70, r75)
#ifndef OMC_EXTENDED
#define OMC_EXTENDED 0 /* 1: libomcgpu_ext.so, the build with the rarely needed features compiled in */
#endif
#ifndef OMC_HEX_LATTICES
#define OMC_HEX_LATTICES OMC_EXTENDED
#endif
#ifndef OMC_SECONDARIES
#define OMC_SECONDARIES OMC_EXTENDED /* secondary neutrons of integer yields: create_secondary / ev_revive */
#endif
#ifndef OMC_RARE_SURFACES
#define OMC_RARE_SURFACES OMC_EXTENDED /* axis cones, general quadrics; white and rotational periodic boundaries */
#endif
#ifndef OMC_AFTER_COLLISION_TALLIES
#define OMC_AFTER_COLLISION_TALLIES OMC_EXTENDED /* analog estimator: collision records */
#endif

namespace omc {

#if OMC_CUDA
__constant__ omcgpu_model c_dm;
__constant__ PState c_ps; // the particle columns never move after omcgpu_init (out-of-line device functions read it)
#else
static omcgpu_model c_dm;
#endif
#define DM c_dm

// read-only loads go through the non-coherent path on the device
template<class T>
OMC_HD T ld(const T* p)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  return __ldg(p);
#else
  return *p;
#endif
}

// ---- CSG tables staged in shared memory ---------------------------------------------------------
// Surfaces, region tokens, cells, universes' cell lists, half-space candidate lists and lattices
// are uploaded into ONE block of device memory (the geometry block: a few KB for the configurations
// of BASELINE.json), and the geometry pointers of c_dm hold OFFSETS into it.  Kernels that walk the
// geometry are launched with the block's size as dynamic shared memory and copy it there once per
// thread block (stage_geometry); every read of a geometry table goes through gld(): base + offset,
// where the base is the thread block's shared-memory copy when the launch staged one, else the
// block itself in global memory (a geometry larger than the budget OMCGPU_GEO_SMEM, kernels
// launched without dynamic shared memory such as k_score).  One generic base, no per-access test.
// Measured (profiles/r52-r54, README): C2 unchanged, C3 / C4 1-3 % slower than plain __ldg loads,
// an LDS-only build (no global fallback, -DOMC_GEO_ALWAYS) 1-2 % faster: the geometry tables are
// L1 hits either way and not what bounds the event kernels.
#if OMC_CUDA
__constant__ const char* c_geo_blob;
#if defined(__CUDA_ARCH__)
extern __shared__ __align__(16) unsigned char g_geo_smem[];
OMC_HD uint32_t geo_staged_bytes()
{
  uint32_t r;
  asm("mov.u32 %0, %%dynamic_smem_size;" : "=r"(r));
  return r;
}
#endif
#endif
template<class T>
OMC_HD T gld(const T* p)
{
#if OMC_CUDA && defined(__CUDA_ARCH__) && defined(OMC_GEO_ALWAYS)
  // measurement build: shared memory only (omcgpu_init refuses geometries beyond the budget)
  return *reinterpret_cast<const T*>(g_geo_smem + (uint32_t)reinterpret_cast<uintptr_t>(p));
#elif OMC_CUDA && defined(__CUDA_ARCH__) && defined(OMC_GEO_NEVER)
  return __ldg(p); // measurement build: real pointers, global memory only
#elif OMC_CUDA && defined(__CUDA_ARCH__)
  const char* base = geo_staged_bytes() ? reinterpret_cast<const char*>(g_geo_smem) : c_geo_blob;
  return *reinterpret_cast<const T*>(base + reinterpret_cast<uintptr_t>(p));
#else
  return *p;
#endif
}
#if OMC_CUDA
// every thread of the block, before any of them leaves the kernel
OMC_HD void stage_geometry()
{
#if defined(__CUDA_ARCH__)
  const uint32_t nb = geo_staged_bytes();
  if (nb) {
    const uint4* src = reinterpret_cast<const uint4*>(c_geo_blob);
    uint4* dst = reinterpret_cast<uint4*>(g_geo_smem);
    for (uint32_t k = threadIdx.x; k < (nb >> 4); k += blockDim.x)
      dst[k] = __ldg(src + k);
    __syncthreads();
  }
#endif
}
#endif

// =============================================================================
// RNG -- src/random_lcg.cpp:11-13 (constants), 32-44 (prn), 91-120 (future_seed)
// =============================================================================
constexpr uint64_t PRN_MULT = 6364136223846793005ULL;
constexpr uint64_t PRN_ADD = 1442695040888963407ULL;

OMC_BOTH double prn(uint64_t& seed)
{
  seed = PRN_MULT * seed + PRN_ADD;
  uint64_t word = ((seed >> ((seed >> 59u) + 5u)) ^ seed) * 12605985483714917081ull;
  uint64_t result = (word >> 43u) ^ word;
  // ldexp(result, -64): uint64 -> double rounds to nearest, the scale is exact
  return static_cast<double>(result) * 5.421010862427522170037264004349708557e-20;
}

OMC_BOTH uint64_t future_seed(uint64_t n, uint64_t seed)
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

// include/openmc/search.h:14-20 -- index of the interval containing `value`
template<class T>
OMC_HD int64_t lower_bound_index(const T* a, int64_t n, T value)
{
  if (ld(a) == value)
    return 0;
  int64_t lo = 0, len = n; // std::lower_bound
  while (len > 0) {
    int64_t half = len >> 1;
    if (ld(a + lo + half) < value) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  return lo - 1;
}

// src/math_functions.cpp:763-792
OMC_OOL V3 rotate_angle(V3 u, double mu, uint64_t& seed)
{
  double phi = 2.0 * PI * prn(seed);
  double sinphi = sin(phi);
  double cosphi = cos(phi);
  double a = sqrt(fmax(0., 1. - mu * mu));
  double b = sqrt(fmax(0., 1. - u.z * u.z));
  if (b > 1e-10) {
    return v3(mu * u.x + a * (u.x * u.z * cosphi - u.y * sinphi) / b,
      mu * u.y + a * (u.y * u.z * cosphi + u.x * sinphi) / b, mu * u.z - a * b * cosphi);
  } else {
    b = sqrt(1. - u.y * u.y);
    return v3(mu * u.x + a * (-u.x * u.y * sinphi + u.z * cosphi) / b, mu * u.y + a * b * sinphi,
      mu * u.z - a * (u.y * u.z * sinphi + u.x * cosphi) / b);
  }
}

// src/math_functions.cpp:98-121 (recursion fused with the sum; same operation order)
OMC_HD double evaluate_legendre(int n, const double* data, double x)
{
  double pm1 = 1.0, p = x;
  double val = 0.5 * ld(data) * 1.0; // (0 + 0.5) * data[0] * pnx[0]
  val = 0.0 + val;
  if (n >= 1)
    val += (1 + 0.5) * ld(data + 1) * p;
  for (int l = 1; l < n; l++) {
    double pn = ((2 * l + 1) * x * p - l * pm1) / (l + 1);
    val += ((l + 1) + 0.5) * ld(data + l + 1) * pn;
    pm1 = p;
    p = pn;
  }
  return val;
}

// src/random_dist.cpp:20-39
OMC_HD double maxwell_spectrum(double T, uint64_t& seed)
{
  double r1 = prn(seed);
  double r2 = prn(seed);
  double r3 = prn(seed);
  double c = cos(PI / 2. * r3);
  return -T * (log(r1) + log(r2) * c * c);
}
OMC_HD double watt_spectrum(double a, double b, uint64_t& seed)
{
  double w = maxwell_spectrum(a, seed);
  double xi = -1. + (1. - (-1.)) * prn(seed);
  return w + 0.25 * a * a * b + xi * sqrt(a * a * b * w);
}

// =============================================================================
// Surfaces -- src/surface.cpp:116-130, 207-220, 354-375, 399-456, 598-648
// =============================================================================
// General quadric A x^2 + B y^2 + C z^2 + D xy + E yz + F xz + G x + H y + J z + K = 0,
// src/surface.cpp:861-868 (evaluate), 870-939 (distance), 941-949 (normal); out of line: rare
OMC_OOL double general_quadric_evaluate(const double* c4, V3 r)
{
  const double* c = DM.quad_coef + 10 * (int)gld(c4);
  const double A = gld(c), B = gld(c + 1), C = gld(c + 2), D = gld(c + 3), E = gld(c + 4), F = gld(c + 5),
               G = gld(c + 6), H = gld(c + 7), J = gld(c + 8), K = gld(c + 9);
  const double x = r.x, y = r.y, z = r.z;
  return x * (A * x + D * y + G) + y * (B * y + E * z + H) + z * (C * z + F * x + J) + K;
}
OMC_OOL double general_quadric_distance(const double* c4, V3 r, V3 ang, bool coincident)
{
  const double* c = DM.quad_coef + 10 * (int)gld(c4);
  const double A = gld(c), B = gld(c + 1), C = gld(c + 2), D = gld(c + 3), E = gld(c + 4), F = gld(c + 5),
               G = gld(c + 6), H = gld(c + 7), J = gld(c + 8), K = gld(c + 9);
  const double x = r.x, y = r.y, z = r.z, u = ang.x, v = ang.y, w = ang.z;
  const double a = A * u * u + B * v * v + C * w * w + D * u * v + E * v * w + F * u * w;
  const double k = A * u * x + B * v * y + C * w * z +
                   0.5 * (D * (u * y + v * x) + E * (v * z + w * y) + F * (w * x + u * z) + G * u + H * v + J * w);
  const double cc = A * x * x + B * y * y + C * z * z + D * x * y + E * y * z + F * x * z + G * x + H * y + J * z + K;
  double quad = k * k - a * cc;
  double d;
  if (quad < 0.0) {
    return INFTY;
  } else if (coincident || fabs(cc) < FP_COINCIDENT) {
    if (a == 0.0)
      d = INFTY;
    else if (k >= 0.0)
      d = (-k - sqrt(quad)) / a;
    else
      d = (-k + sqrt(quad)) / a;
  } else if (a == 0.0) {
    d = -0.5 * cc / k;
  } else {
    quad = sqrt(quad);
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
OMC_OOL V3 general_quadric_normal(const double* c4, V3 r)
{
  const double* c = DM.quad_coef + 10 * (int)gld(c4);
  const double A = gld(c), B = gld(c + 1), C = gld(c + 2), D = gld(c + 3), E = gld(c + 4), F = gld(c + 5),
               G = gld(c + 6), H = gld(c + 7), J = gld(c + 8);
  return v3(2.0 * A * r.x + D * r.y + F * r.z + G, 2.0 * B * r.y + D * r.x + E * r.z + H,
    2.0 * C * r.z + E * r.y + F * r.x + J);
}

OMC_HD double surf_evaluate(int s, V3 r)
{
  const double* c = DM.surf_coef + OMCGPU_SURF_NCOEF * s;
  const double c0 = gld(c), c1 = gld(c + 1), c2 = gld(c + 2), c3 = gld(c + 3);
  switch (gld(DM.surf_type + s)) {
  case OMCGPU_SURF_XPLANE:
    return r.x - c0;
  case OMCGPU_SURF_YPLANE:
    return r.y - c0;
  case OMCGPU_SURF_ZPLANE:
    return r.z - c0;
  case OMCGPU_SURF_PLANE:
    return c0 * r.x + c1 * r.y + c2 * r.z - c3;
  case OMCGPU_SURF_XCYLINDER: {
    const double a = r.y - c0, b = r.z - c1;
    return a * a + b * b - c2 * c2;
  }
  case OMCGPU_SURF_YCYLINDER: {
    const double a = r.x - c0, b = r.z - c1;
    return a * a + b * b - c2 * c2;
  }
  case OMCGPU_SURF_ZCYLINDER: {
    const double a = r.x - c0, b = r.y - c1;
    return a * a + b * b - c2 * c2;
  }
  case OMCGPU_SURF_SPHERE: {
    const double x = r.x - c0, y = r.y - c1, z = r.z - c2;
    return x * x + y * y + z * z - c3 * c3;
  }
#if OMC_RARE_SURFACES
  case OMCGPU_SURF_XCONE: { // axis_aligned_cone_evaluate, src/surface.cpp:688-696
    const double r1 = r.x - c0, r2 = r.y - c1, r3 = r.z - c2;
    return r2 * r2 + r3 * r3 - c3 * r1 * r1;
  }
  case OMCGPU_SURF_YCONE: {
    const double r1 = r.y - c1, r2 = r.x - c0, r3 = r.z - c2;
    return r2 * r2 + r3 * r3 - c3 * r1 * r1;
  }
  case OMCGPU_SURF_QUADRIC:
    return general_quadric_evaluate(c, r);
  default: { // z-cone
    const double r1 = r.z - c2, r2 = r.x - c0, r3 = r.y - c1;
    return r2 * r2 + r3 * r3 - c3 * r1 * r1;
  }
#else
  default: // (cones and quadrics: extended build; omcgpu_init refuses them here)
    return 0.0;
#endif
  }
}

OMC_HD V3 surf_normal(int s, V3 r)
{
  const double* c = DM.surf_coef + OMCGPU_SURF_NCOEF * s;
  switch (gld(DM.surf_type + s)) {
  case OMCGPU_SURF_XPLANE:
    return v3(1., 0., 0.);
  case OMCGPU_SURF_YPLANE:
    return v3(0., 1., 0.);
  case OMCGPU_SURF_ZPLANE:
    return v3(0., 0., 1.);
  case OMCGPU_SURF_PLANE:
    return v3(gld(c), gld(c + 1), gld(c + 2));
  case OMCGPU_SURF_XCYLINDER:
    return v3(0.0, 2.0 * (r.y - gld(c)), 2.0 * (r.z - gld(c + 1)));
  case OMCGPU_SURF_YCYLINDER:
    return v3(2.0 * (r.x - gld(c)), 0.0, 2.0 * (r.z - gld(c + 1)));
  case OMCGPU_SURF_ZCYLINDER:
    return v3(2.0 * (r.x - gld(c)), 2.0 * (r.y - gld(c + 1)), 0.0);
  case OMCGPU_SURF_SPHERE:
    return v3(2.0 * (r.x - gld(c)), 2.0 * (r.y - gld(c + 1)), 2.0 * (r.z - gld(c + 2)));
#if OMC_RARE_SURFACES
  case OMCGPU_SURF_XCONE: // axis_aligned_cone_normal, src/surface.cpp:747-756
    return v3(-2.0 * gld(c + 3) * (r.x - gld(c)), 2.0 * (r.y - gld(c + 1)), 2.0 * (r.z - gld(c + 2)));
  case OMCGPU_SURF_YCONE:
    return v3(2.0 * (r.x - gld(c)), -2.0 * gld(c + 3) * (r.y - gld(c + 1)), 2.0 * (r.z - gld(c + 2)));
  case OMCGPU_SURF_QUADRIC:
    return general_quadric_normal(c, r);
  default:
    return v3(2.0 * (r.x - gld(c)), 2.0 * (r.y - gld(c + 1)), -2.0 * gld(c + 3) * (r.z - gld(c + 2)));
#else
  default:
    return v3(0., 0., 1.);
#endif
  }
}

OMC_HD bool surf_sense(int s, V3 r, V3 u)
{
  const double f = surf_evaluate(s, r);
  if (fabs(f) < FP_COINCIDENT)
    return dot(u, surf_normal(s, r)) > 0.0;
  return f > 0.0;
}

OMC_HD double plane_distance(double f, double proj, bool coincident)
{ // f = offset - r_i (axis planes) ; general plane passes f = -(A x + .. - D)
  if (coincident || fabs(f) < FP_COINCIDENT || proj == 0.0)
    return INFTY;
  const double d = f / proj;
  return d < 0.0 ? INFTY : d;
}

// quadratic a t^2 + 2 k t + c = 0 shared by cylinders (a = 1 - u_axis^2) and
// the sphere (a = 1, no division): src/surface.cpp:399-445, 606-643
template<bool DIVIDE>
OMC_HD double quadric_distance(double a, double k, double c, bool coincident)
{
  const double quad = k * k - a * c;
  if (quad < 0.0)
    return INFTY;
  if (coincident || fabs(c) < FP_COINCIDENT) {
    if (k >= 0.0)
      return INFTY;
    return DIVIDE ? (-k + sqrt(quad)) / a : -k + sqrt(quad);
  }
  if (c < 0.0)
    return DIVIDE ? (-k + sqrt(quad)) / a : -k + sqrt(quad);
  const double d = DIVIDE ? (-k - sqrt(quad)) / a : -k - sqrt(quad);
  return d < 0.0 ? INFTY : d;
}

OMC_HD double cylinder_distance(double u_axis, double q2, double q3, double u2, double u3,
  double radius, bool coincident)
{
  const double a = 1.0 - u_axis * u_axis;
  if (a == 0.0)
    return INFTY;
  const double k = q2 * u2 + q3 * u3;
  const double c = q2 * q2 + q3 * q3 - radius * radius;
  return quadric_distance<true>(a, k, c, coincident);
}

// axis_aligned_cone_distance, src/surface.cpp:701-742 (index 1 = the cone's axis)
OMC_HD double cone_distance(double r1, double r2, double r3, double u1, double u2, double u3, double radius_sq,
  bool coincident)
{
  const double a = u2 * u2 + u3 * u3 - radius_sq * u1 * u1;
  const double k = r2 * u2 + r3 * u3 - radius_sq * r1 * u1;
  const double c = r2 * r2 + r3 * r3 - radius_sq * r1 * r1;
  double quad = k * k - a * c;
  double d;
  if (quad < 0.0) {
    return INFTY;
  } else if (coincident || fabs(c) < FP_COINCIDENT) {
    if (k >= 0.0)
      d = (-k - sqrt(quad)) / a;
    else
      d = (-k + sqrt(quad)) / a;
  } else {
    quad = sqrt(quad);
    d = (-k - quad) / a;
    const double b = (-k + quad) / a;
    if (d < 0.0) {
      if (b > 0.0)
        d = b;
    } else if (b > 0.0 && b < d) {
      d = b;
    }
  }
  if (d <= 0.0)
    return INFTY;
  return d;
}

OMC_HD double surf_distance(int s, V3 r, V3 u, bool coincident)
{
  const double* c = DM.surf_coef + OMCGPU_SURF_NCOEF * s;
  const double c0 = gld(c), c1 = gld(c + 1), c2 = gld(c + 2), c3 = gld(c + 3);
  switch (gld(DM.surf_type + s)) {
  case OMCGPU_SURF_XPLANE:
    return plane_distance(c0 - r.x, u.x, coincident);
  case OMCGPU_SURF_YPLANE:
    return plane_distance(c0 - r.y, u.y, coincident);
  case OMCGPU_SURF_ZPLANE:
    return plane_distance(c0 - r.z, u.z, coincident);
  case OMCGPU_SURF_PLANE: {
    const double f = c0 * r.x + c1 * r.y + c2 * r.z - c3;
    const double projection = c0 * u.x + c1 * u.y + c2 * u.z;
    if (coincident || fabs(f) < FP_COINCIDENT || projection == 0.0)
      return INFTY;
    const double d = -f / projection;
    return d < 0.0 ? INFTY : d;
  }
  case OMCGPU_SURF_XCYLINDER:
    return cylinder_distance(u.x, r.y - c0, r.z - c1, u.y, u.z, c2, coincident);
  case OMCGPU_SURF_YCYLINDER:
    return cylinder_distance(u.y, r.x - c0, r.z - c1, u.x, u.z, c2, coincident);
  case OMCGPU_SURF_ZCYLINDER:
    return cylinder_distance(u.z, r.x - c0, r.y - c1, u.x, u.y, c2, coincident);
  case OMCGPU_SURF_SPHERE: {
    const double x = r.x - c0, y = r.y - c1, z = r.z - c2;
    const double k = x * u.x + y * u.y + z * u.z;
    const double cc = x * x + y * y + z * z - c3 * c3;
    return quadric_distance<false>(1.0, k, cc, coincident);
  }
#if OMC_RARE_SURFACES
  case OMCGPU_SURF_XCONE:
    return cone_distance(r.x - c0, r.y - c1, r.z - c2, u.x, u.y, u.z, c3, coincident);
  case OMCGPU_SURF_YCONE:
    return cone_distance(r.y - c1, r.x - c0, r.z - c2, u.y, u.x, u.z, c3, coincident);
  case OMCGPU_SURF_QUADRIC:
    return general_quadric_distance(c, r, u, coincident);
  default:
    return cone_distance(r.z - c2, r.x - c0, r.y - c1, u.z, u.x, u.y, c3, coincident);
#else
  default:
    return INFTY;
#endif
  }
}

// Surface::diffuse_reflect, src/surface.cpp:142-161: cosine-distributed direction about the
// surface normal n (not normalised on entry), on the side the particle came from
OMC_OOL V3 diffuse_reflect(V3 n, V3 u, uint64_t& seed)
{
  const double nn = sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
  n = v3(n.x / nn, n.y / nn, n.z / nn);
  const double projection = dot(n, u);
  const double mu = (projection >= 0.0) ? -sqrt(prn(seed)) : sqrt(prn(seed));
  V3 d = rotate_angle(n, mu, seed);
  const double dn = sqrt(d.x * d.x + d.y * d.y + d.z * d.z);
  return v3(d.x / dn, d.y / dn, d.z / dn);
}

// =============================================================================
// Regions -- src/cell.cpp:910-936 (distance), 940-970 (contains_simple),
// 974-1029 (contains_complex)
// =============================================================================
OMC_HD bool cell_contains(int c, V3 r, V3 u, int on_surface)
{
  const int t0 = gld(DM.cell_region_ofs + c), t1 = gld(DM.cell_region_ofs + c + 1);
  const int32_t* tok = DM.region_tokens;
  if (gld(DM.cell_simple + c)) {
    for (int it = t0; it < t1; ++it) {
      const int32_t token = gld(tok + it);
      if (token == on_surface)
        continue;
      if (-token == on_surface)
        return false;
      if (surf_sense(abs(token) - 1, r, u) != (token > 0))
        return false;
    }
    return true;
  }
  bool in_cell = true;
  int total_depth = 0;
  for (int it = t0; it < t1; ++it) {
    const int32_t token = gld(tok + it);
    if (token < OP_UNION) {
      if (token == on_surface)
        in_cell = true;
      else if (-token == on_surface)
        in_cell = false;
      else
        in_cell = (surf_sense(abs(token) - 1, r, u) == (token > 0));
    } else if ((token == OP_UNION && in_cell) || (token == OP_INTERSECTION && !in_cell)) {
      if (total_depth == 0)
        return in_cell;
      total_depth--;
      int depth = 1;
      do {
        ++it;
        const int32_t next_token = gld(tok + it);
        if (next_token > OP_COMPLEMENT)
          depth += (next_token == OP_RIGHT_PAREN) ? -1 : 1;
      } while (depth > 0);
    } else if (token == OP_LEFT_PAREN) {
      total_depth++;
    } else if (token == OP_RIGHT_PAREN) {
      total_depth--;
    }
  }
  return in_cell;
}

OMC_HD void cell_distance(int c, V3 r, V3 u, int on_surface, double& min_dist, int& i_surf)
{
  min_dist = INFTY;
  i_surf = 2147483647;
  const int t0 = gld(DM.cell_region_ofs + c), t1 = gld(DM.cell_region_ofs + c + 1);
  const int abs_on = abs(on_surface);
  for (int it = t0; it < t1; ++it) {
    const int32_t token = gld(DM.region_tokens + it);
    if (token >= OP_UNION)
      continue;
    const double d = surf_distance(abs(token) - 1, r, u, abs(token) == abs_on);
    if (d < min_dist) {
      if (min_dist - d >= FP_PRECISION * min_dist) {
        min_dist = d;
        i_surf = -token;
      }
    }
  }
}

// =============================================================================
// Rectangular lattices -- src/lattice.cpp:243-248, 252-306, 310-342, 352-377
// =============================================================================
// Position::rotate with the fill cell's matrix, include/openmc/position.h:97-104
OMC_HD V3 cell_rotate(int c, V3 v)
{
  const double* m = DM.cell_rotation + 9 * c;
  return v3(v.x * gld(m) + v.y * gld(m + 1) + v.z * gld(m + 2), v.x * gld(m + 3) + v.y * gld(m + 4) + v.z * gld(m + 5),
    v.x * gld(m + 6) + v.y * gld(m + 7) + v.z * gld(m + 8));
}

struct LatIdx {
  int x, y, z;
};

// ---- hexagonal lattices, src/lattice.cpp:727-1007: rare, out of line; lat_shape = (2 n_rings - 1,
// 2 n_rings - 1, n_axial), lat_pitch = (radial, axial), lat_lower_left = the centre ----------------
// Compiled in only with -DOMC_EXTENDED=1 (libomcgpu_ext.so; the loaders pick it for models with
// hexagonal lattices or secondary neutrons): the dispatch in the lattice functions cost every event kernel 2-4 % in
// registers and spills even for models without a lattice (profiles/r68-r69), so the default
// build does not carry it and omcgpu_init refuses such models there, naming the other build.

OMC_HD bool lat_is_hex(int l)
{
#if OMC_HEX_LATTICES
  return DM.n_hex_lattices > 0 && gld(DM.lat_type + l) != OMCGPU_LATTICE_RECT;
#else
  (void)l;
  return false;
#endif
}
OMC_OOL bool hex_valid(int l, LatIdx i)
{ // HexLattice::are_valid_indices :727-735
  const int n_rings = (gld(DM.lat_shape + 3 * l) + 1) / 2;
  return (i.x >= 0) && (i.y >= 0) && (i.z >= 0) && (i.x < 2 * n_rings - 1) && (i.y < 2 * n_rings - 1) &&
         (i.x + i.y > n_rings - 2) && (i.x + i.y < 3 * n_rings - 2) && (i.z < gld(DM.lat_shape + 3 * l + 2));
}
OMC_OOL V3 hex_local_position(int l, V3 r, LatIdx i)
{ // HexLattice::get_local_position :983-1007
  const double* c = DM.lat_lower_left + 3 * l;
  const double p0 = gld(DM.lat_pitch + 3 * l);
  const int n_rings = (gld(DM.lat_shape + 3 * l) + 1) / 2;
  if (gld(DM.lat_type + l) == OMCGPU_LATTICE_HEX_Y) {
    r.x -= gld(c) + sqrt(3.0) / 2.0 * (i.x - n_rings + 1) * p0;
    r.y -= (gld(c + 1) + (i.y - n_rings + 1) * p0 + (i.x - n_rings + 1) * p0 / 2.0);
  } else {
    r.x -= (gld(c) + (i.x - n_rings + 1) * p0 + (i.y - n_rings + 1) * p0 / 2.0);
    r.y -= gld(c + 1) + sqrt(3.0) / 2.0 * (i.y - n_rings + 1) * p0;
  }
  if (gld(DM.lat_is3d + l)) {
    const int n_axial = gld(DM.lat_shape + 3 * l + 2);
    r.z -= gld(c + 2) - (0.5 * n_axial - i.z - 0.5) * gld(DM.lat_pitch + 3 * l + 1);
  }
  return r;
}
OMC_OOL LatIdx hex_get_indices(int l, V3 r, V3 u)
{ // HexLattice::get_indices :879-973
  const double* c = DM.lat_lower_left + 3 * l;
  const double p0 = gld(DM.lat_pitch + 3 * l);
  const int n_rings = (gld(DM.lat_shape + 3 * l) + 1) / 2;
  const bool is_3d = gld(DM.lat_is3d + l);
  V3 r_o = v3(r.x - gld(c), r.y - gld(c + 1), r.z);
  if (is_3d)
    r_o.z -= gld(c + 2);
  LatIdx res;
  res.z = 0;
  if (is_3d) {
    const int n_axial = gld(DM.lat_shape + 3 * l + 2);
    const double iz_ = r_o.z / gld(DM.lat_pitch + 3 * l + 1) + 0.5 * n_axial;
    const long long iz_close = llround(iz_);
    if (fabs(iz_ - (double)iz_close) < FP_COINCIDENT)
      res.z = (u.z > 0) ? (int)iz_close : (int)iz_close - 1;
    else
      res.z = (int)floor(iz_);
  }
  if (gld(DM.lat_type + l) == OMCGPU_LATTICE_HEX_Y) {
    const double alpha = r_o.y - r_o.x / sqrt(3.0);
    res.x = (int)floor(r_o.x / (0.5 * sqrt(3.0) * p0));
    res.y = (int)floor(alpha / p0);
  } else {
    const double alpha = r_o.y - r_o.x * sqrt(3.0);
    res.x = (int)floor(-alpha / (sqrt(3.0) * p0));
    res.y = (int)floor(r_o.y / (0.5 * sqrt(3.0) * p0));
  }
  res.x += n_rings - 1;
  res.y += n_rings - 1;
  int i1_chg = 0, i2_chg = 0;
  double d_min = INFTY, dp_min = INFTY;
  for (int i = 0; i < 2; i++) {
    for (int j = 0; j < 2; j++) {
      LatIdx t;
      t.x = res.x + j;
      t.y = res.y + i;
      t.z = 0;
      V3 r_t = hex_local_position(l, r, t);
      const double d = r_t.x * r_t.x + r_t.y * r_t.y;
      const bool on_boundary = fabs(1.0 - d_min / d) < FP_COINCIDENT;
      if (d < d_min || on_boundary) {
        const double nrm = sqrt(d);
        r_t = v3(r_t.x / nrm, r_t.y / nrm, r_t.z / nrm);
        const double dp = u.x * r_t.x + u.y * r_t.y;
        if (on_boundary && dp > dp_min)
          continue;
        d_min = d;
        i1_chg = j;
        i2_chg = i;
        dp_min = dp;
      }
    }
  }
  res.x += i1_chg;
  res.y += i2_chg;
  return res;
}
OMC_HD int pack_trans(int tx, int ty, int tz);
// HexLattice::distance :738-875: r is the position at the lattice's own level (z: the tile's), idx
// the current tile -- distances are measured in the frames of the NEIGHBOUR tiles
struct HexDist {
  double d;
  int trans;
};
OMC_HD HexDist hex_distance(int l, V3 r, V3 u, LatIdx idx)
{
  int trans;
  const double p0 = gld(DM.lat_pitch + 3 * l);
  const bool oy = gld(DM.lat_type + l) == OMCGPU_LATTICE_HEX_Y;
  double beta_dir, gamma_dir, delta_dir;
  if (oy) {
    beta_dir = u.x * sqrt(3.0) / 2.0 + u.y / 2.0;
    gamma_dir = u.x * sqrt(3.0) / 2.0 - u.y / 2.0;
    delta_dir = u.y;
  } else {
    beta_dir = u.x;
    gamma_dir = u.x / 2.0 - u.y * sqrt(3.0) / 2.0;
    delta_dir = u.x / 2.0 + u.y * sqrt(3.0) / 2.0;
  }
  double d = INFTY;
  trans = pack_trans(0, 0, 0);
  LatIdx t = idx;
  double edge = -copysign(0.5 * p0, beta_dir);
  t.x = idx.x + (beta_dir > 0 ? 1 : -1);
  V3 r_t = hex_local_position(l, r, t);
  const double beta = oy ? r_t.x * sqrt(3.0) / 2.0 + r_t.y / 2.0 : r_t.x;
  if ((fabs(beta - edge) > FP_PRECISION) && beta_dir != 0) {
    d = (edge - beta) / beta_dir;
    trans = pack_trans(beta_dir > 0 ? 1 : -1, 0, 0);
  }
  edge = -copysign(0.5 * p0, gamma_dir);
  t.x = idx.x + (gamma_dir > 0 ? 1 : -1);
  t.y = idx.y - (gamma_dir > 0 ? 1 : -1);
  r_t = hex_local_position(l, r, t);
  const double gamma = oy ? r_t.x * sqrt(3.0) / 2.0 - r_t.y / 2.0 : r_t.x / 2.0 - r_t.y * sqrt(3.0) / 2.0;
  if ((fabs(gamma - edge) > FP_PRECISION) && gamma_dir != 0) {
    const double this_d = (edge - gamma) / gamma_dir;
    if (this_d < d) {
      trans = pack_trans(gamma_dir > 0 ? 1 : -1, gamma_dir > 0 ? -1 : 1, 0);
      d = this_d;
    }
  }
  edge = -copysign(0.5 * p0, delta_dir);
  t.x = idx.x;
  t.y = idx.y + (delta_dir > 0 ? 1 : -1);
  r_t = hex_local_position(l, r, t);
  const double delta = oy ? r_t.y : r_t.x / 2.0 + r_t.y * sqrt(3.0) / 2.0;
  if ((fabs(delta - edge) > FP_PRECISION) && delta_dir != 0) {
    const double this_d = (edge - delta) / delta_dir;
    if (this_d < d) {
      trans = pack_trans(0, delta_dir > 0 ? 1 : -1, 0);
      d = this_d;
    }
  }
  if (gld(DM.lat_is3d + l)) {
    const double z0 = copysign(0.5 * gld(DM.lat_pitch + 3 * l + 1), u.z);
    if ((fabs(r.z - z0) > FP_PRECISION) && u.z != 0) {
      const double this_d = (z0 - r.z) / u.z;
      if (this_d < d) {
        d = this_d;
        trans = pack_trans(0, 0, u.z > 0 ? 1 : -1);
      }
    }
  }
  HexDist out;
  out.d = d;
  out.trans = trans;
  return out;
}
OMC_HD V3 cell_rotate(int c, V3 v);
// src/geometry.cpp:392-400: the distance is taken from the position at the lattice's OWN level --
// the parent level's position through the fill cell's translation / rotation, z of the tile
OMC_OOL HexDist hex_level_distance(int l, int c_up, V3 r_up, double z_local, V3 u, int ix, int iy, int iz)
{
  const double* t = DM.cell_translation + 3 * c_up;
  V3 r_hex = v3(r_up.x - gld(t), r_up.y - gld(t + 1), r_up.z - gld(t + 2));
  if (gld(DM.cell_rotated + c_up))
    r_hex = cell_rotate(c_up, r_hex);
  r_hex.z = z_local;
  LatIdx li;
  li.x = ix;
  li.y = iy;
  li.z = iz;
  return hex_distance(l, r_hex, u, li);
}

OMC_HD bool rect_valid(int l, LatIdx i)
{
  const int32_t* n = DM.lat_shape + 3 * l;
  return i.x >= 0 && i.x < gld(n) && i.y >= 0 && i.y < gld(n + 1) && i.z >= 0 && i.z < gld(n + 2);
}
OMC_HD bool lat_valid(int l, LatIdx i)
{
  if (lat_is_hex(l))
    return hex_valid(l, i);
  return rect_valid(l, i);
}
OMC_HD int lat_flat(int l, LatIdx i)
{
  const int32_t* n = DM.lat_shape + 3 * l;
  const int nx = gld(n), ny = gld(n + 1);
  return nx * ny * i.z + nx * i.y + i.x;
}
OMC_HD int lat_universe(int l, LatIdx i)
{
  return gld(DM.lat_universes + gld(DM.lat_univ_ofs + l) + lat_flat(l, i));
}
OMC_HD int lat_axis_index(double r, double ll, double pitch, double u)
{
  const double ix_ = (r - ll) / pitch;
  const long long ix_close = llround(ix_);
  if (fabs(ix_ - (double)ix_close) < FP_COINCIDENT)
    return (u > 0) ? (int)ix_close : (int)ix_close - 1;
  return (int)floor(ix_);
}
OMC_HD LatIdx lat_get_indices(int l, V3 r, V3 u)
{
  if (lat_is_hex(l))
    return hex_get_indices(l, r, u);
  const double* ll = DM.lat_lower_left + 3 * l;
  const double* p = DM.lat_pitch + 3 * l;
  LatIdx i;
  i.x = lat_axis_index(r.x, gld(ll), gld(p), u.x);
  i.y = lat_axis_index(r.y, gld(ll + 1), gld(p + 1), u.y);
  i.z = gld(DM.lat_is3d + l) ? lat_axis_index(r.z, gld(ll + 2), gld(p + 2), u.z) : 0;
  return i;
}
OMC_HD V3 rect_local_position(int l, V3 r, LatIdx i);
OMC_HD V3 lat_local_position(int l, V3 r, LatIdx i)
{
  if (lat_is_hex(l))
    return hex_local_position(l, r, i);
  return rect_local_position(l, r, i);
}
OMC_HD V3 rect_local_position(int l, V3 r, LatIdx i)
{
  const double* ll = DM.lat_lower_left + 3 * l;
  const double* p = DM.lat_pitch + 3 * l;
  r.x -= (gld(ll) + (i.x + 0.5) * gld(p));
  r.y -= (gld(ll + 1) + (i.y + 0.5) * gld(p + 1));
  if (gld(DM.lat_is3d + l))
    r.z -= (gld(ll + 2) + (i.z + 0.5) * gld(p + 2));
  return r;
}
// returns distance; translation packed as (tx+1) | (ty+1)<<2 | (tz+1)<<4
OMC_HD int pack_trans(int tx, int ty, int tz)
{
  return (tx + 1) | ((ty + 1) << 2) | ((tz + 1) << 4);
}
constexpr int TRANS_NONE = 1 | (1 << 2) | (1 << 4);
OMC_HD double lat_distance(int l, V3 r, V3 u, int& trans)
{
  const double* p = DM.lat_pitch + 3 * l;
  const double x0 = copysign(0.5 * gld(p), u.x);
  const double y0 = copysign(0.5 * gld(p + 1), u.y);
  double d = INFTY;
  trans = TRANS_NONE;
  if ((fabs(r.x - x0) > FP_PRECISION) && u.x != 0) {
    d = (x0 - r.x) / u.x;
    trans = pack_trans(u.x > 0 ? 1 : -1, 0, 0);
  }
  if ((fabs(r.y - y0) > FP_PRECISION) && u.y != 0) {
    const double this_d = (y0 - r.y) / u.y;
    if (this_d < d) {
      d = this_d;
      trans = pack_trans(0, u.y > 0 ? 1 : -1, 0);
    }
  }
  if (gld(DM.lat_is3d + l)) {
    const double z0 = copysign(0.5 * gld(p + 2), u.z);
    if ((fabs(r.z - z0) > FP_PRECISION) && u.z != 0) {
      const double this_d = (z0 - r.z) / u.z;
      if (this_d < d) {
        d = this_d;
        trans = pack_trans(0, 0, u.z > 0 ? 1 : -1);
      }
    }
  }
  return d;
}

// =============================================================================
// Per-particle accessors over the SoA columns
// =============================================================================
#define LV(col, lvl) (ps.col[(int64_t)(lvl)*ps.cap + i])

OMC_HD V3 get_r(const PState& ps, int64_t i, int lvl)
{
  return v3(LV(rx, lvl), LV(ry, lvl), LV(rz, lvl));
}
OMC_HD V3 get_u(const PState& ps, int64_t i, int lvl)
{
  return v3(LV(ux, lvl), LV(uy, lvl), LV(uz, lvl));
}
OMC_HD void set_r(const PState& ps, int64_t i, int lvl, V3 r)
{
  LV(rx, lvl) = r.x;
  LV(ry, lvl) = r.y;
  LV(rz, lvl) = r.z;
}
OMC_HD void set_u(const PState& ps, int64_t i, int lvl, V3 u)
{
  LV(ux, lvl) = u.x;
  LV(uy, lvl) = u.y;
  LV(uz, lvl) = u.z;
}
// Hexagonal lattices at a particle's level l, for distance_to_boundary and cross_lattice.  For a
// particle in the global columns (the list kernels) the out-of-line functions take (i, l) and read
// the state through the constant-memory copy of the column pointers: the call sites hold two
// integers, not twenty registers of arguments (measured: passing the values cost the event
// kernels 2-4 %, profiles/r68).  A particle staged in thread-local memory (tail kernel) passes values.
OMC_HD HexDist hex_distance_values(const PState& ps, int64_t i, int l)
{
  LatIdx li;
  li.x = LV(lat_x, l);
  li.y = LV(lat_y, l);
  li.z = LV(lat_z, l);
  return hex_level_distance(LV(lattice, l), LV(cell, l - 1), v3(LV(rx, l - 1), LV(ry, l - 1), LV(rz, l - 1)),
    LV(rz, l), v3(LV(ux, l), LV(uy, l), LV(uz, l)), li.x, li.y, li.z);
}
// cross_lattice for a hex lattice: local position in the new tile; false when the tile is outside
OMC_HD bool hex_cross_values(const PState& ps, int64_t i, int lvl)
{
  const int lat = LV(lattice, lvl);
  LatIdx li;
  li.x = LV(lat_x, lvl);
  li.y = LV(lat_y, lvl);
  li.z = LV(lat_z, lvl);
  const int c_up = LV(cell, lvl - 1);
  const double* t = DM.cell_translation + 3 * c_up;
  V3 r = v3(LV(rx, lvl - 1) - gld(t), LV(ry, lvl - 1) - gld(t + 1), LV(rz, lvl - 1) - gld(t + 2));
  if (gld(DM.cell_rotated + c_up))
    r = cell_rotate(c_up, r);
  r = hex_local_position(lat, r, li);
  LV(rx, lvl) = r.x;
  LV(ry, lvl) = r.y;
  LV(rz, lvl) = r.z;
  if (!hex_valid(lat, li))
    return false;
  LV(universe, lvl) = lat_universe(lat, li);
  return true;
}
#if OMC_CUDA
__device__ __noinline__ HexDist hex_distance_ool(int64_t i, int l)
{
  return hex_distance_values(c_ps, i, l);
}
__device__ __noinline__ bool hex_cross_ool(int64_t i, int lvl)
{
  return hex_cross_values(c_ps, i, lvl);
}
#endif
OMC_HD HexDist hex_distance_at_level(const PState& ps, int64_t i, int l)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  if (ps.is_local == 0)
    return hex_distance_ool(i, l);
#endif
  return hex_distance_values(ps, i, l);
}
OMC_HD bool hex_cross(const PState& ps, int64_t i, int lvl)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  if (ps.is_local == 0)
    return hex_cross_ool(i, lvl);
#endif
  return hex_cross_values(ps, i, lvl);
}

OMC_HD void reset_level(const PState& ps, int64_t i, int lvl)
{ // LocalCoord::reset, include/openmc/particle_data.h
  LV(cell, lvl) = C_NONE;
  LV(universe, lvl) = C_NONE;
  LV(lattice, lvl) = C_NONE;
  LV(lat_x, lvl) = 0;
  LV(lat_y, lvl) = 0;
  LV(lat_z, lvl) = 0;
}

// =============================================================================
// One particle's state staged in thread-local memory, laid out as the sector groups of PState so
// that the very same event functions run on it: a "local view" is a PState whose columns point at
// the thread's copy (index 0).  Re-reads of a field an earlier event of the kernel wrote are then
// L1 hits of the thread's own local memory instead of round trips to L2 for gathered global
// sectors.  NL > 0: everything staged, NL coordinate levels (the tail kernel keeps a history here
// from take-over to death).  NL == 0: only the groups P1 P2 P4 P5 (128 bytes); coordinate levels
// and the k accumulators stay in the global columns, pre-offset to the particle.
// =============================================================================
template<int NL>
struct LevelStore {
  alignas(32) double gA[NL][4];
  alignas(32) double gB[NL][4];
  alignas(16) int32_t gC[NL][4];
  alignas(32) double p3[4];
};
template<>
struct LevelStore<0> {};
template<int NL>
struct LocalParticle : LevelStore<NL> {
  alignas(32) double p1[4], p2[4];
  alignas(32) int32_t p4[8];
  alignas(32) double p5[4];
  int32_t n_progeny;
  uint8_t next_ev;
};
template<int NL>
OMC_HD PState local_view(LocalParticle<NL>& lp, const PState& g, int64_t index)
{
  PState ps;
  ps.id_shift = index;
  ps.is_local = NL + 1;
  ps.n_prefix = g.n_prefix;
  ps.xs_prefix = g.xs_prefix;
#define D_(col, base, slot) ps.col.p = reinterpret_cast<double*>(base) + (slot)
#define I_(col, base, slot) ps.col.p = reinterpret_cast<int32_t*>(base) + (slot)
  if constexpr (NL > 0) {
    ps.cap = 1;
    D_(rx, lp.gA, 0); D_(ry, lp.gA, 1); D_(rz, lp.gA, 2); I_(cell, lp.gA, 6); I_(lattice, lp.gA, 7);
    D_(ux, lp.gB, 0); D_(uy, lp.gB, 1); D_(uz, lp.gB, 2); I_(universe, lp.gB, 6);
    I_(lat_x, lp.gC, 0); I_(lat_y, lp.gC, 1); I_(lat_z, lp.gC, 2);
    D_(k_abs, lp.p3, 0); D_(k_col, lp.p3, 1); D_(k_trk, lp.p3, 2); D_(k_leak, lp.p3, 3);
    ps.n_progeny = &lp.n_progeny;
  } else { // global columns, pre-offset so that index 0 is this particle
    ps.cap = g.cap;
    ps.rx.p = &g.rx[index]; ps.ry.p = &g.ry[index]; ps.rz.p = &g.rz[index];
    ps.cell.p = &g.cell[index]; ps.lattice.p = &g.lattice[index];
    ps.ux.p = &g.ux[index]; ps.uy.p = &g.uy[index]; ps.uz.p = &g.uz[index]; ps.universe.p = &g.universe[index];
    ps.lat_x.p = &g.lat_x[index]; ps.lat_y.p = &g.lat_y[index]; ps.lat_z.p = &g.lat_z[index];
    ps.k_abs.p = &g.k_abs[index]; ps.k_col.p = &g.k_col[index]; ps.k_trk.p = &g.k_trk[index];
    ps.k_leak.p = &g.k_leak[index];
    ps.n_progeny = g.n_progeny + index;
  }
  D_(E, lp.p1, 0); D_(wgt, lp.p1, 1); D_(time, lp.p1, 2);
  ps.seed.p = reinterpret_cast<uint64_t*>(lp.p1) + 3;
  D_(xs_total, lp.p2, 0); D_(xs_absorption, lp.p2, 1); D_(xs_fission, lp.p2, 2); D_(xs_nu_fission, lp.p2, 3);
  I_(n_coord, lp.p4, 0); I_(surface, lp.p4, 1); I_(material, lp.p4, 2); I_(material_last, lp.p4, 3);
  I_(cell_instance, lp.p4, 4); I_(cell_last0, lp.p4, 5); I_(g, lp.p4, 6); I_(n_event, lp.p4, 7);
  D_(sqrtkT, lp.p5, 0); D_(sqrtkT_last, lp.p5, 1);
  I_(b_surface, lp.p5, 4); I_(b_level_trans, lp.p5, 5); I_(g_last, lp.p5, 6);
#undef D_
#undef I_
  ps.next_ev = &lp.next_ev;
  return ps;
}
// whole sector groups, 16 bytes at a time
OMC_HD void copy_group(void* dst, const void* src, int bytes)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  for (int k = 0; k < bytes / 16; ++k)
    reinterpret_cast<int4*>(dst)[k] = reinterpret_cast<const int4*>(src)[k];
#else
  std::memcpy(dst, src, (size_t)bytes);
#endif
}
template<int NL>
OMC_HD void load_local(const PState& g, int64_t i, LocalParticle<NL>& lp)
{
  if constexpr (NL > 0) {
    const int L = DM.n_coord_levels < NL ? DM.n_coord_levels : NL;
    for (int l = 0; l < L; ++l) {
      copy_group(lp.gA[l], &g.rx[(int64_t)l * g.cap + i], 32);
      copy_group(lp.gB[l], &g.ux[(int64_t)l * g.cap + i], 32);
      copy_group(lp.gC[l], &g.lat_x[(int64_t)l * g.cap + i], 16);
    }
    copy_group(lp.p3, &g.k_abs[i], 32);
    lp.n_progeny = g.n_progeny[i];
  }
  copy_group(lp.p1, &g.E[i], 32);
  copy_group(lp.p2, &g.xs_total[i], 32);
  copy_group(lp.p4, &g.n_coord[i], 32);
  copy_group(lp.p5, &g.sqrtkT[i], 32);
  lp.next_ev = 0;
}
// every staged group back to the global columns (whole sectors: no partial-sector writes)
template<int NL>
OMC_HD void store_local_all(const PState& g, int64_t i, const LocalParticle<NL>& lp)
{
  if constexpr (NL > 0) {
    const int L = DM.n_coord_levels < NL ? DM.n_coord_levels : NL;
    for (int l = 0; l < L; ++l) {
      copy_group(&g.rx[(int64_t)l * g.cap + i], lp.gA[l], 32);
      copy_group(&g.ux[(int64_t)l * g.cap + i], lp.gB[l], 32);
      copy_group(&g.lat_x[(int64_t)l * g.cap + i], lp.gC[l], 16);
    }
    copy_group(&g.k_abs[i], lp.p3, 32);
    g.n_progeny[i] = lp.n_progeny;
  }
  copy_group(&g.E[i], lp.p1, 32);
  copy_group(&g.xs_total[i], lp.p2, 32);
  copy_group(&g.n_coord[i], lp.p4, 32);
  copy_group(&g.sqrtkT[i], lp.p5, 32);
}
// what the rest of the generation reads of a finished history: the k / leakage accumulators and
// the event count (k_reduce), the progeny count (fission-bank sort)
template<int NL>
OMC_HD void store_local_results(const PState& g, int64_t i, const LocalParticle<NL>& lp)
{
  static_assert(NL > 0, "the tail kernel stages everything");
  copy_group(&g.k_abs[i], lp.p3, 32);
  copy_group(&g.n_coord[i], lp.p4, 32);
  copy_group(&g.E[i], lp.p1, 32); // (weight 0 = dead: what a later probe of the state expects)
  g.n_progeny[i] = lp.n_progeny;
}

// =============================================================================
// Locate -- src/geometry.cpp:66-98, 102-252, 256-297 ; src/universe.cpp:40-60
// The walk keeps only the current level in registers and streams the deeper
// levels straight into their SoA columns.  Neighbour lists only change the
// search order inside a universe, so the universe's own cell list is searched.
// =============================================================================
OMC_HD int cell_instance_at_level(const PState& ps, int64_t i, int level)
{
  const int c = LV(cell, level);
  const int map = gld(DM.cell_distribcell_index + c);
  if (map == C_NONE)
    return C_NONE;
  int instance = 0;
  for (int k = 0; k < level; k++) {
    const int c_k = LV(cell, k);
    const int type = gld(DM.cell_type + c_k);
    if (type == OMCGPU_FILL_UNIVERSE) {
      instance += gld(DM.cell_offset + gld(DM.cell_offset_ofs + c_k) + map);
    } else if (type == OMCGPU_FILL_LATTICE) {
      instance += gld(DM.cell_offset + gld(DM.cell_offset_ofs + c_k) + map);
      const int lat = LV(lattice, k + 1);
      LatIdx li;
      li.x = LV(lat_x, k + 1);
      li.y = LV(lat_y, k + 1);
      li.z = LV(lat_z, k + 1);
      if (lat_valid(lat, li)) {
        const int32_t* n = DM.lat_shape + 3 * lat;
        instance += gld(DM.lat_offsets + gld(DM.lat_offset_ofs + lat) +
                       gld(n) * gld(n + 1) * gld(n + 2) * map + lat_flat(lat, li));
      }
    }
  }
  return instance;
}

// Starting at level `lvl` (whose universe/r/u columns are already set), descend
// until a material cell is found.  Returns false when the particle is in no cell
// of the current universe (or left a lattice without an outer universe).
OMC_HD bool find_cell_inner(const PState& ps, int64_t i, int lvl, int surface)
{
  V3 r = get_r(ps, i, lvl);
  V3 u = get_u(ps, i, lvl);
  int univ = LV(universe, lvl);
  bool crossed = surface != SURFACE_NONE; // first level searched after a surface crossing
  for (;; ++lvl) {
    int found = C_NONE;
    if (crossed) {
      // the cell entered through a surface names that half-space in its region
      // (cells of a valid model do not overlap): search those cells of this
      // universe first -- same result as the full list, far fewer candidates
      const int slot = 2 * (abs(surface) - 1) + (surface > 0 ? 1 : 0);
      const int k1 = gld(DM.tok_cell_ofs + slot + 1);
      for (int k = gld(DM.tok_cell_ofs + slot); k < k1; ++k) {
        const int c = gld(DM.tok_cells + k);
        if (gld(DM.cell_universe + c) == univ && cell_contains(c, r, u, surface)) {
          found = c;
          break;
        }
      }
      crossed = false;
    }
    if (found == C_NONE) { // Universe::find_cell
      const int k1 = gld(DM.univ_cell_ofs + univ + 1);
      for (int k = gld(DM.univ_cell_ofs + univ); k < k1; ++k) {
        const int c = gld(DM.univ_cells + k);
        if (cell_contains(c, r, u, surface)) {
          found = c;
          break;
        }
      }
    }
    if (found == C_NONE) {
      ps.n_coord[i] = lvl + 1;
      return false;
    }
    LV(cell, lvl) = found;
    const int type = gld(DM.cell_type + found);
    if (type == OMCGPU_FILL_MATERIAL) {
      ps.n_coord[i] = lvl + 1;
      int inst = 0;
      if (gld(DM.cell_distribcell_index + found) >= 0)
        inst = cell_instance_at_level(ps, i, lvl);
      ps.cell_instance[i] = inst;
      const int m0 = gld(DM.cell_mat_ofs + found), m1 = gld(DM.cell_mat_ofs + found + 1);
      const int t0 = gld(DM.cell_sqrtkT_ofs + found), t1 = gld(DM.cell_sqrtkT_ofs + found + 1);
      ps.material_last[i] = ps.material[i];
      ps.material[i] = gld(DM.cell_material + m0 + ((m1 - m0) > 1 ? inst : 0));
      ps.sqrtkT_last[i] = ps.sqrtkT[i];
      ps.sqrtkT[i] = gld(DM.cell_sqrtkT + t0 + ((t1 - t0) > 1 ? inst : 0));
      return true;
    }
    if (lvl + 1 >= OMCGPU_MAX_COORD) {
      ps.n_coord[i] = lvl + 1;
      return false;
    }
    const double* t = DM.cell_translation + 3 * found;
    r.x -= gld(t);
    r.y -= gld(t + 1);
    r.z -= gld(t + 2);
    if (gld(DM.cell_rotated + found)) { // LocalCoord::rotate, src/geometry.cpp:200-203, 219-222
      r = cell_rotate(found, r);
      u = cell_rotate(found, u);
    }
    const int nl = lvl + 1;
    const int fill = gld(DM.cell_fill + found);
    if (type == OMCGPU_FILL_UNIVERSE) {
      univ = fill;
      // lattice column of the new level stays C_NONE (levels were reset)
    } else {
      const LatIdx li = lat_get_indices(fill, r, u);
      r = lat_local_position(fill, r, li);
      LV(lattice, nl) = fill;
      LV(lat_x, nl) = li.x;
      LV(lat_y, nl) = li.y;
      LV(lat_z, nl) = li.z;
      if (lat_valid(fill, li)) {
        univ = lat_universe(fill, li);
      } else {
        univ = gld(DM.lat_outer + fill);
        if (univ == C_NONE) {
          ps.n_coord[i] = nl;
          return false;
        }
      }
    }
    LV(universe, nl) = univ;
    set_r(ps, i, nl, r);
    set_u(ps, i, nl, u);
  }
}

OMC_HD void reset_deeper(const PState& ps, int64_t i, int n_coord)
{
  const int L = DM.n_coord_levels < OMCGPU_MAX_COORD ? DM.n_coord_levels : OMCGPU_MAX_COORD;
  for (int l = n_coord; l < L; ++l)
    reset_level(ps, i, l);
}

// (cell_rotate is defined with the lattice helpers above)
// exhaustive_find_cell / neighbor_list_find_cell with n_coord already set.
// Out of line on the device: the crossing event calls it from six places, and inlined
// copies made k_surface_xa 247 KB of SASS -- with the GPU nearly empty (the last ~100
// sweeps of a generation) a quarter of its stall cycles were instruction fetches
// (profiles/README.md, r17).
#if OMC_CUDA
__device__ __noinline__ bool find_cell_from_ool(int64_t i, int n_coord, int surface)
{
  reset_deeper(c_ps, i, n_coord);
  return find_cell_inner(c_ps, i, n_coord - 1, surface);
}
// the same for a particle staged in thread-local memory: the view is rebuilt from the staging
// area's address (so the caller's view never has to exist in memory)
template<int NL>
__device__ __noinline__ bool find_cell_from_local_ool(LocalParticle<NL>* lp, int64_t index, int n_coord, int surface)
{
  const PState ps = local_view(*lp, c_ps, index);
  reset_deeper(ps, 0, n_coord);
  return find_cell_inner(ps, 0, n_coord - 1, surface);
}
template<int NL>
OMC_HD LocalParticle<NL>* staging_area(const PState& ps)
{ // p1 is where the view's E column points
  return reinterpret_cast<LocalParticle<NL>*>(reinterpret_cast<char*>(ps.E.p) - offsetof(LocalParticle<NL>, p1));
}
#endif
OMC_HD bool find_cell_from(const PState& ps, int64_t i, int n_coord, int surface)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  switch (ps.is_local) { // a compile-time constant wherever a view is built
  case 0:
    return find_cell_from_ool(i, n_coord, surface);
  case 1:
    return find_cell_from_local_ool<0>(staging_area<0>(ps), ps.id_shift, n_coord, surface);
  case 2:
    return find_cell_from_local_ool<1>(staging_area<1>(ps), ps.id_shift, n_coord, surface);
  case 4:
    return find_cell_from_local_ool<3>(staging_area<3>(ps), ps.id_shift, n_coord, surface);
  default:
    return find_cell_from_local_ool<OMCGPU_MAX_COORD>(staging_area<OMCGPU_MAX_COORD>(ps), ps.id_shift, n_coord, surface);
  }
#else
  reset_deeper(ps, i, n_coord);
  return find_cell_inner(ps, i, n_coord - 1, surface);
#endif
}

OMC_HD bool find_cell_from_root(const PState& ps, int64_t i, int surface)
{
  LV(universe, 0) = DM.root_universe;
  return find_cell_from(ps, i, 1, surface);
}

// =============================================================================
// distance_to_boundary -- src/geometry.cpp:361-453
// =============================================================================
struct Boundary {
  double distance;
  int surface;
  int level;
  int trans;
};

OMC_HD Boundary distance_to_boundary(const PState& ps, int64_t i, int n_coord, int surface,
  bool& negative_lattice_distance)
{
  Boundary info;
  info.distance = inf_d();
  info.surface = SURFACE_NONE;
  info.level = 0;
  info.trans = TRANS_NONE;
  double d_lat = inf_d();
  int level_lat_trans = TRANS_NONE;
  negative_lattice_distance = false;
  for (int l = 0; l < n_coord; l++) {
    const V3 r = get_r(ps, i, l);
    const V3 u = get_u(ps, i, l);
    const int c = LV(cell, l);
    double d_surf;
    int level_surf_cross;
    cell_distance(c, r, u, surface, d_surf, level_surf_cross);
    const int lat = LV(lattice, l);
    if (lat != C_NONE) {
      if (lat_is_hex(lat)) { // out of line; the common path carries one uniform test
        const HexDist hd = hex_distance_at_level(ps, i, l);
        d_lat = hd.d;
        level_lat_trans = hd.trans;
      } else {
        d_lat = lat_distance(lat, r, u, level_lat_trans);
      }
      if (d_lat < 0)
        negative_lattice_distance = true;
    }
    const double d = info.distance;
    if (d_surf < d_lat - FP_COINCIDENT) {
      if (d == inf_d() || (d - d_surf) / d >= FP_REL_PRECISION) {
        info.distance = d_surf;
        if (gld(DM.cell_simple + c) || d_surf == INFTY) {
          info.surface = level_surf_cross;
        } else {
          const V3 r_hit = v3(r.x + d_surf * u.x, r.y + d_surf * u.y, r.z + d_surf * u.z);
          const V3 nrm = surf_normal(abs(level_surf_cross) - 1, r_hit);
          info.surface = (dot(u, nrm) > 0) ? abs(level_surf_cross) : -abs(level_surf_cross);
        }
        info.trans = TRANS_NONE;
        info.level = l + 1;
      }
    } else {
      if (d == inf_d() || (d - d_lat) / d >= FP_REL_PRECISION) {
        info.distance = d_lat;
        info.surface = SURFACE_NONE;
        info.trans = level_lat_trans;
        info.level = l + 1;
      }
    }
  }
  return info;
}

// =============================================================================
// Multigroup cross sections -- src/mgxs.cpp:599-675 (the (t,a) cache of the
// reference is a pure memoisation of these two index functions; the device
// recomputes them instead of carrying 48 bytes of cache per particle)
// =============================================================================
OMC_HD int mg_temperature_index(int m, double sqrtkT)
{
  const int b0 = ld(DM.mgm_blk_ofs + m);
  const int nt = ld(DM.mgm_n_temps + m);
  if (nt == 1)
    return 0;
  const double kT = sqrtkT * sqrtkT;
  int best = 0;
  double bestv = fabs(ld(DM.mgm_kTs + b0) - kT);
  for (int t = 1; t < nt; ++t) {
    const double v = fabs(ld(DM.mgm_kTs + b0 + t) - kT);
    if (v < bestv) {
      bestv = v;
      best = t;
    }
  }
  return best;
}
OMC_HD int mg_angle_index(int m, V3 u)
{
  if (ld(DM.mgm_isotropic + m))
    return 0;
  const int n_pol = ld(DM.mgm_n_pol + m), n_azi = ld(DM.mgm_n_azi + m);
  const double my_pol = acos(u.z);
  const double my_azi = atan2(u.y, u.x);
  double delta_angle = PI / n_pol;
  const int pp = (int)floor(my_pol / delta_angle);
  delta_angle = 2. * PI / n_azi;
  const int aa = (int)floor((my_azi + PI) / delta_angle);
  return n_azi * pp + aa;
}
OMC_HD int mg_ba(int m, int t, int a)
{
  return ld(DM.blk_ang_ofs + ld(DM.mgm_blk_ofs + m) + t) + a;
}

// src/mgxs.cpp:537-586
OMC_HD void mg_sample_fission_energy(int ba, int gin, int& dg, int& gout, uint64_t& seed)
{
  const int G = DM.mg_n_groups, DG = DM.mg_n_delayed;
  const double nu_fission = ld(DM.mg_nu_fission + ba * G + gin);
  double prob_prompt = ld(DM.mg_prompt_nu_fission + ba * G + gin);
  const double xi_pd = prn(seed) * nu_fission;
  const double xi_gout = prn(seed);
  if (xi_pd <= prob_prompt) {
    dg = -1;
    double prob_gout = 0.;
    const double* chi = DM.mg_chi_prompt + ((int64_t)ba * G + gin) * G;
    for (gout = 0; gout < G; ++gout) {
      prob_gout += ld(chi + gout);
      if (xi_gout < prob_gout)
        break;
    }
  } else {
    for (dg = 0; dg < DG; ++dg) {
      prob_prompt += ld(DM.mg_delayed_nu_fission + ((int64_t)ba * DG + dg) * G + gin);
      if (xi_pd < prob_prompt)
        break;
    }
    dg = dg < DG - 1 ? dg : DG - 1;
    double prob_gout = 0.;
    const double* chi = DM.mg_chi_delayed + (((int64_t)ba * DG + dg) * G + gin) * G;
    for (gout = 0; gout < G; ++gout) {
      prob_gout += ld(chi + gout);
      if (xi_gout < prob_gout)
        break;
    }
  }
}

// src/scattdata.cpp:158-170, 341-367, 528-558, 729-775
OMC_HD void mg_sample_scatter(int ba, int gin, int& gout, double& mu, double& wgt, uint64_t& seed)
{
  const int G = DM.mg_n_groups;
  const int gmin = ld(DM.sc_gmin + ba * G + gin), gmax = ld(DM.sc_gmax + ba * G + gin);
  const int e0 = ld(DM.sc_gofs + ba * G + gin);
  const int order = ld(DM.sc_order + ba);
  const double xi = prn(seed);
  double prob = 0.;
  int i_gout = 0;
  for (gout = gmin; gout < gmax; ++gout) {
    prob += ld(DM.sc_energy + e0 + i_gout);
    if (xi < prob)
      break;
    ++i_gout;
  }
  const int64_t d0 = (int64_t)ld(DM.sc_dofs + ba * G + gin) + (int64_t)i_gout * order;
  const double* dist = DM.sc_dist + d0;
  const int fmt = ld(DM.sc_format + ba);
  if (fmt == OMCGPU_SCATT_LEGENDRE) {
    const double Mx = ld(DM.sc_maxval + e0 + i_gout);
    for (int samples = 0; samples < MAX_SAMPLE; ++samples) {
      mu = 2. * prn(seed) - 1.;
      const double f = evaluate_legendre(order - 1, dist, mu);
      if (f > 0.) {
        const double uu = prn(seed) * Mx;
        if (uu <= f)
          break;
      }
    }
  } else if (fmt == OMCGPU_SCATT_HISTOGRAM) {
    const double* mugrid = DM.sc_mu + ld(DM.sc_mu_ofs + ba);
    const double xi2 = prn(seed);
    int imu;
    if (xi2 < ld(dist)) {
      imu = 0;
    } else { // std::upper_bound
      int lo = 0, len = order;
      while (len > 0) {
        int half = len >> 1;
        if (!(xi2 < ld(dist + lo + half))) {
          lo += half + 1;
          len -= half + 1;
        } else {
          len = half;
        }
      }
      imu = lo;
    }
    mu = prn(seed) * ld(DM.sc_dmu + ba) + ld(mugrid + imu);
    if (mu < -1.)
      mu = -1.;
    else if (mu > 1.)
      mu = 1.;
  } else {
    const double* fmu = DM.sc_fmu + d0;
    const double* mugrid = DM.sc_mu + ld(DM.sc_mu_ofs + ba);
    const int NP = order;
    const double xi2 = prn(seed);
    double c_k = ld(dist);
    int k;
    for (k = 0; k < NP - 1; k++) {
      const double c_k1 = ld(dist + k + 1);
      if (xi2 < c_k1)
        break;
      c_k = c_k1;
    }
    k = k < NP - 2 ? k : NP - 2;
    const double p0 = ld(fmu + k), mu0 = ld(mugrid + k);
    const double p1 = ld(fmu + k + 1), mu1 = ld(mugrid + k + 1);
    if (p0 == p1) {
      mu = mu0 + (xi2 - c_k) / p0;
    } else {
      const double frac = (p1 - p0) / (mu1 - mu0);
      mu = mu0 + (sqrt(fmax(0., p0 * p0 + 2. * frac * (xi2 - c_k))) - p0) / frac;
    }
    if (mu < -1.)
      mu = -1.;
    else if (mu > 1.)
      mu = 1.;
  }
  wgt *= ld(DM.sc_mult + e0 + i_gout);
}

// ScattData::get_xs(SCATTER, gin) -- src/scattdata.cpp:174-225
OMC_HD double mg_scatter_xs(int ba, int g)
{
  const int G = DM.mg_n_groups;
  const int e0 = ld(DM.sc_gofs + ba * G + g);
  const int n = ld(DM.sc_gmax + ba * G + g) - ld(DM.sc_gmin + ba * G + g) + 1;
  double ip = 0.0;
  for (int k = 0; k < n; ++k)
    ip = ip + ld(DM.sc_mult + e0 + k) * ld(DM.sc_energy + e0 + k);
  return ld(DM.sc_scattxs + ba * G + g) / ip;
}

// =============================================================================
// Continuous-energy cross sections -- src/material.cpp:829-901,
// src/nuclide.cpp:616-775, src/reaction.cpp:98-112
// =============================================================================
struct MicroXS {
  double total, absorption, fission, nu_fission;
  int index_grid;
  double interp_factor;
};

// Hash bin of the union log grid: int(log(E/E_min)/spacing), src/material.cpp:833-834.
// The device log() may differ from the host libm in the last ulp, which could
// move an energy that sits on a bin edge into the neighbouring bin.  hash_edge[k]
// is the smallest double the HOST libm maps to bin >= k (built once at init),
// so one or two comparisons make the bin bit-identical to the reference's.
OMC_HD int ce_log_index(double E)
{
  int k = (int)(log(E / DM.energy_min) / DM.log_spacing);
  const int nb = DM.n_log_bins;
  k = k < 0 ? 0 : (k > nb - 1 ? nb - 1 : k);
  while (k > 0 && E < ld(DM.hash_edge + k))
    --k;
  while (k < nb - 1 && E >= ld(DM.hash_edge + k + 1))
    ++k;
  return k; // clamped to the last valid window, see oracle.cpp ce_log_index
}

// Fine hash (derived by omcgpu_init): DM.n_fine_bins = FINE_PER_LOG_BIN x n_log_bins
// equal-lethargy bins; fine_window[j][nuc] = {lo, hi} (bin-major: the nuclides of a
// material share the sectors of bin j) brackets, for every energy the
// device may map to bin j, K = the number of grid points below E.  The windows are
// built from bin edges widened by 1e-6 of a bin, so the bin does not have to be
// bit-exact (the device log() may differ from libm in the last ulp) and holds ~1
// point: the ~4 dependent probes of the search inside the reference's 8000-bin
// window shrink to 0-2.  The grid index is still the reference's: for E strictly
// between two points every bracketing window yields i_grid = K - 1; an E that sits
// ON a grid point, below the first or above the last point takes the reference's
// own window (ce_grid_index_reference), because lower_bound_index
// (include/openmc/search.h:14-20) then depends on where that window starts.
OMC_HD int ce_fine_index(double E)
{
  const int nf = DM.n_fine_bins;
  int j = (int)(log(E / DM.energy_min) * DM.fine_inv_spacing);
  return j < 0 ? 0 : (j > nf - 1 ? nf - 1 : j);
}

// Nuclide::calculate_xs, src/nuclide.cpp:725-742, with the reference's hash window
OMC_OOL int ce_grid_index_reference(int nuc, const double* energy, int n_grid, double E)
{
  int i_grid;
  if (E < ld(energy)) {
    i_grid = 0;
  } else if (E > ld(energy + n_grid - 1)) {
    i_grid = n_grid - 2;
  } else {
    const int i_log = ce_log_index(E);
    const int32_t* gi = DM.nuc_grid_index + (int64_t)nuc * (DM.n_log_bins + 1) + i_log;
    const int i_low = ld(gi);
    const int i_high = ld(gi + 1) + 1;
    i_grid = i_low + (int)lower_bound_index(energy + i_low, (int64_t)(i_high - i_low), E);
  }
  if (ld(energy + i_grid) == ld(energy + i_grid + 1))
    ++i_grid;
  return i_grid;
}

// grid index + the bracketing energies
OMC_HD int ce_grid_index(int nuc, int j_fine, const double* energy, int n_grid, double E, double& e0,
  double& e1)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  const int2 w = __ldg(reinterpret_cast<const int2*>(DM.fine_window) + (int64_t)j_fine * DM.n_nuclides + nuc);
  int lo = w.x, len = w.y - w.x;
#else
  const int32_t* w = DM.fine_window + 2 * ((int64_t)j_fine * DM.n_nuclides + nuc);
  int lo = w[0], len = w[1] - w[0];
#endif
  while (len > 0) { // first index in [lo, hi] whose energy is >= E
    const int half = len >> 1;
    if (ld(energy + lo + half) < E) {
      lo += half + 1;
      len -= half + 1;
    } else {
      len = half;
    }
  }
  if (lo >= 1 && lo <= n_grid - 1) {
    e1 = ld(energy + lo);
    if (e1 != E) {
      e0 = ld(energy + lo - 1);
      return lo - 1;
    }
  }
  const int i_grid = ce_grid_index_reference(nuc, energy, n_grid, E);
  e0 = ld(energy + i_grid);
  e1 = ld(energy + i_grid + 1);
  return i_grid;
}

OMC_HD void ce_micro_xs(int nuc, int j_fine, double E, MicroXS& micro)
{
  const int64_t g0 = ld(DM.nuc_grid_ofs + nuc);
  const int n_grid = (int)(ld(DM.nuc_grid_ofs + nuc + 1) - g0);
  double e0, e1;
  const int i_grid = ce_grid_index(nuc, j_fine, DM.nuc_energy + g0, n_grid, E, e0, e1);
  const double f = (E - e0) / (e1 - e0);
  micro.index_grid = i_grid;
  micro.interp_factor = f;
#if OMC_CUDA && defined(__CUDA_ARCH__)
  // two adjacent 32-byte rows: exactly two sectors
  const double2* row = reinterpret_cast<const double2*>(DM.nuc_xs) + 2 * (g0 + i_grid);
  const double2 a01 = __ldg(row), a23 = __ldg(row + 1);
  const double2 b01 = __ldg(row + 2), b23 = __ldg(row + 3);
  const double a0 = a01.x, a1 = a01.y, a2 = a23.x, a3 = a23.y;
  const double b0 = b01.x, b1 = b01.y, b2 = b23.x, b3 = b23.y;
#else
  const double* row = DM.nuc_xs + 4 * (g0 + i_grid);
  const double a0 = row[0], a1 = row[1], a2 = row[2], a3 = row[3];
  const double b0 = row[4], b1 = row[5], b2 = row[6], b3 = row[7];
#endif
  micro.total = (1.0 - f) * a0 + f * b0;
  micro.absorption = (1.0 - f) * a1 + f * b1;
  if (ld(DM.nuc_fissionable + nuc)) {
    micro.fission = (1.0 - f) * a2 + f * b2;
    micro.nu_fission = (1.0 - f) * a3 + f * b3;
  } else {
    micro.fission = 0.0;
    micro.nu_fission = 0.0;
  }
}

OMC_HD double ce_reaction_xs(int rx, const MicroXS& micro)
{
  const int thr = ld(DM.rx_threshold + rx);
  if (micro.index_grid < thr)
    return 0.0;
  const double* v = DM.rx_xs + ld(DM.rx_xs_ofs + rx) + (micro.index_grid - thr);
  return (1.0 - micro.interp_factor) * ld(v) + micro.interp_factor * ld(v + 1);
}

// Several temperatures per nuclide (TemperatureMethod::NEAREST, src/nuclide.cpp:682-691): every
// nuc_* / rx_* array is indexed by TABLE = nuclide at one loaded temperature, and omcgpu_init
// resolves "the table nearest to the cell's kT" once per (material, cell temperature) pair into
// derived materials whose entries are table indices (engine.cu, expand_temperatures) -- the
// lookup and collision code below never sees a temperature.
// Materials with many nuclides (depleted fuel, config C5: ~300): NB nuclides walk the
// lookup chain (window -> probes -> bracketing energies -> two xs rows) in lockstep, so
// that the loads of one stage are independent and overlap; the partial sums are added
// in nuclide order afterwards, i.e. every sum sees the same operands in the same order
// as the one-at-a-time walk (bit-identical results).
#ifndef OMC_XS_BATCH
#define OMC_XS_BATCH 4
#endif
#ifndef OMC_XS_BATCH_MIN
#define OMC_XS_BATCH_MIN 16 /* nuclides in the material from which the lockstep walk is used */
#endif
template<int NB>
OMC_HD void ce_macro_xs_batch(int k0, int k1, int j_fine, double E, double& t, double& a, double& f,
  double& nf, int32_t* i_grid_out, int max_nuc, double* prefix_out)
{
  static_assert(XS_PREFIX_STRIDE % NB == 0, "a prefix boundary must fall between two batches");
  for (int kb = k0; kb < k1; kb += NB) {
    int nuc[NB], n_grid[NB], lo[NB], len[NB], ig[NB];
    const double* energy[NB];
    int64_t g0[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j)
      nuc[j] = ld(DM.mat_nuclide + (kb + j < k1 ? kb + j : k1 - 1));
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      g0[j] = ld(DM.nuc_grid_ofs + nuc[j]);
      n_grid[j] = (int)(ld(DM.nuc_grid_ofs + nuc[j] + 1) - g0[j]);
      energy[j] = DM.nuc_energy + g0[j];
#if OMC_CUDA && defined(__CUDA_ARCH__)
      const int2 w = __ldg(reinterpret_cast<const int2*>(DM.fine_window) + (int64_t)j_fine * DM.n_nuclides + nuc[j]);
      lo[j] = w.x;
      len[j] = w.y - w.x;
#else
      const int32_t* w = DM.fine_window + 2 * ((int64_t)j_fine * DM.n_nuclides + nuc[j]);
      lo[j] = w[0];
      len[j] = w[1] - w[0];
#endif
    }
    bool any = true;
    while (any) { // lower bound of E inside every window, in step
      any = false;
      double probe[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (len[j] > 0)
          probe[j] = ld(energy[j] + lo[j] + (len[j] >> 1));
#pragma unroll
      for (int j = 0; j < NB; ++j)
        if (len[j] > 0) {
          const int half = len[j] >> 1;
          if (probe[j] < E) {
            lo[j] += half + 1;
            len[j] -= half + 1;
          } else {
            len[j] = half;
          }
          any = any || len[j] > 0;
        }
    }
    double e0[NB], e1[NB];
    bool fast[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      fast[j] = lo[j] >= 1 && lo[j] <= n_grid[j] - 1;
      const int k = fast[j] ? lo[j] : 1;
      e1[j] = ld(energy[j] + k);
      e0[j] = ld(energy[j] + k - 1);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (fast[j] && e1[j] != E) {
        ig[j] = lo[j] - 1;
      } else { // on a grid point or outside the grid: the reference's own window decides
        ig[j] = ce_grid_index_reference(nuc[j], energy[j], n_grid[j], E);
        e0[j] = ld(energy[j] + ig[j]);
        e1[j] = ld(energy[j] + ig[j] + 1);
      }
    }
    double ra[NB][4], rb[NB][4], rho[NB];
    int fis[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j) {
#if OMC_CUDA && defined(__CUDA_ARCH__)
      const double2* row = reinterpret_cast<const double2*>(DM.nuc_xs) + 2 * (g0[j] + ig[j]);
      const double2 a01 = __ldg(row), a23 = __ldg(row + 1), b01 = __ldg(row + 2), b23 = __ldg(row + 3);
      ra[j][0] = a01.x; ra[j][1] = a01.y; ra[j][2] = a23.x; ra[j][3] = a23.y;
      rb[j][0] = b01.x; rb[j][1] = b01.y; rb[j][2] = b23.x; rb[j][3] = b23.y;
#else
      const double* row = DM.nuc_xs + 4 * (g0[j] + ig[j]);
      for (int c = 0; c < 4; ++c) {
        ra[j][c] = row[c];
        rb[j][c] = row[4 + c];
      }
#endif
      rho[j] = ld(DM.mat_atom_density + (kb + j < k1 ? kb + j : k1 - 1));
      fis[j] = ld(DM.nuc_fissionable + nuc[j]);
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      if (kb + j >= k1)
        continue;
      const double fr = (E - e0[j]) / (e1[j] - e0[j]);
      t += rho[j] * ((1.0 - fr) * ra[j][0] + fr * rb[j][0]);
      a += rho[j] * ((1.0 - fr) * ra[j][1] + fr * rb[j][1]);
      f += rho[j] * (fis[j] ? (1.0 - fr) * ra[j][2] + fr * rb[j][2] : 0.0);
      nf += rho[j] * (fis[j] ? (1.0 - fr) * ra[j][3] + fr * rb[j][3] : 0.0);
      if (i_grid_out && (kb + j - k0) < max_nuc)
        i_grid_out[kb + j - k0] = ig[j];
    }
    // running total after nuclides 0 .. 32 b + 31 of the material: exactly the partial sum the
    // sampling walk of sample_nuclide (src/physics.cpp:494-518) holds at that point
    if (prefix_out && (kb + NB - k0) % XS_PREFIX_STRIDE == 0)
      prefix_out[(kb + NB - k0) / XS_PREFIX_STRIDE - 1] = t;
  }
}

// Material::calculate_neutron_xs (src/material.cpp:829-901): out = total, absorption,
// fission, nu-fission, accumulated in nuclide order.  Few nuclides (C2/C3: 1-5 per
// material): one at a time -- measured faster there than the lockstep walk, which
// costs registers (profiles/README.md, step r07).
// Total cross sections of NB nuclides of a material (entries kb .. kb+NB-1, clamped to
// k1-1) at one energy, walked in lockstep -- the sampling walk of sample_nuclide
// (src/physics.cpp:494-518) over a many-nuclide material needs up to n of them in a row.
// sig_t[j] is bit-identical to ce_micro_xs(...).total.
template<int NB>
OMC_HD void ce_total_xs_batch(int kb, int k1, int j_fine, double E, double* sig_t)
{
  int nuc[NB], n_grid[NB], lo[NB], len[NB], ig[NB];
  const double* energy[NB];
  int64_t g0[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j)
    nuc[j] = ld(DM.mat_nuclide + (kb + j < k1 ? kb + j : k1 - 1));
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    g0[j] = ld(DM.nuc_grid_ofs + nuc[j]);
    n_grid[j] = (int)(ld(DM.nuc_grid_ofs + nuc[j] + 1) - g0[j]);
    energy[j] = DM.nuc_energy + g0[j];
#if OMC_CUDA && defined(__CUDA_ARCH__)
    const int2 w = __ldg(reinterpret_cast<const int2*>(DM.fine_window) + (int64_t)j_fine * DM.n_nuclides + nuc[j]);
    lo[j] = w.x;
    len[j] = w.y - w.x;
#else
    const int32_t* w = DM.fine_window + 2 * ((int64_t)j_fine * DM.n_nuclides + nuc[j]);
    lo[j] = w[0];
    len[j] = w[1] - w[0];
#endif
  }
  bool any = true;
  while (any) {
    any = false;
    double probe[NB];
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (len[j] > 0)
        probe[j] = ld(energy[j] + lo[j] + (len[j] >> 1));
#pragma unroll
    for (int j = 0; j < NB; ++j)
      if (len[j] > 0) {
        const int half = len[j] >> 1;
        if (probe[j] < E) {
          lo[j] += half + 1;
          len[j] -= half + 1;
        } else {
          len[j] = half;
        }
        any = any || len[j] > 0;
      }
  }
  double e0[NB], e1[NB];
  bool fast[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    fast[j] = lo[j] >= 1 && lo[j] <= n_grid[j] - 1;
    const int k = fast[j] ? lo[j] : 1;
    e1[j] = ld(energy[j] + k);
    e0[j] = ld(energy[j] + k - 1);
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    if (fast[j] && e1[j] != E) {
      ig[j] = lo[j] - 1;
    } else {
      ig[j] = ce_grid_index_reference(nuc[j], energy[j], n_grid[j], E);
      e0[j] = ld(energy[j] + ig[j]);
      e1[j] = ld(energy[j] + ig[j] + 1);
    }
  }
  double a0[NB], b0[NB];
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const double* row = DM.nuc_xs + 4 * (g0[j] + ig[j]);
    a0[j] = ld(row);
    b0[j] = ld(row + 4);
  }
#pragma unroll
  for (int j = 0; j < NB; ++j) {
    const double fr = (E - e0[j]) / (e1[j] - e0[j]);
    sig_t[j] = (1.0 - fr) * a0[j] + fr * b0[j];
  }
}

// BX: the kernel was compiled for models with many-nuclide materials (its register
// budget covers the lockstep walk); the host picks the kernel flavour per model.
template<bool BX>
OMC_HD void ce_macro_xs(int m, double E, double& t, double& a, double& f, double& nf,
  int32_t* i_grid_out, int max_nuc, int batch_min = OMC_XS_BATCH_MIN, double* prefix_out = nullptr)
{
  t = a = f = nf = 0.0;
  const int j_fine = ce_fine_index(E);
  const int k0 = ld(DM.mat_nuc_ofs + m), k1 = ld(DM.mat_nuc_ofs + m + 1);
  if (BX && k1 - k0 >= batch_min) {
    // (a material walked one nuclide at a time has fewer than OMC_XS_BATCH_MIN <= XS_PREFIX_STRIDE
    // nuclides when prefixes are kept: it has no prefix entry and its collisions walk from 0)
    ce_macro_xs_batch<OMC_XS_BATCH>(k0, k1, j_fine, E, t, a, f, nf, i_grid_out, max_nuc,
      k1 - k0 >= 2 * XS_PREFIX_STRIDE ? prefix_out : nullptr);
    return;
  }
  for (int k = k0; k < k1; ++k) {
    MicroXS micro;
    ce_micro_xs(ld(DM.mat_nuclide + k), j_fine, E, micro);
    const double rho = ld(DM.mat_atom_density + k);
    t += rho * micro.total;
    a += rho * micro.absorption;
    f += rho * micro.fission;
    nf += rho * micro.nu_fission;
    if (i_grid_out && (k - k0) < max_nuc)
      i_grid_out[k - k0] = micro.index_grid;
  }
}

// src/endf.cpp:125-134 (Polynomial), 168-226 (Tabulated1D, lin-lin)
OMC_OOL double eval_function(int fn, double x)
{
  const int o0 = ld(DM.fn_ofs + fn), o1 = ld(DM.fn_ofs + fn + 1);
  const double* d = DM.fn_data + o0;
  const int n = o1 - o0;
  if (ld(DM.fn_type + fn) == OMCGPU_FN_POLYNOMIAL) {
    double y = 0.0;
    for (int k = n - 1; k >= 0; --k)
      y = y * x + ld(d + k);
    return y;
  }
  const int np = n / 2;
  const double* ys = d + np;
  if (x < ld(d))
    return ld(ys);
  if (x > ld(d + np - 1))
    return ld(ys + np - 1);
  const int k = (int)lower_bound_index(d, (int64_t)np, x);
  const double x0 = ld(d + k), x1 = ld(d + k + 1), y0 = ld(ys + k), y1 = ld(ys + k + 1);
  const double r = (x - x0) / (x1 - x0);
  return y0 + r * (y1 - y0);
}

// Nuclide::nu, src/nuclide.cpp:535-576 ; mode 0 prompt, 1 delayed, 2 total
OMC_OOL double ce_nu(int nuc, double E, int mode)
{
  if (!ld(DM.nuc_fissionable + nuc))
    return 0.0;
  const int rx = ld(DM.nuc_rx_ofs + nuc) + ld(DM.nuc_fission_rx + nuc);
  const int p0 = ld(DM.rx_prod_ofs + rx), p1 = ld(DM.rx_prod_ofs + rx + 1);
  if (mode == 0)
    return eval_function(ld(DM.prod_yield_fn + p0), E);
  if (mode == 1) {
    if (ld(DM.nuc_n_precursor + nuc) > 0 && DM.create_delayed_neutrons) {
      double nu = 0.0;
      for (int k = p0 + 1; k < p1; ++k)
        if (ld(DM.prod_emission + k) == 1)
          nu += eval_function(ld(DM.prod_yield_fn + k), E);
      return nu;
    }
    return 0.0;
  }
  const int tf = ld(DM.nuc_total_nu_fn + nuc);
  if (tf >= 0 && DM.create_delayed_neutrons)
    return eval_function(tf, E);
  return eval_function(ld(DM.prod_yield_fn + p0), E);
}

// ReactionProduct::sample -> UncorrelatedAngleEnergy::sample (isotropic angle),
// src/reaction_product.cpp:109-130, src/secondary_uncorrelated.cpp:53-66,
// src/distribution_energy.cpp:49-52, 284-298, 309-327, 347-361
OMC_OOL void ce_sample_product(int prod, double E_in, double& E_out, double& mu, uint64_t& seed)
{
  mu = -1. + (1. - (-1.)) * prn(seed);
  const double par0 = ld(DM.prod_par + 2 * prod), par1 = ld(DM.prod_par + 2 * prod + 1);
  const int f0 = ld(DM.prod_fn + 2 * prod), f1 = ld(DM.prod_fn + 2 * prod + 1);
  switch (ld(DM.prod_elaw + prod)) {
  case OMCGPU_ELAW_LEVEL:
    E_out = par1 * (E_in - par0);
    break;
  case OMCGPU_ELAW_MAXWELL: {
    const double theta = eval_function(f0, E_in);
    for (int k = 0; k < MAX_SAMPLE; ++k) {
      E_out = maxwell_spectrum(theta, seed);
      if (E_out <= E_in - par0)
        break;
    }
    break;
  }
  case OMCGPU_ELAW_EVAPORATION: {
    const double theta = eval_function(f0, E_in);
    const double y = (E_in - par0) / theta;
    const double v = 1.0 - exp(-y);
    double x = 0.0;
    for (int k = 0; k < MAX_SAMPLE; ++k) {
      const double a1 = 1.0 - v * prn(seed);
      const double a2 = 1.0 - v * prn(seed);
      x = -log(a1 * a2);
      if (x <= y)
        break;
    }
    E_out = x * theta;
    break;
  }
  case OMCGPU_ELAW_WATT: {
    const double a = eval_function(f0, E_in);
    const double b = eval_function(f1, E_in);
    for (int k = 0; k < MAX_SAMPLE; ++k) {
      E_out = watt_spectrum(a, b, seed);
      if (E_out <= E_in - par0)
        break;
    }
    break;
  }
  default:
    E_out = E_in;
  }
}

// src/physics.cpp:990-1041
OMC_HD V3 sample_cxs_target_velocity(double awr, double E, V3 u, double kT, uint64_t& seed)
{
  const double beta_vn = sqrt(awr * E / kT);
  const double alpha = 1.0 / (1.0 + sqrt(PI) * beta_vn / 2.0);
  double beta_vt_sq = 0.0, mu = 0.0;
  for (int it = 0; it < MAX_SAMPLE; ++it) {
    const double r1 = prn(seed);
    const double r2 = prn(seed);
    if (prn(seed) < alpha) {
      beta_vt_sq = -log(r1 * r2);
    } else {
      const double c = cos(PI / 2.0 * prn(seed));
      beta_vt_sq = -log(r1) - log(r2) * c * c;
    }
    const double beta_vt = sqrt(beta_vt_sq);
    mu = -1. + (1. - (-1.)) * prn(seed);
    const double accept_prob =
      sqrt(beta_vn * beta_vn + beta_vt_sq - 2 * beta_vn * beta_vt * mu) / (beta_vn + beta_vt);
    if (prn(seed) < accept_prob)
      break;
  }
  const double vt = sqrt(beta_vt_sq * kT / awr);
  const V3 d = rotate_angle(u, mu, seed);
  return v3(vt * d.x, vt * d.y, vt * d.z);
}

// =============================================================================
// Fission bank append -- reserves `nu` consecutive slots with one atomic
// (SharedArray::thread_safe_append semantics: sites that do not fit are dropped
// and the progeny counter is not advanced for them)
// =============================================================================
OMC_HD int32_t atomic_add_i32(int32_t* p, int32_t v)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  return atomicAdd(p, v);
#else
  int32_t old = *p;
  *p += v;
  return old;
#endif
}
OMC_HD void atomic_add_u64(unsigned long long* p, unsigned long long v)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  atomicAdd(p, v);
#else
  *p += v;
#endif
}

OMC_HD void bank_store(const Bank& fb, int64_t slot, V3 r, V3 u, double E, double time, double wgt,
  int dg, int64_t parent_id, int64_t progeny_id)
{
  fb.x[slot] = r.x;
  fb.y[slot] = r.y;
  fb.z[slot] = r.z;
  fb.u[slot] = u.x;
  fb.v[slot] = u.y;
  fb.w[slot] = u.z;
  fb.E[slot] = E;
  fb.time[slot] = time;
  fb.wgt[slot] = wgt;
  fb.delayed_group[slot] = dg;
  fb.parent_id[slot] = parent_id;
  fb.progeny_id[slot] = progeny_id;
}

// The shared fission bank is full (the reference's warning "Additional fission sites created in
// this generation will not be banked"): counted, reported in omcgpu_gen_out.n_fission_dropped.
OMC_HD void fission_bank_full(const GenCtx& gc)
{
  atomic_add_i32(gc.n_bank_full, 1);
}

// =============================================================================
// Collision: multigroup -- src/physics_mg.cpp:26-259
// =============================================================================
struct Collider { // registers of the colliding particle
  V3 r, u;
  double E, wgt, time, k_abs;
  int g, material;
  uint64_t seed;
  int n_progeny;
  int64_t id;
  double xs_total, xs_absorption, xs_nu_fission;
};

OMC_HD void russian_roulette(Collider& p, double weight_survive)
{ // src/physics_common.cpp:12-19
  if (weight_survive * prn(p.seed) < p.wgt)
    p.wgt = weight_survive;
  else
    p.wgt = 0.;
}

// shared by MG and CE: number of sites and slot reservation
OMC_HD int fission_site_count(Collider& p, double nu_t)
{
  int nu = (int)nu_t;
  if (prn(p.seed) <= (nu_t - nu))
    ++nu;
  return nu;
}

// collision record of the particle the Collider came from (its index in the generation)
OMC_HD int64_t an_index(const Collider& p, const GenCtx& gc)
{
  return p.id - gc.id_offset - 1;
}
OMC_HD bool an_on(const GenCtx& gc) // records are written only while tallies are active
{
#if OMC_AFTER_COLLISION_TALLIES
  return gc.an_d != nullptr && gc.tallies_active;
#else
  (void)gc;
  return false;
#endif
}
OMC_HD void an_put(const GenCtx& gc, int64_t idx, int k, double v)
{
  gc.an_d[k * gc.an_cap + idx] = v;
}
OMC_HD void an_puti(const GenCtx& gc, int64_t idx, int k, int32_t v)
{
  gc.an_i[k * gc.an_cap + idx] = v;
}

// Particle::create_secondary, src/particle.cpp:75-93: the neutron as it leaves the reaction, put on
// the particle's list; it is transported when the particle itself is dead (ev_revive)
OMC_HD void create_secondary(const Collider& p, const GenCtx& gc)
{
  if (p.E < DM.energy_cutoff)
    return;
  const int32_t slot = atomic_add_i32(gc.sec_count, 1);
  if (slot >= gc.sec_cap) { // the host fails the generation: nothing is dropped silently
    atomic_add_i32(gc.sec_count + 1, 1);
    return;
  }
  double* s = gc.sec_site + (int64_t)slot * SEC_ND;
  s[0] = p.r.x;
  s[1] = p.r.y;
  s[2] = p.r.z;
  s[3] = p.u.x;
  s[4] = p.u.y;
  s[5] = p.u.z;
  s[6] = p.E;
  s[7] = p.wgt;
  s[8] = p.time;
  const int64_t idx = an_index(p, gc);
  gc.sec_next[slot] = gc.sec_head[idx];
  gc.sec_head[idx] = slot;
}

OMC_HD void collision_mg(Collider& p, int ba, const GenCtx& gc, const Bank& fb)
{
  const int G = DM.mg_n_groups;
  if (ld(DM.mat_fissionable + p.material)) { // create_fission_sites :98-237
    const double nu_t = p.wgt / gc.keff * 1.0 * p.xs_nu_fission / p.xs_total;
    const int nu = fission_site_count(p, nu_t);
    if (nu > 0) {
      const int32_t base = atomic_add_i32(gc.n_fission, nu);
      int k = 0;
      for (; k < nu; ++k) {
        const double mu = 2. * prn(p.seed) - 1.;
        const double phi = 2. * PI * prn(p.seed);
        const V3 su = v3(mu, sqrt(1. - mu * mu) * cos(phi), sqrt(1. - mu * mu) * sin(phi));
        int dg, gout;
        mg_sample_fission_energy(ba, p.g, dg, gout, p.seed);
        double t = p.time;
        if (dg != -1) {
          const double decay_rate = ld(DM.mg_decay_rate + mg_ba(p.material, 0, 0) * DM.mg_n_delayed + dg);
          t -= log(prn(p.seed)) / decay_rate;
        }
        const int64_t slot = (int64_t)base + k;
        if (slot >= gc.fission_cap) { // bank full: the reference warns and leaves the loop
          fission_bank_full(gc);      // (thread_safe_append == -1, src/physics_mg.cpp:186-200)
          break;
        }
        bank_store(fb, slot, p.r, su, (double)gout, t, 1.0, dg + 1, p.id, p.n_progeny);
        p.n_progeny++;
      }
      if (an_on(gc) && k > 0) { // p.fission(), p.wgt_bank() = sites stored / weight, src/physics_mg.cpp:223-233
        an_puti(gc, an_index(p, gc), AN_FISSION, 1);
        an_puti(gc, an_index(p, gc), AN_FIS_BASE, base); // p.nu_bank(): the k sites from this slot on
        an_put(gc, an_index(p, gc), AN_WGT_BANK, (double)k);
      }
    }
  }
  if (p.xs_absorption > 0.) { // absorption :239-259
    if (DM.survival_biasing) {
      const double wgt_absorb = p.wgt * p.xs_absorption / p.xs_total;
      p.wgt -= wgt_absorb;
      p.k_abs += wgt_absorb * p.xs_nu_fission / p.xs_absorption;
    } else if (p.xs_absorption > prn(p.seed) * p.xs_total) {
      p.k_abs += p.wgt * p.xs_nu_fission / p.xs_absorption;
      p.wgt = 0.0;
    }
  }
  if (p.wgt == 0.0)
    return;
  if (an_on(gc))
    an_puti(gc, an_index(p, gc), AN_SCATTER, 1); // p.event() = TallyEvent::SCATTER, :95
  int gout;
  double mu;
  mg_sample_scatter(ba, p.g, gout, mu, p.wgt, p.seed); // scatter :83-96
  p.g = gout;
  p.u = rotate_angle(p.u, mu, p.seed);
  p.E = ld(DM.mg_energy_bin_avg + gout);
  (void)G;
  if (DM.survival_biasing && p.wgt < DM.weight_cutoff)
    russian_roulette(p, DM.weight_survive);
}

// =============================================================================
// Collision: continuous energy -- src/physics.cpp:93-285, 494-518, 638-818,
// 1043-1167.  The reference reads the per-particle NuclideMicroXS cache that
// calculate_xs filled at this energy (176 B x n_nuclides per particle); the
// device recomputes the few entries it needs -- same numbers, no cache.
// =============================================================================
template<bool BX = false>
OMC_HD void collision_ce(Collider& p, const GenCtx& gc, const Bank& fb, bool& no_nuclide,
  const double* prefix = nullptr)
{
  no_nuclide = false;
  const int m = p.material;
  const int i_log = ce_fine_index(p.E); // fine-hash bin handed to ce_micro_xs
  const double cutoff = prn(p.seed) * p.xs_total; // sample_nuclide :494-518
  double prob = 0.0;
  int nuc = -1;
  MicroXS micro;
  const int k0 = ld(DM.mat_nuc_ofs + m), k1 = ld(DM.mat_nuc_ofs + m + 1);
  if (BX && k1 - k0 >= OMC_XS_BATCH_MIN) {
    // many-nuclide material: the totals come in lockstep batches, the walk itself keeps
    // the reference's order (same partial sums, same nuclide chosen)
    int kb0 = k0;
    if (prefix && k1 - k0 >= 2 * XS_PREFIX_STRIDE) {
      // the macro lookup left the walk's partial sum after every 32 nuclides: skip the blocks
      // whose partial sum is still below the cutoff (same comparison as the walk's own test, so
      // the nuclide found is the one the walk from 0 finds)
      const int nb = (k1 - k0) / XS_PREFIX_STRIDE;
      int b = 0;
      while (b < nb && !(prefix[b] >= cutoff))
        ++b;
      if (b > 0) {
        prob = prefix[b - 1];
        kb0 = k0 + b * XS_PREFIX_STRIDE;
      }
    }
    for (int kb = kb0; kb < k1 && nuc < 0; kb += OMC_XS_BATCH) {
      double sig_t[OMC_XS_BATCH];
      ce_total_xs_batch<OMC_XS_BATCH>(kb, k1, i_log, p.E, sig_t);
#pragma unroll
      for (int j = 0; j < OMC_XS_BATCH; ++j) {
        if (nuc >= 0 || kb + j >= k1)
          continue;
        prob += ld(DM.mat_atom_density + kb + j) * sig_t[j];
        if (prob >= cutoff)
          nuc = ld(DM.mat_nuclide + kb + j);
      }
    }
    if (nuc >= 0)
      ce_micro_xs(nuc, i_log, p.E, micro);
  } else {
    for (int k = k0; k < k1; ++k) {
      const int cand = ld(DM.mat_nuclide + k);
      ce_micro_xs(cand, i_log, p.E, micro);
      prob += ld(DM.mat_atom_density + k) * micro.total;
      if (prob >= cutoff) {
        nuc = cand;
        break;
      }
    }
  }
  if (nuc < 0) {
    no_nuclide = true;
    p.wgt = 0.0;
    return;
  }
  if (an_on(gc)) { // p.neutron_xs(p.event_nuclide()), read by the analog scores
    const int64_t idx = an_index(p, gc);
    an_put(gc, idx, AN_MICRO_T, micro.total);
    an_put(gc, idx, AN_MICRO_A, micro.absorption);
    an_put(gc, idx, AN_MICRO_F, micro.fission);
    an_put(gc, idx, AN_MICRO_NF, micro.nu_fission);
  }
  const int rx0 = ld(DM.nuc_rx_ofs + nuc);
  if (ld(DM.nuc_fissionable + nuc) && micro.fission > 0.0) { // create_fission_sites :170-285
    const int rx = rx0 + ld(DM.nuc_fission_rx + nuc);
    const double nu_t = p.wgt / gc.keff * 1.0 * micro.nu_fission / micro.total;
    const int nu = fission_site_count(p, nu_t);
    if (nu > 0) {
      const int32_t base = atomic_add_i32(gc.n_fission, nu);
      const int p0 = ld(DM.rx_prod_ofs + rx);
      int k = 0;
      for (; k < nu; ++k) { // sample_fission_neutron :1043-1117
        const double nu_tot = ce_nu(nuc, p.E, 2);
        const double nu_d = ce_nu(nuc, p.E, 1);
        const double beta = nu_d / nu_tot;
        int group = 0;
        double t = p.time;
        if (prn(p.seed) < beta) {
          const double xi = prn(p.seed) * nu_d;
          double pr = 0.0;
          const int n_prec = ld(DM.nuc_n_precursor + nuc);
          for (group = 1; group < n_prec; ++group) {
            pr += eval_function(ld(DM.prod_yield_fn + p0 + group), p.E);
            if (xi < pr)
              break;
          }
          group = group < n_prec ? group : n_prec;
          t -= log(prn(p.seed)) / ld(DM.prod_decay_rate + p0 + group);
        }
        double E_out = 0.0, mu = 0.0;
        for (int ns = 0; ns < MAX_SAMPLE; ++ns) {
          ce_sample_product(p0 + group, p.E, E_out, mu, p.seed);
          if (E_out < DM.energy_max)
            break;
        }
        const V3 su = rotate_angle(p.u, mu, p.seed);
        const int64_t slot = (int64_t)base + k;
        if (slot >= gc.fission_cap) { // bank full: warn + break, src/physics.cpp:233-247
          fission_bank_full(gc);
          break;
        }
        bank_store(fb, slot, p.r, su, E_out, t, 1.0, group, p.id, p.n_progeny);
        p.n_progeny++;
      }
      if (an_on(gc) && k > 0) { // p.fission(), p.wgt_bank() = sites stored / weight, src/physics.cpp:268-281
        an_puti(gc, an_index(p, gc), AN_FISSION, 1);
        an_puti(gc, an_index(p, gc), AN_FIS_BASE, base); // p.nu_bank(): the k sites from this slot on
        an_put(gc, an_index(p, gc), AN_WGT_BANK, (double)k);
      }
    }
  }
  if (micro.absorption > 0.0) { // absorption :638-672
    if (DM.survival_biasing) {
      const double wgt_absorb = p.wgt * micro.absorption / micro.total;
      p.wgt -= wgt_absorb;
      p.k_abs += wgt_absorb * micro.nu_fission / micro.absorption;
    } else if (micro.absorption > prn(p.seed) * micro.total) {
      p.k_abs += p.wgt * micro.nu_fission / micro.absorption;
      p.wgt = 0.0;
    }
  }
  if (p.wgt == 0.0)
    return;
  // scatter :674-752
  if (an_on(gc))
    an_puti(gc, an_index(p, gc), AN_SCATTER, 1); // p.event() = TallyEvent::SCATTER, :740
  const double cut = prn(p.seed) * (micro.total - micro.absorption);
  const double elastic = ce_reaction_xs(rx0, micro); // Nuclide::calculate_elastic_xs
  if (elastic > cut) {                                 // elastic_scatter :754-818
    const double kT = ld(DM.nuc_kT + nuc);
    const double awr = ld(DM.nuc_awr + nuc);
    double vel = sqrt(p.E);
    V3 v_n = v3(vel * p.u.x, vel * p.u.y, vel * p.u.z);
    V3 v_t = v3(0., 0., 0.);
    if (!(p.E >= DM.free_gas_threshold * kT && awr > 1.0))
      v_t = sample_cxs_target_velocity(awr, p.E, p.u, kT, p.seed);
    const V3 v_cm = v3((v_n.x + awr * v_t.x) / (awr + 1.0), (v_n.y + awr * v_t.y) / (awr + 1.0),
      (v_n.z + awr * v_t.z) / (awr + 1.0));
    v_n = v3(v_n.x - v_cm.x, v_n.y - v_cm.y, v_n.z - v_cm.z);
    vel = sqrt(v_n.x * v_n.x + v_n.y * v_n.y + v_n.z * v_n.z);
    const double mu_cm = -1. + (1. - (-1.)) * prn(p.seed);
    const V3 u_cm = v3(v_n.x / vel, v_n.y / vel, v_n.z / vel);
    const V3 rot = rotate_angle(u_cm, mu_cm, p.seed);
    v_n = v3(vel * rot.x + v_cm.x, vel * rot.y + v_cm.y, vel * rot.z + v_cm.z);
    p.E = dot(v_n, v_n);
    vel = sqrt(p.E);
    p.u = v3(v_n.x / vel, v_n.y / vel, v_n.z / vel);
  } else { // inelastic_scatter :1119-1167
    double pr = elastic;
    int rx = rx0;
    const int rx_end = ld(DM.nuc_rx_ofs + nuc + 1);
    for (int j = rx0; j < rx_end && pr < cut; ++j) {
      if (!ld(DM.rx_inelastic + j))
        continue;
      rx = j;
      pr += ce_reaction_xs(j, micro);
    }
    const double E_in = p.E;
    const int prod = ld(DM.rx_prod_ofs + rx);
    double E, mu;
    ce_sample_product(prod, E_in, E, mu, p.seed);
    if (ld(DM.rx_scatter_in_cm + rx)) {
      const double E_cm = E;
      const double A = ld(DM.nuc_awr + nuc);
      E = E_cm + (E_in + 2.0 * mu * (A + 1.0) * sqrt(E_in * E_cm)) / ((A + 1.0) * (A + 1.0));
      mu = mu * sqrt(E_cm / E) + 1.0 / (A + 1.0) * sqrt(E_in / E);
    }
    if (fabs(mu) > 1.0)
      mu = copysign(1.0, mu);
    p.E = E;
    p.u = rotate_angle(p.u, mu, p.seed);
    const double yield = eval_function(ld(DM.prod_yield_fn + prod), E_in);
    if (floor(yield) == yield && yield > 0) {
      // an integral yield creates exactly yield - 1 secondary neutrons (src/physics.cpp:1158-1162)
#if OMC_SECONDARIES
      if (gc.sec_site != nullptr)
        for (int k = 0; k < (int)round(yield) - 1; ++k)
          create_secondary(p, gc);
#endif
    } else {
      p.wgt *= yield;
    }
  }
  if (DM.survival_biasing && p.wgt < DM.weight_cutoff)
    russian_roulette(p, DM.weight_survive);
}

// =============================================================================
// Tally scoring (tracklength) -- src/tallies/tally_scoring.cpp:31-141, 578-700,
// 1622-1990, 2416-2541 ; src/mesh.cpp:874-929, 999-1146, 1328-1369 ;
// src/tallies/filter_{mesh,energy,cell,material,universe}.cpp
// =============================================================================
OMC_D void atomic_add_f64(double* addr, double val)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  // tally arrays are global memory: skip the generic-address dispatch of atomicAdd
  // (no "memory" clobber: tally arrays are never read by the kernels that add to them, and
  // the clobber would make the compiler reload every hoisted model constant after each add)
  asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(val));
#else
  *addr += val;
#endif
}

OMC_D void tally_add(double* addr, double val, bool few_bins = true)
{
#if OMC_CUDA && defined(__CUDA_ARCH__)
  if (!few_bins) { // hundreds of bins (energy spectra, meshes): lanes rarely agree, the vote only costs
    atomic_add_f64(addr, val);
    return;
  }
  // Warp-aggregated when it pays: if every active lane scores the same bin (cell /
  // material tallies with few bins) the warp sums in registers and issues one RED.
  // Otherwise (energy spectra, meshes: lanes rarely agree) each lane issues its own
  // RED -- measured: a match_any + shuffle loop per score cost 12 % of k_score for a
  // 500-bin spectrum tally and saved almost no atomics (profiles/README.md).
  const unsigned active = __activemask();
  if (active == 0xffffffffu) { // (a partial warp -- lanes in different filter branches -- adds directly)
    int same;
    __match_all_sync(active, (unsigned long long)addr, &same);
    if (same) {
      double sum = val;
#pragma unroll
      for (int d = 16; d > 0; d >>= 1)
        sum += __shfl_xor_sync(active, sum, d);
      if ((threadIdx.x & 31) == 0)
        atomic_add_f64(addr, sum);
      return;
    }
  }
  atomic_add_f64(addr, val);
#else
  *addr += val;
#endif
}

struct ScoreCtx { // what the scores need from the particle
  double flux;
  double xs_total, xs_absorption, xs_fission, xs_nu_fission;
  int material, g, mg_ba;
};

OMC_HD double score_value(int score, const ScoreCtx& s, bool& skip)
{
  skip = false;
  if (DM.run_CE) { // score_general_ce_nonanalog
    switch (score) {
    case OMCGPU_SCORE_FLUX:
      return s.flux;
    case OMCGPU_SCORE_TOTAL:
      return s.xs_total * s.flux;
    case OMCGPU_SCORE_SCATTER:
      return (s.xs_total - s.xs_absorption) * s.flux;
    case OMCGPU_SCORE_ABSORPTION:
      return s.xs_absorption * s.flux;
    case OMCGPU_SCORE_FISSION:
      skip = (s.xs_fission == 0);
      return s.xs_fission * s.flux;
    default:
      skip = (s.xs_fission == 0);
      return s.xs_nu_fission * s.flux;
    }
  }
  const int G = DM.mg_n_groups; // score_general_mg, tracklength branch
  const bool fis = s.material >= 0 && ld(DM.mgm_fissionable + s.material);
  switch (score) {
  case OMCGPU_SCORE_FLUX:
    return s.flux;
  case OMCGPU_SCORE_TOTAL:
    return s.xs_total * s.flux;
  case OMCGPU_SCORE_ABSORPTION:
    return s.xs_absorption * s.flux;
  case OMCGPU_SCORE_SCATTER:
    return s.material >= 0 ? s.flux * mg_scatter_xs(s.mg_ba, s.g) : 0.0;
  case OMCGPU_SCORE_FISSION:
    return fis ? s.flux * ld(DM.mg_fission + s.mg_ba * G + s.g) : s.flux * 0.;
  default:
    return fis ? s.flux * ld(DM.mg_nu_fission + s.mg_ba * G + s.g) : s.flux * 0.;
  }
}

struct MeshDist {
  int next_index;
  double distance;
};

// RegularMesh::distance_to_grid_boundary, src/mesh.cpp:1352-1369, for one axis
// held in registers.  Written without branches (the three axes and the two
// directions would otherwise split a warp six ways): the edge ahead is index
// idx (u > 0) or idx - 1 (u <= 0) and the quotient is discarded when there is
// no edge ahead.
struct MeshAxis {
  double r0, u, ll, w;
  int idx, shape;
};
OMC_HD MeshDist mesh_axis_distance(const MeshAxis& a)
{
  const bool pos = a.u > 0;
  const int edge = pos ? a.idx : a.idx - 1;
  const bool ok = !(fabs(a.u) < FP_PRECISION) && (pos ? a.idx <= a.shape : a.idx >= 1);
  const double q = ((a.ll + edge * a.w) - a.r0) / a.u;
  MeshDist d;
  d.distance = ok ? q : INFTY;
  d.next_index = ok ? (pos ? a.idx + 1 : a.idx - 1) : a.idx;
  return d;
}

OMC_HD int mesh_index(double r, double ll, double w)
{
  return (int)ceil((r - ll) / w);
}

// StructuredMesh::raytrace_mesh, src/mesh.cpp:999-1118, over the part [t0, t1]
// of the track r0 + t u, 0 <= t <= total.  (t0, t1) = (0, total) is the
// reference's loop verbatim.  A track can be cut into consecutive parts that
// are traced independently (k_score hands the parts of long tracks to different
// lanes): every crossing distance is measured from r0 with the reference's
// formula, so the parts emit the same (bin, length) pairs as the whole track,
// with the pair that straddles a cut emitted as two pieces.
// emit(bin, w): w = fraction of `total` spent in mesh bin `bin`.
// (kept out of line on the device: the rare path of k_score, whose hot loop is the DDA below)
template<class Emit>
#if OMC_CUDA
__device__ __noinline__
#else
inline
#endif
void mesh_march_exact(int m, V3 r0, V3 u, double total, double t0, double t1, Emit& emit)
{
  const int nd = ld(DM.mesh_ndim + m);
  MeshAxis a0, a1, a2;
  a0.r0 = r0.x; a0.u = u.x;
  a1.r0 = r0.y; a1.u = u.y;
  a2.r0 = r0.z; a2.u = u.z;
  a0.ll = ld(DM.mesh_lower_left + 3 * m); a0.w = ld(DM.mesh_width + 3 * m);
  a0.shape = ld(DM.mesh_shape + 3 * m);
  // a missing axis is one cell that is never left
  a1.ll = a2.ll = 0.0; a1.w = a2.w = 1.0; a1.shape = a2.shape = 1;
  if (nd > 1) {
    a1.ll = ld(DM.mesh_lower_left + 3 * m + 1); a1.w = ld(DM.mesh_width + 3 * m + 1);
    a1.shape = ld(DM.mesh_shape + 3 * m + 1);
  }
  if (nd > 2) {
    a2.ll = ld(DM.mesh_lower_left + 3 * m + 2); a2.w = ld(DM.mesh_width + 3 * m + 2);
    a2.shape = ld(DM.mesh_shape + 3 * m + 2);
  }
  // the first part starts TINY_BIT inside (mesh.cpp:1021); a later part starts at
  // its cut, where the cell is decided by position alone
  const double ts = (t0 == 0.0) ? TINY_BIT : t0;
  a0.idx = mesh_index(a0.r0 + ts * a0.u, a0.ll, a0.w);
  a1.idx = nd > 1 ? mesh_index(a1.r0 + ts * a1.u, a1.ll, a1.w) : 1;
  a2.idx = nd > 2 ? mesh_index(a2.r0 + ts * a2.u, a2.ll, a2.w) : 1;
#define MESH_OUT(a) ((a).idx < 1 || (a).idx > (a).shape)
#define MESH_BIN() (((a2.idx - 1) * a1.shape + (a1.idx - 1)) * a0.shape + a0.idx - 1)
  bool in_mesh = !(MESH_OUT(a0) || MESH_OUT(a1) || MESH_OUT(a2));
  if (total < 2 * TINY_BIT) {
    if (in_mesh)
      emit(MESH_BIN(), 1.0);
    return;
  }
  MeshDist d0 = mesh_axis_distance(a0), d1, d2;
  d1.next_index = d2.next_index = -1;
  d1.distance = d2.distance = INFTY;
  if (nd > 1)
    d1 = mesh_axis_distance(a1);
  if (nd > 2)
    d2 = mesh_axis_distance(a2);
  double traveled = t0;
  for (int guard = 0; guard < 1000000; ++guard) {
    if (in_mesh) {
      int k = 0;
      double dk = d0.distance;
      if (d1.distance < dk) {
        k = 1;
        dk = d1.distance;
      }
      if (d2.distance < dk) {
        k = 2;
        dk = d2.distance;
      }
      emit(MESH_BIN(), ((dk < t1 ? dk : t1) - traveled) / total);
      traveled = dk;
      if (traveled >= t1)
        break;
      // step across the nearest edge: select the axis, one distance, select back
      MeshAxis ak = k == 0 ? a0 : (k == 1 ? a1 : a2);
      ak.idx = k == 0 ? d0.next_index : (k == 1 ? d1.next_index : d2.next_index);
      const MeshDist dn = mesh_axis_distance(ak);
      if (k == 0) {
        a0.idx = ak.idx;
        d0 = dn;
      }
      if (k == 1) {
        a1.idx = ak.idx;
        d1 = dn;
      }
      if (k == 2) {
        a2.idx = ak.idx;
        d2 = dn;
      }
      in_mesh = !MESH_OUT(ak);
    } else {
      int k_max = -1;
      if (MESH_OUT(a0) && d0.distance > traveled) {
        traveled = d0.distance;
        k_max = 0;
      }
      if (nd > 1 && MESH_OUT(a1) && d1.distance > traveled) {
        traveled = d1.distance;
        k_max = 1;
      }
      if (nd > 2 && MESH_OUT(a2) && d2.distance > traveled) {
        traveled = d2.distance;
        k_max = 2;
      }
      if (k_max == -1)
        traveled += TINY_BIT;
      if (traveled >= t1)
        break;
      const double tt = traveled + TINY_BIT;
      a0.idx = mesh_index(a0.r0 + tt * a0.u, a0.ll, a0.w);
      if (nd > 1)
        a1.idx = mesh_index(a1.r0 + tt * a1.u, a1.ll, a1.w);
      if (nd > 2)
        a2.idx = mesh_index(a2.r0 + tt * a2.u, a2.ll, a2.w);
      in_mesh = !(MESH_OUT(a0) || MESH_OUT(a1) || MESH_OUT(a2));
      d0 = mesh_axis_distance(a0);
      if (nd > 1)
        d1 = mesh_axis_distance(a1);
      if (nd > 2)
        d2 = mesh_axis_distance(a2);
    }
  }
#undef MESH_OUT
#undef MESH_BIN
}

template<int ND, class Emit>
OMC_D void mesh_march_nd(int m, V3 r0, V3 u, double total, double t0, double t1, Emit& emit)
{
  const double ts = (t0 == 0.0) ? TINY_BIT : t0;
  double tnext[ND], tdel[ND]; // distance of the next crossing of axis k, distance between two of them
  int left[ND], stride[ND];   // cells ahead before the mesh ends, bin increment of a step
  bool in_mesh = true;
  int bin = 0, cells = 1;
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    const double ll = ld(DM.mesh_lower_left + 3 * m + k), w = ld(DM.mesh_width + 3 * m + k);
    const int shape = ld(DM.mesh_shape + 3 * m + k);
    const double rk = comp(r0, k), uk = comp(u, k);
    const int idx = mesh_index(rk + ts * uk, ll, w);
    in_mesh = in_mesh && idx >= 1 && idx <= shape;
    const bool pos = uk > 0;
    tnext[k] = INFTY;
    tdel[k] = 0.0;
    if (!(fabs(uk) < FP_PRECISION)) {
      tnext[k] = ((ll + (pos ? idx : idx - 1) * w) - rk) / uk; // the reference's formula for the first crossing
      tdel[k] = w / fabs(uk);
    }
    left[k] = pos ? shape - idx : idx - 1;
    stride[k] = pos ? cells : -cells;
    bin += (idx - 1) * cells;
    cells *= shape;
  }
  if (!in_mesh || total < 2 * TINY_BIT) {
    mesh_march_exact(m, r0, u, total, t0, t1, emit);
    return;
  }
  const double inv_total = 1.0 / total;
  double traveled = t0;
  for (int guard = 0; guard < 1000000; ++guard) {
    // nearest crossing; ties go to the lower axis as in the reference's loop
    int k = 0;
    double dk = tnext[0];
    if (ND > 1 && tnext[1] < dk) {
      k = 1;
      dk = tnext[1];
    }
    if (ND > 2 && tnext[2] < dk) {
      k = 2;
      dk = tnext[2];
    }
    emit(bin, ((dk < t1 ? dk : t1) - traveled) * inv_total);
    traveled = dk;
    if (traveled >= t1)
      break;
    // step across that edge without branching on the axis (the lanes of a warp cross different
    // axes: an if / else chain would run the bodies one after another)
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      const bool ka = k == a;
      if (ka && left[a] == 0) // the mesh ends here (a straight track cannot re-enter a box)
        return;
      left[a] -= ka ? 1 : 0;
      bin += ka ? stride[a] : 0;
      tnext[a] += ka ? tdel[a] : 0.0;
    }
  }
}

// The same ray trace as a DDA (Amanatides & Woo) for a part that starts INSIDE
// the mesh -- the case of every flight when the mesh covers the geometry.  The
// first crossing of each axis uses the reference's formula; every later one adds w / |u|
// instead of a fresh division, and a length fraction is (length * 1/total).  Crossing
// distances therefore differ from the reference's by rounding (a few ulp per cell crossed,
// over at most the cells of ONE part, since every part starts afresh from its own position),
// i.e. bin sequences are the same and segment lengths agree to ~1e-14 relative; tally sums are
// not bit-reproducible in the reference either (atomic order).  A part that starts outside the
// mesh takes the exact routine above.  Compiled per mesh dimension: a 2-D mesh (configs C1-C4)
// carries no third axis through the loop.
template<class Emit>
OMC_D void mesh_march(int m, V3 r0, V3 u, double total, double t0, double t1, Emit& emit)
{
  const int nd = ld(DM.mesh_ndim + m);
  if (nd == 2)
    mesh_march_nd<2>(m, r0, u, total, t0, t1, emit);
  else if (nd == 3)
    mesh_march_nd<3>(m, r0, u, total, t0, t1, emit);
  else
    mesh_march_nd<1>(m, r0, u, total, t0, t1, emit);
}

// -----------------------------------------------------------------------------
// The same DDA with the per-FLIGHT constants taken out (k_score, second version).  A flight's own
// lane computes them once -- the only divisions of the ray trace: first crossing of every axis (the
// reference's formula), distance between crossings w / |u| and its inverse -- and publishes them;
// any lane can then start ANY part [ta, tb] of the flight with a multiply per axis
// (mesh_part_begin: crossings of axis k before ta = floor((ta - tnext0) |u| / w) + 1), so parts
// can be a few cells short and the lanes of a warp stay in step.  Crossing distances inside a part
// are tnext0 + n tdel followed by additions of tdel: as in mesh_march_nd they differ from the
// reference's per-crossing quotient by rounding only, the bins are the same, and a piece that
// straddles a cut is emitted as two pieces of the same bin.
// -----------------------------------------------------------------------------
constexpr int MESH_RAY_AXES = 3;
struct MeshRay {
  double tnext0[MESH_RAY_AXES], tdel[MESH_RAY_AXES], inv_tdel[MESH_RAY_AXES];
  int left[MESH_RAY_AXES], stride[MESH_RAY_AXES]; // cells ahead before the mesh ends, bin increment of a step
  int bin0;
};

// false: the flight starts outside the mesh (the exact routine traces it).  Axes the mesh does not
// have are never crossed (tnext0 = INFTY).
OMC_HD bool mesh_ray_setup(int m, V3 r0, V3 u, MeshRay& ray)
{
  const int nd = ld(DM.mesh_ndim + m);
  bool in_mesh = true;
  int bin = 0, cells = 1;
#pragma unroll
  for (int k = 0; k < MESH_RAY_AXES; ++k) {
    ray.tnext0[k] = INFTY;
    ray.tdel[k] = 0.0;
    ray.inv_tdel[k] = 0.0;
    ray.left[k] = 0;
    ray.stride[k] = 0;
    if (k < nd) {
      const double ll = ld(DM.mesh_lower_left + 3 * m + k), w = ld(DM.mesh_width + 3 * m + k);
      const int shape = ld(DM.mesh_shape + 3 * m + k);
      const double rk = comp(r0, k), uk = comp(u, k);
      const int idx = mesh_index(rk + TINY_BIT * uk, ll, w);
      in_mesh = in_mesh && idx >= 1 && idx <= shape;
      const bool pos = uk > 0;
      if (!(fabs(uk) < FP_PRECISION)) {
        ray.tnext0[k] = ((ll + (pos ? idx : idx - 1) * w) - rk) / uk; // the reference's formula for the first crossing
        ray.tdel[k] = w / fabs(uk);
        ray.inv_tdel[k] = fabs(uk) / w;
      }
      ray.left[k] = pos ? shape - idx : idx - 1;
      ray.stride[k] = pos ? cells : -cells;
      bin += (idx - 1) * cells;
      cells *= shape;
    }
  }
  ray.bin0 = bin;
  return in_mesh;
}

// crossings of one axis strictly before distance t, at most left + 1 (= the ray has left the mesh);
// tn: distance of the first crossing at or after t
OMC_HD int mesh_ray_crossed(double tnext0, double tdel, double inv_tdel, int left, double t, double& tn)
{
  int n = 0;
  tn = tnext0;
  if (t > tnext0) {
    const double q = (t - tnext0) * inv_tdel;
    n = q >= (double)left + 1.0 ? left + 1 : (int)q + 1;
    tn = tnext0 + n * tdel;
    if (tn < t) { // the quotient was rounded down across an integer
      ++n;
      tn += tdel;
    }
  }
  return n;
}

// mesh cells the flight visits inside the mesh (to balance the parts)
OMC_HD int mesh_ray_cells(const MeshRay& ray, double total)
{
  int n = 1;
#pragma unroll
  for (int k = 0; k < MESH_RAY_AXES; ++k) {
    double tn;
    const int c = mesh_ray_crossed(ray.tnext0[k], ray.tdel[k], ray.inv_tdel[k], ray.left[k], total, tn);
    n += c > ray.left[k] ? ray.left[k] : c;
  }
  return n;
}

// the part [ta, tb] of a flight that starts inside the mesh: emit(bin, fraction of total).
// RAY: anything with the members of MeshRay read as ray.member(k) -- registers on the host,
// the warp's staging area in k_score
template<int ND, bool LENGTHS = false, class Ray, class Emit>
OMC_D void mesh_part_march(const Ray& ray, double ta, double tb, double inv_total, Emit& emit)
{ // LENGTHS: emit(bin, length) -- the caller has folded 1 / total into its scores
  double tnext[ND], tdel[ND];
  int left[ND], stride[ND];
  int bin = ray.bin0();
#pragma unroll
  for (int k = 0; k < ND; ++k) {
    tdel[k] = ray.tdel(k);
    stride[k] = ray.stride(k);
    const int l = ray.left(k);
    const int n = mesh_ray_crossed(ray.tnext0(k), tdel[k], ray.inv_tdel(k), l, ta, tnext[k]);
    if (n > l)
      return; // the flight left the mesh before this part
    left[k] = l - n;
    bin += n * stride[k];
  }
  double traveled = ta;
#pragma unroll 1
  for (;;) { // ends: every crossing distance grows by tdel > 0, an axis that is never crossed sits at INFTY
    int k = 0;
    double dk = tnext[0];
    if (ND > 1 && tnext[1] < dk) {
      k = 1;
      dk = tnext[1];
    }
    if (ND > 2 && tnext[2] < dk) {
      k = 2;
      dk = tnext[2];
    }
    const double len = (dk < tb ? dk : tb) - traveled;
    emit(bin, LENGTHS ? len : len * inv_total);
    traveled = dk;
    if (traveled >= tb)
      break;
#pragma unroll
    for (int a = 0; a < ND; ++a) {
      const bool ka = k == a;
      if (ka && left[a] == 0)
        return;
      left[a] -= ka ? 1 : 0;
      bin += ka ? stride[a] : 0;
      tnext[a] += ka ? tdel[a] : 0.0;
    }
  }
}

struct MeshRayRegs { // Ray view of a MeshRay in registers (host compile, single-lane paths)
  const MeshRay* r;
  OMC_HD double tnext0(int k) const { return r->tnext0[k]; }
  OMC_HD double tdel(int k) const { return r->tdel[k]; }
  OMC_HD double inv_tdel(int k) const { return r->inv_tdel[k]; }
  OMC_HD int left(int k) const { return r->left[k]; }
  OMC_HD int stride(int k) const { return r->stride[k]; }
  OMC_HD int bin0() const { return r->bin0; }
};

// Number of mesh cells the track r0 -> r1 visits (exact for a track inside the
// mesh, an estimate otherwise) -- only used to balance work between lanes.
OMC_HD int mesh_cells_visited(int m, V3 r0, V3 r1, V3 u)
{
  const int nd = ld(DM.mesh_ndim + m);
  int n = 1;
  for (int k = 0; k < nd; ++k) {
    const int sk = ld(DM.mesh_shape + 3 * m + k);
    const double ll = ld(DM.mesh_lower_left + 3 * m + k), w = ld(DM.mesh_width + 3 * m + k);
    int a = mesh_index(comp(r0, k) + TINY_BIT * comp(u, k), ll, w);
    int b = mesh_index(comp(r1, k), ll, w);
    a = a < 0 ? 0 : (a > sk + 1 ? sk + 1 : a);
    b = b < 0 ? 0 : (b > sk + 1 ? sk + 1 : b);
    n += a > b ? a - b : b - a;
  }
  return n;
}

// A tally evaluated for one track: the matches of its non-mesh filters
// (FilterBinIter, src/tallies/tally_scoring.cpp:31-141); a mesh filter, if
// present at position `mesh_pos`, streams its (bin, weight) pairs from the ray
// trace.
struct TallyEval {
  int nf;
  int mesh_pos;
  int mesh;
  int nb[OMCGPU_MAX_FILTERS];
  int bins[OMCGPU_MAX_FILTERS][OMCGPU_MAX_COORD];
  int stride[OMCGPU_MAX_FILTERS];
};

// StructuredMesh::get_bin, src/mesh.cpp:874-929: bin of a point, -1 outside the mesh
OMC_HD int mesh_get_bin(int m, V3 r)
{
  const int nd = ld(DM.mesh_ndim + m);
  int bin = 0, cells = 1;
  for (int k = 0; k < nd; ++k) {
    const int shape = ld(DM.mesh_shape + 3 * m + k);
    const int idx = mesh_index(comp(r, k), ld(DM.mesh_lower_left + 3 * m + k), ld(DM.mesh_width + 3 * m + k));
    if (idx < 1 || idx > shape)
      return -1;
    bin += (idx - 1) * cells;
    cells *= shape;
  }
  return bin;
}

// DistribcellFilter::get_all_bins, src/tallies/filter_distribcell.cpp:40-63: the instance of
// `cell` the particle is in, or -1
OMC_HD int distribcell_bin(const PState& ps, int64_t i, int n_coord, int cell)
{
  const int map = gld(DM.cell_distribcell_index + cell);
  int offset = 0;
  for (int l = 0; l < n_coord; ++l) {
    const int c = LV(cell, l);
    const int type = gld(DM.cell_type + c);
    if (type == OMCGPU_FILL_UNIVERSE) {
      offset += gld(DM.cell_offset + gld(DM.cell_offset_ofs + c) + map);
    } else if (type == OMCGPU_FILL_LATTICE) {
      const int lat = LV(lattice, l + 1);
      LatIdx li;
      li.x = LV(lat_x, l + 1);
      li.y = LV(lat_y, l + 1);
      li.z = LV(lat_z, l + 1);
      if (lat_valid(lat, li)) {
        const int32_t* n = DM.lat_shape + 3 * lat;
        offset += gld(DM.lat_offsets + gld(DM.lat_offset_ofs + lat) + gld(n) * gld(n + 1) * gld(n + 2) * map +
                     lat_flat(lat, li)) +
                  gld(DM.cell_offset + gld(DM.cell_offset_ofs + c) + map);
      }
    }
    if (c == cell)
      return offset;
  }
  return -1;
}

// returns false when some filter of tally t has no matching bin.  `r_point` non-null: the tally
// is scored at a point with the collision estimator (the mesh filter matches the bin of the
// point, MeshFilter::get_all_bins src/tallies/filter_mesh.cpp:61-66; sc.g is then the
// pre-collision group, EnergyFilter::get_all_bins src/tallies/filter_energy.cpp:62-67)
OMC_HD bool tally_prepare(const PState& ps, int64_t i, int n_coord, int t, double E_last,
  const ScoreCtx& sc, TallyEval& te, const V3* r_point = nullptr, double E_post = 0.0, int g_post = 0)
{
  const int G = DM.mg_n_groups;
  te.nf = ld(DM.tally_n_filters + t);
  te.mesh_pos = -1;
  te.mesh = -1;
  for (int k = 0; k < te.nf; ++k) {
    const int f = ld(DM.tally_filters + OMCGPU_MAX_FILTERS * t + k);
    te.stride[k] = ld(DM.tally_strides + OMCGPU_MAX_FILTERS * t + k);
    te.nb[k] = 0;
    const int o = ld(DM.filter_data_ofs + f);
    switch (ld(DM.filter_type + f)) {
    case OMCGPU_FILTER_MESH:
      if (r_point) {
        const int b = mesh_get_bin(ld(DM.filter_mesh + f), *r_point);
        if (b >= 0)
          te.bins[k][te.nb[k]++] = b;
      } else {
        te.mesh_pos = k;
        te.mesh = ld(DM.filter_mesh + f);
        te.nb[k] = 1;
      }
      break;
    case OMCGPU_FILTER_ENERGY: {
      const int nbin = ld(DM.filter_n_bins + f);
      if (!DM.run_CE && ld(DM.filter_matches_groups + f)) {
        te.bins[k][0] = G - sc.g - 1;
        te.nb[k] = 1;
      } else {
        const double* e = DM.filter_ddata + o;
        if (E_last >= ld(e) && E_last <= ld(e + nbin)) {
          te.bins[k][0] = (int)lower_bound_index(e, (int64_t)nbin + 1, E_last);
          te.nb[k] = 1;
        }
      }
      break;
    }
    case OMCGPU_FILTER_ENERGYOUT: { // EnergyoutFilter::get_all_bins, src/tallies/filter_energy.cpp:100-112:
      const int nbin = ld(DM.filter_n_bins + f); // the energy / group the particle LEFT the collision with
      if (!DM.run_CE && ld(DM.filter_matches_groups + f)) {
        te.bins[k][0] = G - g_post - 1;
        te.nb[k] = 1;
      } else {
        const double* e = DM.filter_ddata + o;
        if (E_post >= ld(e) && E_post <= ld(e + nbin)) {
          te.bins[k][0] = (int)lower_bound_index(e, (int64_t)nbin + 1, E_post);
          te.nb[k] = 1;
        }
      }
      break;
    }
    case OMCGPU_FILTER_CELL:
      for (int l = 0; l < n_coord; ++l) {
        const int b = ld(DM.filter_idata + o + LV(cell, l));
        if (b >= 0)
          te.bins[k][te.nb[k]++] = b;
      }
      break;
    case OMCGPU_FILTER_MATERIAL:
      if (sc.material >= 0) {
        // (derived materials of a multi-temperature model map back to the model's material)
        const int b = ld(DM.filter_idata + o + (DM.mat_base ? ld(DM.mat_base + sc.material) : sc.material));
        if (b >= 0)
          te.bins[k][te.nb[k]++] = b;
      }
      break;
    case OMCGPU_FILTER_DISTRIBCELL: {
      const int b = distribcell_bin(ps, i, n_coord, ld(DM.filter_idata + o));
      if (b >= 0)
        te.bins[k][te.nb[k]++] = b;
      break;
    }
    default: // universe
      for (int l = 0; l < n_coord; ++l) {
        const int b = ld(DM.filter_idata + o + LV(universe, l));
        if (b >= 0)
          te.bins[k][te.nb[k]++] = b;
      }
      break;
    }
    if (te.nb[k] == 0)
      return false;
  }
  return true;
}

// SCORE: (score code, skip&) -> value of the score for this event
template<class Score>
OMC_D void tally_emit_with(const TallyEval& te, int t, int mesh_bin, double mesh_w, const Score& score_of,
  double* values)
{
  const int ns = ld(DM.tally_n_scores + t);
  const int32_t* scores = DM.tally_scores + OMCGPU_MAX_SCORES * t;
  const int64_t r_ofs = ld(DM.tally_result_ofs + t);
  double* results = values + r_ofs;
  const bool few_bins = ld(DM.tally_result_ofs + t + 1) - r_ofs <= 64 * ns; // warp-aggregate only small tallies
  // cartesian product over the non-mesh filters (at most MAX_COORD^2 combos; 1 in practice)
  int n0 = 1, n1 = 1, n2 = 1;
  if (te.nf > 0)
    n0 = (te.mesh_pos == 0) ? 1 : te.nb[0];
  if (te.nf > 1)
    n1 = (te.mesh_pos == 1) ? 1 : te.nb[1];
  if (te.nf > 2)
    n2 = (te.mesh_pos == 2) ? 1 : te.nb[2];
  for (int a = 0; a < n0; ++a)
    for (int b = 0; b < n1; ++b)
      for (int c = 0; c < n2; ++c) {
        int64_t index = 0;
        if (te.nf > 0)
          index += (int64_t)((te.mesh_pos == 0) ? mesh_bin : te.bins[0][a]) * te.stride[0];
        if (te.nf > 1)
          index += (int64_t)((te.mesh_pos == 1) ? mesh_bin : te.bins[1][b]) * te.stride[1];
        if (te.nf > 2)
          index += (int64_t)((te.mesh_pos == 2) ? mesh_bin : te.bins[2][c]) * te.stride[2];
        for (int s = 0; s < ns; ++s) {
          bool skip;
          const double score = score_of(ld(scores + s), skip);
          if (!skip)
            tally_add(results + index * ns + s, score * mesh_w, few_bins);
        }
      }
}

struct NonAnalogScore { // score_general_ce_nonanalog / score_general_mg (tracklength, collision)
  const ScoreCtx* sc;
  OMC_D double operator()(int score, bool& skip) const { return score_value(score, *sc, skip); }
};
OMC_D void tally_emit(const TallyEval& te, int t, int mesh_bin, double mesh_w, const ScoreCtx& sc,
  double* values)
{
  NonAnalogScore f;
  f.sc = &sc;
  tally_emit_with(te, t, mesh_bin, mesh_w, f, values);
}

// score_general_ce_analog (src/tallies/tally_scoring.cpp:1116-1330) and the ANALOG branches of
// score_general_mg (:1622-1990), i_nuclide = -1, flux = 1 (neutrons): the six scores of the path
struct AnalogScore {
  double wgt_last, wgt_bank, keff;
  double xs_total, xs_absorption, xs_fission; // macroscopic, before the collision
  double micro_t, micro_a, micro_f, micro_nf; // of the nuclide collided with (CE)
  double mg_fis_over_abs, mg_nufis_over_abs;  // table values at (angle, temperature, p_g) (MG)
  bool scatter, fission;                      // p.event() == SCATTER, p.fission()
  bool eout;                                  // the tally has an EnergyoutFilter: nu-fission goes site by site
  OMC_D double operator()(int score, bool& skip) const
  {
    skip = false;
    if (eout && score == OMCGPU_SCORE_NU_FISSION) { // score_fission_eout (score_after_collision), or nothing
      skip = true;
      return 0.0;
    }
    const bool sb = DM.survival_biasing != 0;
    if (DM.run_CE) {
      const double wgt_absorb = sb ? wgt_last * micro_a / micro_t : 0.0;
      switch (score) {
      case OMCGPU_SCORE_FLUX:
        return 1.0 * wgt_last / xs_total;
      case OMCGPU_SCORE_TOTAL:
        return wgt_last * 1.0;
      case OMCGPU_SCORE_SCATTER:
        skip = !scatter;
        return (wgt_last - wgt_absorb) * 1.0;
      case OMCGPU_SCORE_ABSORPTION:
        if (sb)
          return wgt_absorb * 1.0;
        skip = scatter;
        return wgt_last * 1.0;
      case OMCGPU_SCORE_FISSION:
        if (xs_fission == 0) {
          skip = true;
          return 0.0;
        }
        if (sb)
          return micro_t > 0 ? wgt_last * micro_f / micro_t * 1.0 : 0.;
        skip = scatter;
        return wgt_last * micro_f / micro_a * 1.0;
      default: // nu-fission
        if (xs_fission == 0) {
          skip = true;
          return 0.0;
        }
        if (sb)
          return micro_t > 0 ? wgt_last * micro_nf / micro_t * 1.0 : 0.;
        skip = !fission;
        return keff * wgt_bank * 1.0;
      }
    }
    const double wgt_absorb = sb ? wgt_last * xs_absorption / xs_total : 0.0;
    switch (score) {
    case OMCGPU_SCORE_FLUX:
      return 1.0 * wgt_last / xs_total;
    case OMCGPU_SCORE_TOTAL:
      return 1.0 * wgt_last;
    case OMCGPU_SCORE_SCATTER:
      skip = !scatter;
      return (wgt_last - wgt_absorb) * 1.0;
    case OMCGPU_SCORE_ABSORPTION:
      if (sb)
        return wgt_absorb * 1.0;
      skip = scatter;
      return wgt_last * 1.0;
    case OMCGPU_SCORE_FISSION: {
      double s;
      if (sb) {
        s = wgt_absorb * 1.0;
      } else {
        skip = scatter;
        s = wgt_last * 1.0;
      }
      return s * mg_fis_over_abs;
    }
    default: // nu-fission
      if (sb)
        return wgt_absorb * 1.0 * mg_nufis_over_abs;
      skip = !fission;
      return keff * wgt_bank * 1.0;
    }
  }
};

struct GenericEmit { // one lane scores its whole track: any filter combination
  const TallyEval* te;
  int t;
  const ScoreCtx* sc;
  double* values;
  OMC_D void operator()(int bin, double w) const { tally_emit(*te, t, bin, w, *sc, values); }
};

// Track geometry of the mesh ray trace: total_distance is |r1 - r0|, not the
// flight distance (src/mesh.cpp:1008-1010)
OMC_HD double track_length(V3 r0, V3 r1)
{
  const double dx = r1.x - r0.x, dy = r1.y - r0.y, dz = r1.z - r0.z;
  return sqrt(dx * dx + dy * dy + dz * dz);
}

// score_tracklength_tally by ONE lane for ONE track (the tail kernel and the
// rare tallies with several matches per filter); src/tallies/tally_scoring.cpp:2416-2541
OMC_OOL void score_tracklength(const PState& ps, int64_t i, int n_coord, V3 r0, V3 r1, V3 u,
  double E_last, const ScoreCtx& sc, double* values)
{
  for (int t = 0; t < DM.n_tallies; ++t) {
    if (ld(DM.tally_estimator + t) != OMCGPU_ESTIMATOR_TRACKLENGTH)
      continue;
    TallyEval te;
    if (!tally_prepare(ps, i, n_coord, t, E_last, sc, te))
      continue;
    if (te.mesh_pos < 0) {
      tally_emit(te, t, 0, 1.0, sc, values);
      continue;
    }
    const double total = track_length(r0, r1);
    if (total == 0.0)
      continue;
    GenericEmit em;
    em.te = &te;
    em.t = t;
    em.sc = &sc;
    em.values = values;
    mesh_march(te.mesh, r0, u, total, 0.0, total, em);
  }
}

// which tallies are scored AFTER the collision from the collision record: the analog estimator,
// and the collision estimator of multigroup runs with survival biasing (score_general_mg takes
// the group from whether the particle is still alive after the roulette, :1632-1657)
OMC_HD bool tally_scored_after_collision(int t)
{
  const int est = ld(DM.tally_estimator + t);
  return est == OMCGPU_ESTIMATOR_ANALOG ||
         (est == OMCGPU_ESTIMATOR_COLLISION && !DM.run_CE && DM.survival_biasing != 0);
}
OMC_HD bool tally_scored_before_collision(int t)
{
  return ld(DM.tally_estimator + t) == OMCGPU_ESTIMATOR_COLLISION && !tally_scored_after_collision(t);
}

// score_collision_tally, src/tallies/tally_scoring.cpp:2543-2620: the tallies with the collision
// estimator, at the point of a collision.  sc carries flux = wgt_last / Sigma_t and the
// PRE-collision macroscopic cross sections, group and (MG) angle block; E_last the pre-collision
// energy -- nothing a score of score_general_ce_nonanalog / score_general_mg reads changes in the
// collision, so the collision kernel scores before it calls collision().
OMC_OOL void score_collision(const PState& ps, int64_t i, int n_coord, V3 r, double E_last, const ScoreCtx& sc,
  double* values)
{
  for (int t = 0; t < DM.n_tallies; ++t) {
    if (!tally_scored_before_collision(t))
      continue;
    TallyEval te;
    if (!tally_prepare(ps, i, n_coord, t, E_last, sc, te, &r))
      continue;
    tally_emit(te, t, 0, 1.0, sc, values);
  }
}

// score_analog_tally_ce / _mg (src/tallies/tally_scoring.cpp:2314-2414) and, for the case above,
// score_collision_tally, from the record the collision left for particle idx
OMC_OOL void score_after_collision(const PState& ps, int64_t i, int n_coord, const GenCtx& gc, int64_t idx,
  double* values, const double* bank_E)
{
  const double* d = gc.an_d + idx;
  const int32_t* n = gc.an_i + idx;
  const int64_t cap = gc.an_cap;
  const V3 r = v3(d[AN_RX * cap], d[AN_RY * cap], d[AN_RZ * cap]);
  const double E_last = d[AN_E_LAST * cap];
  const int g_last = n[AN_G_LAST * cap];
  // the group of the scores: the new one when the roulette killed the particle after its scatter
  const int p_g = (DM.survival_biasing && !n[AN_ALIVE * cap]) ? n[AN_G_POST * cap] : g_last;
  ScoreCtx sc;
  sc.xs_total = d[AN_XS_T * cap];
  sc.xs_absorption = d[AN_XS_A * cap];
  sc.xs_fission = d[AN_XS_F * cap];
  sc.xs_nu_fission = d[AN_XS_NF * cap];
  sc.flux = d[AN_WGT_LAST * cap] / sc.xs_total;
  sc.material = n[AN_MATERIAL * cap];
  sc.mg_ba = n[AN_MG_BA * cap];
  AnalogScore an;
  an.wgt_last = d[AN_WGT_LAST * cap];
  an.wgt_bank = d[AN_WGT_BANK * cap];
  an.keff = gc.keff;
  an.xs_total = sc.xs_total;
  an.xs_absorption = sc.xs_absorption;
  an.xs_fission = sc.xs_fission;
  an.micro_t = d[AN_MICRO_T * cap];
  an.micro_a = d[AN_MICRO_A * cap];
  an.micro_f = d[AN_MICRO_F * cap];
  an.micro_nf = d[AN_MICRO_NF * cap];
  an.scatter = n[AN_SCATTER * cap] != 0;
  an.fission = n[AN_FISSION * cap] != 0;
  an.eout = false;
  an.mg_fis_over_abs = an.mg_nufis_over_abs = 0.0;
  if (!DM.run_CE) {
    const int G = DM.mg_n_groups;
    const bool fis = ld(DM.mgm_fissionable + sc.material);
    const double a = ld(DM.mg_absorption + sc.mg_ba * G + p_g);
    an.mg_fis_over_abs = (fis ? ld(DM.mg_fission + sc.mg_ba * G + p_g) : 0.) / a;
    an.mg_nufis_over_abs = (fis ? ld(DM.mg_nu_fission + sc.mg_ba * G + p_g) : 0.) / a;
  }
  for (int t = 0; t < DM.n_tallies; ++t) {
    if (!tally_scored_after_collision(t))
      continue;
    TallyEval te;
    sc.g = g_last; // the energy filter matches the group the particle collided in
    if (!tally_prepare(ps, i, n_coord, t, E_last, sc, te, &r, d[AN_E_POST * cap], n[AN_G_POST * cap]))
      continue;
    if (ld(DM.tally_estimator + t) == OMCGPU_ESTIMATOR_ANALOG) {
      int eo_pos = -1; // tally.energyout_filter_
      for (int k = 0; k < te.nf; ++k)
        if (ld(DM.filter_type + ld(DM.tally_filters + OMCGPU_MAX_FILTERS * t + k)) == OMCGPU_FILTER_ENERGYOUT)
          eo_pos = k;
      an.eout = eo_pos >= 0;
      tally_emit_with(te, t, 0, 1.0, an, values);
      // score_fission_eout, src/tallies/tally_scoring.cpp:383-470: nu-fission of a tally with an
      // EnergyoutFilter, once per fission site banked in this collision (the wgt_bank sites of the
      // fission bank from AN_FIS_BASE on), into the outgoing-energy bin of that site
      if (eo_pos >= 0 && (DM.survival_biasing || an.fission) && an.fission && !(DM.run_CE && sc.xs_fission == 0)) {
        const int ns = ld(DM.tally_n_scores + t);
        const int fe = ld(DM.tally_filters + OMCGPU_MAX_FILTERS * t + eo_pos);
        const double* eb = DM.filter_ddata + ld(DM.filter_data_ofs + fe);
        const int neb = ld(DM.filter_n_bins + fe);
        const int base = n[AN_FIS_BASE * cap], n_bank = (int)an.wgt_bank;
        double* results = values + ld(DM.tally_result_ofs + t);
        for (int s = 0; s < ns; ++s) {
          if (ld(DM.tally_scores + OMCGPU_MAX_SCORES * t + s) != OMCGPU_SCORE_NU_FISSION)
            continue;
          for (int b = 0; b < n_bank; ++b) {
            const double E_site = bank_E[base + b];
            int bin_out;
            if (!DM.run_CE && ld(DM.filter_matches_groups + fe)) {
              bin_out = neb - (int)E_site - 1;
            } else {
              const double E_out = DM.run_CE ? E_site : ld(DM.mg_energy_bin_avg + (int)E_site);
              if (E_out < ld(eb) || E_out > ld(eb + neb))
                continue;
              bin_out = (int)lower_bound_index(eb, (int64_t)neb + 1, E_out);
            }
            // the other filters' matches as tally_emit_with walks them
            int n0 = te.nf > 0 ? (eo_pos == 0 ? 1 : te.nb[0]) : 1, n1 = te.nf > 1 ? (eo_pos == 1 ? 1 : te.nb[1]) : 1,
                n2 = te.nf > 2 ? (eo_pos == 2 ? 1 : te.nb[2]) : 1;
            for (int a = 0; a < n0; ++a)
              for (int bb = 0; bb < n1; ++bb)
                for (int c = 0; c < n2; ++c) {
                  int64_t index = 0;
                  if (te.nf > 0)
                    index += (int64_t)(eo_pos == 0 ? bin_out : te.bins[0][a]) * te.stride[0];
                  if (te.nf > 1)
                    index += (int64_t)(eo_pos == 1 ? bin_out : te.bins[1][bb]) * te.stride[1];
                  if (te.nf > 2)
                    index += (int64_t)(eo_pos == 2 ? bin_out : te.bins[2][c]) * te.stride[2];
                  tally_add(results + index * ns + s, gc.keff * 1.0, false); // keff * site weight (1 / weight)
                }
          }
        }
      }
    } else {
      sc.g = p_g;
      tally_emit(te, t, 0, 1.0, sc, values);
    }
  }
}

} // namespace omc

