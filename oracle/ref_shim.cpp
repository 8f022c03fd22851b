// extern "C" doorway into the reference's OWN translation units (compiled
// verbatim from /root/reference by oracle/Makefile into oracle/_ref/).  Used by
// tests/test_oracle_vs_reference.py to validate the restatement in oracle.cpp
// and, through it, the device code.  TEST INFRASTRUCTURE ONLY.
#include <cstdint>

#include "openmc/math_functions.h"
#include "openmc/position.h"
#include "openmc/random_dist.h"
#include "openmc/random_lcg.h"
#include "openmc/search.h"

extern "C" {

double ref_prn(uint64_t* seed) { return openmc::prn(seed); }
uint64_t ref_future_seed(uint64_t n, uint64_t seed) { return openmc::future_seed(n, seed); }
uint64_t ref_init_seed(int64_t id, int offset) { return openmc::init_seed(id, offset); }
void ref_set_seed(int64_t s) { openmc::openmc_set_seed(s); }
void ref_set_stride(uint64_t s) { openmc::openmc_set_stride(s); }
double ref_evaluate_legendre(int n, const double* data, double x)
{
  return openmc::evaluate_legendre(n, data, x);
}
void ref_rotate_angle(double* uvw, double mu, const double* phi, uint64_t* seed)
{
  openmc::rotate_angle_c(uvw, mu, phi, seed);
}
double ref_maxwell_spectrum(double T, uint64_t* seed) { return openmc::maxwell_spectrum(T, seed); }
double ref_watt_spectrum(double a, double b, uint64_t* seed)
{
  return openmc::watt_spectrum(a, b, seed);
}
int64_t ref_lower_bound_index(const double* a, int64_t n, double v)
{
  return openmc::lower_bound_index(a, a + n, v);
}
void ref_reflect(double* u, const double* n)
{
  openmc::Position uu {u[0], u[1], u[2]};
  openmc::Position nn {n[0], n[1], n[2]};
  auto r = uu.reflect(nn);
  u[0] = r.x;
  u[1] = r.y;
  u[2] = r.z;
}
}
