/* hdf5_hl.h STAND-IN (same content as hdf5.h) -- test infrastructure only.
 *
 * The reference's geometry translation units (src/surface.cpp, cell.cpp, lattice.cpp,
 * universe.cpp, geometry.cpp, geometry_aux.cpp, ...) include "hdf5.h" only for the
 * handle types in their to_hdf5() writers; HDF5 itself is not installed in this image.
 * This header declares just those types and type-id objects so that the UNMODIFIED
 * reference sources compile into oracle/_ref/libref_geom.so (recipe: oracle/Makefile),
 * where tests use the reference's own find_cell / distance_to_boundary to pin the
 * oracle and the CUDA path.  Nothing here is HDF5 code; the writers are never called.
 */
#pragma once
#include <cstddef>
#include <cstdint>
typedef int64_t hid_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;
typedef int herr_t;
typedef int htri_t;
extern hid_t H5T_NATIVE_DOUBLE_g, H5T_NATIVE_FLOAT_g, H5T_NATIVE_INT_g, H5T_NATIVE_UINT_g,
  H5T_NATIVE_INT8_g, H5T_NATIVE_INT64_g, H5T_NATIVE_UINT32_g, H5T_NATIVE_UINT64_g, H5T_NATIVE_LONG_g,
  H5T_NATIVE_ULONG_g, H5T_NATIVE_LLONG_g, H5T_NATIVE_ULLONG_g, H5T_NATIVE_SCHAR_g, H5T_NATIVE_UCHAR_g,
  H5T_NATIVE_SHORT_g, H5T_NATIVE_USHORT_g, H5T_C_S1_g;
#define H5T_NATIVE_DOUBLE H5T_NATIVE_DOUBLE_g
#define H5T_NATIVE_FLOAT H5T_NATIVE_FLOAT_g
#define H5T_NATIVE_INT H5T_NATIVE_INT_g
#define H5T_NATIVE_UINT H5T_NATIVE_UINT_g
#define H5T_NATIVE_INT8 H5T_NATIVE_INT8_g
#define H5T_NATIVE_INT64 H5T_NATIVE_INT64_g
#define H5T_NATIVE_UINT32 H5T_NATIVE_UINT32_g
#define H5T_NATIVE_UINT64 H5T_NATIVE_UINT64_g
#define H5T_NATIVE_LONG H5T_NATIVE_LONG_g
#define H5T_NATIVE_ULONG H5T_NATIVE_ULONG_g
#define H5T_NATIVE_LLONG H5T_NATIVE_LLONG_g
#define H5T_NATIVE_ULLONG H5T_NATIVE_ULLONG_g
#define H5T_NATIVE_SCHAR H5T_NATIVE_SCHAR_g
#define H5T_NATIVE_UCHAR H5T_NATIVE_UCHAR_g
#define H5T_NATIVE_SHORT H5T_NATIVE_SHORT_g
#define H5T_NATIVE_USHORT H5T_NATIVE_USHORT_g
#define H5T_C_S1 H5T_C_S1_g
#define H5S_ALL 0
#define H5P_DEFAULT 0
#define H5T_VARIABLE ((size_t)-1)
