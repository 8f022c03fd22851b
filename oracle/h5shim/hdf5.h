/* hdf5.h of oracle/h5shim -- TEST / BASELINE INFRASTRUCTURE, not HDF5.
 *
 * libhdf5 is not installed in this image and cannot be fetched, and the reference
 * (CMakeLists.txt:170 `find_package(HDF5 REQUIRED ...)`) needs it for its data
 * libraries, summary and statepoint files.  This header declares the ~60 entry
 * points the reference calls (all of them listed by
 * `grep -ohE '\bH5[A-Z]\w*' /root/reference/src /root/reference/include`),
 * and h5shim.cpp implements them over a private container file with the same
 * tree semantics (groups, datasets, attributes, simple dataspaces with hyperslab
 * selections, numeric / fixed-string / compound types with conversion on read and
 * write).  With it the UNMODIFIED reference sources compile and link here into
 * oracle/_ref/libopenmc_ref.so and oracle/_ref/openmc (recipe: oracle/Makefile), which
 * read XML inputs plus data libraries written in that container by oracle/omh5.py.
 * The on-disk bytes are NOT HDF5; the tree (names, shapes, types, attributes) is
 * what the reference's readers and writers see (route 1, option ii).
 */
#ifndef OMC_H5SHIM_HDF5_H
#define OMC_H5SHIM_HDF5_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t hid_t;
typedef unsigned long long hsize_t;
typedef long long hssize_t;
typedef int herr_t;
typedef int htri_t;
typedef long ssize_t_h5;

/* predefined datatypes: fixed handle values (compile-time constants) */
#define H5T_NATIVE_CHAR    ((hid_t)1)
#define H5T_NATIVE_SCHAR   ((hid_t)1)
#define H5T_NATIVE_INT8    ((hid_t)1)
#define H5T_NATIVE_UCHAR   ((hid_t)2)
#define H5T_NATIVE_UINT8   ((hid_t)2)
#define H5T_NATIVE_SHORT   ((hid_t)3)
#define H5T_NATIVE_USHORT  ((hid_t)4)
#define H5T_NATIVE_INT     ((hid_t)5)
#define H5T_NATIVE_INT32   ((hid_t)5)
#define H5T_NATIVE_UINT    ((hid_t)6)
#define H5T_NATIVE_UINT32  ((hid_t)6)
#define H5T_NATIVE_LONG    ((hid_t)7)
#define H5T_NATIVE_INT64   ((hid_t)7)
#define H5T_NATIVE_LLONG   ((hid_t)7)
#define H5T_NATIVE_ULONG   ((hid_t)8)
#define H5T_NATIVE_UINT64  ((hid_t)8)
#define H5T_NATIVE_ULLONG  ((hid_t)8)
#define H5T_NATIVE_FLOAT   ((hid_t)9)
#define H5T_NATIVE_DOUBLE  ((hid_t)10)
#define H5T_C_S1           ((hid_t)11)
#define H5T_VARIABLE       ((size_t)-1)

typedef enum { H5T_INTEGER = 0, H5T_FLOAT = 1, H5T_STRING = 3, H5T_COMPOUND = 6 } H5T_class_t;
typedef enum { H5T_STR_NULLTERM = 0, H5T_STR_NULLPAD = 1, H5T_STR_SPACEPAD = 2 } H5T_str_t;
#define HOFFSET(S, M) (offsetof(S, M))

#define H5S_ALL ((hid_t)0)
#define H5P_DEFAULT ((hid_t)0)
typedef enum { H5S_SCALAR = 0, H5S_SIMPLE = 1 } H5S_class_t;
typedef enum { H5S_SELECT_SET = 0 } H5S_seloper_t;

#define H5F_ACC_RDONLY 0u
#define H5F_ACC_RDWR 1u
#define H5F_ACC_TRUNC 2u
#define H5F_ACC_EXCL 4u

#define H5P_FILE_ACCESS ((hid_t)21)
#define H5P_DATASET_XFER ((hid_t)22)
#define H5FD_MPIO ((hid_t)-2)
typedef enum { H5FD_MPIO_INDEPENDENT = 0, H5FD_MPIO_COLLECTIVE = 1 } H5FD_mpio_xfer_t;

typedef enum { H5I_BADID = -1, H5I_FILE = 1, H5I_GROUP, H5I_DATATYPE, H5I_DATASPACE, H5I_DATASET,
  H5I_ATTR, H5I_GENPROP_LST } H5I_type_t;
typedef enum { H5_INDEX_NAME = 0, H5_INDEX_CRT_ORDER = 1 } H5_index_t;
typedef enum { H5_ITER_INC = 0, H5_ITER_DEC = 1, H5_ITER_NATIVE = 2 } H5_iter_order_t;
typedef enum { H5O_TYPE_UNKNOWN = -1, H5O_TYPE_GROUP = 0, H5O_TYPE_DATASET = 1,
  H5O_TYPE_NAMED_DATATYPE = 2 } H5O_type_t;
typedef struct { hsize_t nlinks; } H5G_info_t;
typedef struct { H5O_type_t type; } H5O_info_t;

/* files */
hid_t H5Fcreate(const char* name, unsigned flags, hid_t fcpl, hid_t fapl);
hid_t H5Fopen(const char* name, unsigned flags, hid_t fapl);
herr_t H5Fclose(hid_t file);
hid_t H5Fget_access_plist(hid_t file);
/* groups, objects, links */
hid_t H5Gcreate(hid_t loc, const char* name, hid_t lcpl, hid_t gcpl, hid_t gapl);
hid_t H5Gopen(hid_t loc, const char* name, hid_t gapl);
herr_t H5Gclose(hid_t group);
herr_t H5Gget_info(hid_t group, H5G_info_t* info);
hid_t H5Oopen(hid_t loc, const char* name, hid_t lapl);
herr_t H5Oclose(hid_t obj);
herr_t H5Oget_info_by_idx(hid_t loc, const char* group_name, H5_index_t idx, H5_iter_order_t order,
  hsize_t n, H5O_info_t* info, hid_t lapl);
long H5Lget_name_by_idx(hid_t loc, const char* group_name, H5_index_t idx, H5_iter_order_t order,
  hsize_t n, char* name, size_t size, hid_t lapl);
htri_t H5LTpath_valid(hid_t loc, const char* path, int check_object_valid);
/* identifiers */
long H5Iget_name(hid_t id, char* name, size_t size);
H5I_type_t H5Iget_type(hid_t id);
hid_t H5Iget_file_id(hid_t id);
/* datasets */
hid_t H5Dcreate(hid_t loc, const char* name, hid_t type, hid_t space, hid_t lcpl, hid_t dcpl, hid_t dapl);
hid_t H5Dopen(hid_t loc, const char* name, hid_t dapl);
herr_t H5Dclose(hid_t dset);
hid_t H5Dget_space(hid_t dset);
hid_t H5Dget_type(hid_t dset);
herr_t H5Dread(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t plist, void* buf);
herr_t H5Dwrite(hid_t dset, hid_t mem_type, hid_t mem_space, hid_t file_space, hid_t plist, const void* buf);
/* attributes */
htri_t H5Aexists_by_name(hid_t loc, const char* obj_name, const char* attr_name, hid_t lapl);
hid_t H5Aopen(hid_t obj, const char* name, hid_t aapl);
hid_t H5Acreate(hid_t obj, const char* name, hid_t type, hid_t space, hid_t acpl, hid_t aapl);
herr_t H5Aclose(hid_t attr);
hid_t H5Aget_space(hid_t attr);
hid_t H5Aget_type(hid_t attr);
herr_t H5Aread(hid_t attr, hid_t mem_type, void* buf);
herr_t H5Awrite(hid_t attr, hid_t mem_type, const void* buf);
/* dataspaces */
hid_t H5Screate(H5S_class_t cls);
hid_t H5Screate_simple(int rank, const hsize_t* dims, const hsize_t* maxdims);
herr_t H5Sclose(hid_t space);
int H5Sget_simple_extent_ndims(hid_t space);
int H5Sget_simple_extent_dims(hid_t space, hsize_t* dims, hsize_t* maxdims);
herr_t H5Sselect_hyperslab(hid_t space, H5S_seloper_t op, const hsize_t* start, const hsize_t* stride,
  const hsize_t* count, const hsize_t* block);
herr_t H5Soffset_simple(hid_t space, const hssize_t* offset);
/* datatypes */
hid_t H5Tcreate(H5T_class_t cls, size_t size);
hid_t H5Tcopy(hid_t type);
herr_t H5Tclose(hid_t type);
herr_t H5Tinsert(hid_t type, const char* name, size_t offset, hid_t member);
herr_t H5Tset_size(hid_t type, size_t size);
size_t H5Tget_size(hid_t type);
herr_t H5Tset_strpad(hid_t type, H5T_str_t pad);
int H5Tget_nmembers(hid_t type);
char* H5Tget_member_name(hid_t type, unsigned idx);   /* malloc'ed: release with H5free_memory / free */
herr_t H5free_memory(void* p);
/* property lists */
hid_t H5Pcreate(hid_t cls);
herr_t H5Pclose(hid_t plist);
hid_t H5Pget_driver(hid_t plist);
herr_t H5Pset_dxpl_mpio(hid_t plist, H5FD_mpio_xfer_t mode);

#ifdef __cplusplus
}
#endif
#endif
