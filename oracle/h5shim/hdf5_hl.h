/* hdf5_hl.h of oracle/h5shim -- see hdf5.h in this directory (H5LTpath_valid is declared there). */
#ifndef OMC_H5SHIM_HDF5_HL_H
#define OMC_H5SHIM_HDF5_HL_H
#include "hdf5.h"
#endif
