// TEST INFRASTRUCTURE ONLY (oracle/): link-time glue for building the
// reference into oracle/_ref without HDF5/Eigen.
// the two H5::PredType statics declared in h5stub/H5Cpp.h.
#include <cstdio>
#include <cstdlib>
#include "simulator.h"

const H5::PredType H5::PredType::NATIVE_DOUBLE(0);
const H5::PredType H5::PredType::NATIVE_ULLONG(1);
