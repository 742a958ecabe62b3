// TEST INFRASTRUCTURE ONLY (oracle/): empty shadow of the reference's
// include/eigen3-hdf5.hpp, which needs libhdf5.  Only used when compiling
// oracle/_ref (see oracle/build_ref.py).
