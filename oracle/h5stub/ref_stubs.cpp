// TEST INFRASTRUCTURE ONLY (oracle/): link-time glue for building the
// reference into oracle/_ref without HDF5/Eigen.
//  * the two H5::PredType statics declared in h5stub/H5Cpp.h;
//  * stand-ins for the two Simulator members defined in the reference's
//    src/simulator/time_dependent.cpp (TDMC mode, out of scope per SURVEY §2),
//    which is left out of the oracle build because it pulls in Eigen.
#include <cstdio>
#include <cstdlib>
#include "simulator.h"

const H5::PredType H5::PredType::NATIVE_DOUBLE(0);
const H5::PredType H5::PredType::NATIVE_ULLONG(1);

void Simulator::time_hit( Particle& P )
{
    std::printf("[oracle/_ref] TDMC mode is not built into the oracle\n");
    std::exit(EXIT_FAILURE);
}
Particle Simulator::forced_decay( const Particle& P, const std::shared_ptr<Nuclide>& N,
                                  const double initial, const double interval, const int p_tdmc )
{
    std::printf("[oracle/_ref] TDMC mode is not built into the oracle\n");
    std::exit(EXIT_FAILURE);
}
