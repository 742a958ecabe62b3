// trmm_eigen.h — eigen-pairs of a general real matrix (the TRMM post-processor, reference TRMM.cpp:31-67)
#ifndef MCB_TRMM_EIGEN_H
#define MCB_TRMM_EIGEN_H

#include <complex>
#include <string>
#include <vector>

namespace mcbhost {

// A: n x n, row-major.  w: the n eigenvalues, by descending real part (conjugate pairs adjacent, positive imaginary part
// first).  V: n x n row-major, column j the unit-norm eigenvector of w[j], largest component real and positive.
bool eigen_general(int n, const double* A, std::vector<std::complex<double>>& w, std::vector<std::complex<double>>& V, std::string& error);

// What TRMM.exe does (reference TRMM.cpp:10-81): reads "TRM" and "inverse_speed" from the run's output file, solves the
// forward and the adjoint eigen-problem and writes alpha, alpha_adj, phi_mode, phi_mode_adj to output_TRMM.h5 next to it.
bool trmm_postprocess(const std::string& output_h5, std::string& error);

}  // namespace mcbhost
#endif
