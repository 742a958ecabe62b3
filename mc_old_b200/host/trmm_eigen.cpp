// trmm_eigen.cpp — eigen-pairs of the Transition Rate Matrix: what TRMM.exe does after the run (reference TRMM.cpp:10-81).
//
// The reference hands TRM (N x N real, N = G energy groups + 6 precursor groups) to Eigen::EigenSolver and writes the
// eigenvalues ("alpha") and eigenvectors ("phi_mode"), then the same for the adjoint matrix.  There is no Eigen in this
// image, and a post-processing step of a 26 x 26 matrix has no business on the GPU, so this is a small host solver:
//   1. Householder reduction to upper Hessenberg form;
//   2. eigenvalues by the shifted QR iteration in complex arithmetic (Wilkinson shift, Givens rotations, deflation);
//   3. one eigenvector per eigenvalue by inverse iteration on the ORIGINAL matrix (complex LU with partial pivoting),
//      which also polishes nothing and hides nothing: the residual |A v - alpha v| is what the tests check.
// Conventions: eigenvalues sorted by descending real part (the fundamental mode first; conjugate pairs adjacent, the
// member with positive imaginary part first); eigenvectors have unit 2-norm and their largest component real and
// positive.  Eigen's order (position on its Schur form's diagonal) and phase are artefacts of its iteration; the
// reference's own consumer (examples/infinite_GCR_TRMM/plot.py) sorts the eigenvalues before use and its expansion
// coefficients do not depend on the normalisation.
#include "trmm_eigen.h"

#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>

namespace mcbhost {

namespace {

typedef std::complex<double> cplx;
const double EPS = std::numeric_limits<double>::epsilon();

// Diagonal similarity D^-1 A D with powers of two (no rounding) that makes the norms of row i and column i
// comparable: the TRM mixes rates of 1e7 /s (fast groups) with precursor decay constants of 1e-2 /s, and the QR
// iteration's error is relative to the norm of what it is given.
void balance(int n, std::vector<double>& A)
{
    bool again = true;
    while (again) {
        again = false;
        for (int i = 0; i < n; i++) {
            double c = 0.0, r = 0.0;
            for (int j = 0; j < n; j++) if (j != i) { c += std::fabs(A[j * n + i]); r += std::fabs(A[i * n + j]); }
            if (c == 0.0 || r == 0.0) continue;
            double f = 1.0;
            const double s = c + r;
            while (c < r / 2.0) { c *= 2.0; r /= 2.0; f *= 2.0; }
            while (c >= r * 2.0) { c /= 2.0; r *= 2.0; f /= 2.0; }
            if (c + r < 0.95 * s && f != 1.0) {
                again = true;
                for (int j = 0; j < n; j++) A[j * n + i] *= f;
                for (int j = 0; j < n; j++) A[i * n + j] /= f;
            }
        }
    }
}

// A (n x n, row-major) -> upper Hessenberg, similarity by Householder reflectors (eigenvalues only: not accumulated)
void to_hessenberg(int n, std::vector<double>& A)
{
    std::vector<double> v(n);
    for (int k = 0; k + 2 < n; k++) {
        double norm = 0.0;
        for (int i = k + 1; i < n; i++) norm += A[i * n + k] * A[i * n + k];
        norm = std::sqrt(norm);
        if (norm == 0.0) continue;
        const double a = A[(k + 1) * n + k];
        const double alpha = a > 0.0 ? -norm : norm;
        for (int i = 0; i < n; i++) v[i] = 0.0;
        v[k + 1] = a - alpha;
        for (int i = k + 2; i < n; i++) v[i] = A[i * n + k];
        double vv = 0.0;
        for (int i = k + 1; i < n; i++) vv += v[i] * v[i];
        if (vv == 0.0) continue;
        // A <- (I - 2 v v^T / vv) A (I - 2 v v^T / vv)
        for (int j = 0; j < n; j++) {
            double s = 0.0;
            for (int i = k + 1; i < n; i++) s += v[i] * A[i * n + j];
            s *= 2.0 / vv;
            for (int i = k + 1; i < n; i++) A[i * n + j] -= s * v[i];
        }
        for (int i = 0; i < n; i++) {
            double s = 0.0;
            for (int j = k + 1; j < n; j++) s += A[i * n + j] * v[j];
            s *= 2.0 / vv;
            for (int j = k + 1; j < n; j++) A[i * n + j] -= s * v[j];
        }
        for (int i = k + 2; i < n; i++) A[i * n + k] = 0.0;
    }
}

// eigenvalues of a complex upper Hessenberg matrix by the explicitly shifted QR iteration; H is destroyed
bool hessenberg_eigenvalues(int n, std::vector<cplx>& H, std::vector<cplx>& w, std::string& error)
{
    w.assign(n, cplx(0.0, 0.0));
    int hi = n - 1, iter = 0, total = 0;
    std::vector<cplx> cs(n), sn(n);
    while (hi >= 0) {
        // the active block [lo, hi]: lo is the first row below a negligible subdiagonal entry
        int lo = hi;
        while (lo > 0) {
            const double sub = std::abs(H[lo * n + lo - 1]);
            double scale = std::abs(H[lo * n + lo]) + std::abs(H[(lo - 1) * n + lo - 1]);
            if (scale == 0.0) scale = 1.0;
            if (sub <= EPS * scale) { H[lo * n + lo - 1] = 0.0; break; }
            lo--;
        }
        if (lo == hi) { w[hi] = H[hi * n + hi]; hi--; iter = 0; continue; }
        if (++total > 200 * n + 1000) { error = "QR iteration did not converge"; return false; }
        // Wilkinson shift: the eigenvalue of the trailing 2 x 2 block closer to its last diagonal entry
        const cplx a = H[(hi - 1) * n + hi - 1], b = H[(hi - 1) * n + hi], c = H[hi * n + hi - 1], d = H[hi * n + hi];
        cplx mu;
        if (++iter % 11 == 0) mu = d + cplx(std::abs(c.real()) + std::abs(c.imag()), 0.0);  // exceptional shift
        else {
            const cplx tr = a + d, det = a * d - b * c;
            const cplx disc = std::sqrt(tr * tr - 4.0 * det);
            const cplx l1 = 0.5 * (tr + disc), l2 = 0.5 * (tr - disc);
            mu = std::abs(l1 - d) < std::abs(l2 - d) ? l1 : l2;
        }
        // H - mu I = Q R by Givens rotations, then R Q + mu I, on the active block only (eigenvalues only)
        for (int i = lo; i <= hi; i++) H[i * n + i] -= mu;
        for (int k = lo; k < hi; k++) {
            const cplx x = H[k * n + k], y = H[(k + 1) * n + k];
            const double r = std::sqrt(std::norm(x) + std::norm(y));
            if (r == 0.0) { cs[k] = 1.0; sn[k] = 0.0; continue; }
            cs[k] = x / r; sn[k] = y / r;  // [conj(c) conj(s); -s c] applied from the left zeroes y
            for (int j = k; j <= hi; j++) {
                const cplx t1 = H[k * n + j], t2 = H[(k + 1) * n + j];
                H[k * n + j] = std::conj(cs[k]) * t1 + std::conj(sn[k]) * t2;
                H[(k + 1) * n + j] = -sn[k] * t1 + cs[k] * t2;
            }
        }
        for (int k = lo; k < hi; k++) {  // times Q = G_lo^H ... G_{hi-1}^H from the right
            const int last = std::min(k + 2, hi);
            for (int i = lo; i <= last; i++) {
                const cplx t1 = H[i * n + k], t2 = H[i * n + k + 1];
                H[i * n + k] = t1 * cs[k] + t2 * sn[k];
                H[i * n + k + 1] = -t1 * std::conj(sn[k]) + t2 * std::conj(cs[k]);
            }
        }
        for (int i = lo; i <= hi; i++) H[i * n + i] += mu;
    }
    return true;
}

// x <- (A - shift I)^-1 x by LU with partial pivoting; tiny pivots are replaced (inverse iteration wants them)
void solve_shifted(int n, const std::vector<double>& A, cplx shift, double tiny, std::vector<cplx>& x)
{
    std::vector<cplx> M((size_t)n * n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) M[i * n + j] = cplx(A[i * n + j], 0.0) - (i == j ? shift : cplx(0.0, 0.0));
    for (int k = 0; k < n; k++) {
        int piv = k;
        for (int i = k + 1; i < n; i++) if (std::abs(M[i * n + k]) > std::abs(M[piv * n + k])) piv = i;
        if (piv != k) { for (int j = 0; j < n; j++) std::swap(M[k * n + j], M[piv * n + j]); std::swap(x[k], x[piv]); }
        if (std::abs(M[k * n + k]) < tiny) M[k * n + k] = tiny;
        for (int i = k + 1; i < n; i++) {
            const cplx f = M[i * n + k] / M[k * n + k];
            if (f == cplx(0.0, 0.0)) continue;
            for (int j = k + 1; j < n; j++) M[i * n + j] -= f * M[k * n + j];
            x[i] -= f * x[k];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        cplx s = x[i];
        for (int j = i + 1; j < n; j++) s -= M[i * n + j] * x[j];
        x[i] = s / M[i * n + i];
    }
}

void normalise(std::vector<cplx>& x)
{
    double nrm = 0.0;
    size_t big = 0;
    for (size_t i = 0; i < x.size(); i++) { nrm += std::norm(x[i]); if (std::abs(x[i]) > std::abs(x[big])) big = i; }
    nrm = std::sqrt(nrm);
    if (nrm == 0.0) return;
    const cplx phase = std::conj(x[big]) / std::abs(x[big]);  // the largest component becomes real and positive
    for (cplx& v : x) v = v * phase / nrm;
}

}  // namespace

bool eigen_general(int n, const double* A_in, std::vector<std::complex<double>>& w, std::vector<std::complex<double>>& V, std::string& error)
{
    if (n <= 0) { error = "empty matrix"; return false; }
    std::vector<double> A(A_in, A_in + (size_t)n * n);
    double anorm = 0.0;
    for (double v : A) {
        if (!std::isfinite(v)) { error = "matrix has non-finite entries"; return false; }
        anorm = std::max(anorm, std::fabs(v));
    }
    std::vector<double> Hr = A;
    balance(n, Hr);
    to_hessenberg(n, Hr);
    std::vector<cplx> H((size_t)n * n);
    for (size_t i = 0; i < H.size(); i++) H[i] = cplx(Hr[i], 0.0);
    if (!hessenberg_eigenvalues(n, H, w, error)) return false;
    // a real matrix: eigenvalues are real or come in conjugate pairs; tidy what the complex iteration left.  An
    // eigenvalue is paired with the one closest to its conjugate when that one is close indeed (the two were computed
    // independently and agree to rounding); without such a partner a small imaginary part is noise on a real eigenvalue
    std::vector<char> done(n, 0);
    for (int i = 0; i < n; i++) {
        if (done[i]) continue;
        const double size = std::max(std::abs(w[i]), anorm * EPS);
        if (w[i].imag() == 0.0) { done[i] = 1; continue; }
        int best = -1;
        for (int j = 0; j < n; j++) {
            if (j == i || done[j]) continue;
            if (best < 0 || std::abs(w[j] - std::conj(w[i])) < std::abs(w[best] - std::conj(w[i]))) best = j;
        }
        if (best >= 0 && w[best].imag() * w[i].imag() < 0.0 && std::abs(w[best] - std::conj(w[i])) <= 1e-6 * size) {
            const cplx m = 0.5 * (w[i] + std::conj(w[best]));
            w[i] = m; w[best] = std::conj(m);
            done[i] = done[best] = 1;
        } else if (std::fabs(w[i].imag()) <= 1e-6 * size) {
            w[i] = cplx(w[i].real(), 0.0);
            done[i] = 1;
        }  // else: left as computed
    }
    std::vector<int> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
        if (w[a].real() != w[b].real()) return w[a].real() > w[b].real();
        return w[a].imag() > w[b].imag();
    });
    std::vector<cplx> ws(n);
    for (int i = 0; i < n; i++) ws[i] = w[order[i]];
    w = ws;
    // eigenvectors: inverse iteration with the eigenvalue nudged off the spectrum
    V.assign((size_t)n * n, cplx(0.0, 0.0));
    const double tiny = std::max(anorm, std::numeric_limits<double>::min() / EPS) * EPS;
    std::vector<cplx> x(n);
    for (int j = 0; j < n; j++) {
        if (j > 0 && w[j] == std::conj(w[j - 1]) && w[j].imag() != 0.0) {  // the conjugate pair's second member
            for (int i = 0; i < n; i++) V[(size_t)i * n + j] = std::conj(V[(size_t)i * n + j - 1]);
            continue;
        }
        const cplx shift = w[j] + cplx(std::max(std::abs(w[j]), anorm * EPS) * 8.0 * EPS, 0.0);
        for (int i = 0; i < n; i++) x[i] = cplx(1.0 + 0.1 * ((i * 7 + j * 3) % 11), 0.05 * ((i * 5 + j) % 7));
        for (int it = 0; it < 3; it++) {
            solve_shifted(n, A, shift, tiny, x);
            normalise(x);
        }
        if (w[j].imag() == 0.0) for (cplx& v : x) v = cplx(v.real(), 0.0);
        normalise(x);
        for (int i = 0; i < n; i++) V[(size_t)i * n + j] = x[i];
    }
    return true;
}

}  // namespace mcbhost

// ---------------------------------------------------------------------------------------------
// the post-processing step itself
// ---------------------------------------------------------------------------------------------
#include "h5lite.h"

namespace mcbhost {

bool trmm_postprocess(const std::string& file_name, std::string& error)
{
    // I/O directory (TRMM.cpp:13-15): output_TRMM.h5 goes next to the file that was read
    const size_t last = file_name.find_last_of('/');
    const std::string io_dir = last == std::string::npos ? std::string() : file_name.substr(0, last + 1);
    std::vector<uint64_t> dims;
    std::vector<double> TRM, speed_inv;
    if (!h5lite::read_root_f64(file_name, "TRM", dims, TRM, error)) return false;
    if (dims.size() != 2 || dims[0] != dims[1] || dims[0] == 0) { error = file_name + ": TRM is not a square matrix"; return false; }
    const int N = (int)dims[0], J = 6, G = N - J;  // TRMM.cpp:26-28
    if (G < 1) { error = file_name + ": TRM has fewer than 7 rows (6 precursor groups + energy groups)"; return false; }
    if (!h5lite::read_root_f64(file_name, "inverse_speed", dims, speed_inv, error)) return false;
    if (speed_inv.size() < (size_t)G) { error = file_name + ": inverse_speed has fewer entries than TRM has energy groups"; return false; }
    std::vector<std::complex<double>> alpha, phi, alpha_adj, phi_adj;
    if (!eigen_general(N, TRM.data(), alpha, phi, error)) { error = "TRM: " + error; return false; }
    // the adjoint matrix (TRMM.cpp:47-58): rows of the energy groups times 1/v, transpose, rows divided by 1/v again
    std::vector<double> A = TRM;
    for (int i = 0; i < G; i++) for (int j = 0; j < N; j++) A[(size_t)i * N + j] *= speed_inv[i];
    std::vector<double> At((size_t)N * N);
    for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) At[(size_t)i * N + j] = A[(size_t)j * N + i];
    for (int i = 0; i < G; i++) for (int j = 0; j < N; j++) At[(size_t)i * N + j] /= speed_inv[i];
    if (!eigen_general(N, At.data(), alpha_adj, phi_adj, error)) { error = "adjoint TRM: " + error; return false; }
    // output_TRMM.h5 (TRMM.cpp:72-78): vectors as N x 1, matrices N x N (row-major, column = mode), complex = {r, i}
    h5lite::File out;
    auto pairs = [](const std::vector<std::complex<double>>& v) {
        std::vector<double> p(2 * v.size());
        for (size_t i = 0; i < v.size(); i++) { p[2 * i] = v[i].real(); p[2 * i + 1] = v[i].imag(); }
        return p;
    };
    const uint64_t n = (uint64_t)N;
    out.root.dataset_c128("alpha", {n, 1}, pairs(alpha).data());
    out.root.dataset_c128("alpha_adj", {n, 1}, pairs(alpha_adj).data());
    out.root.dataset_c128("phi_mode", {n, n}, pairs(phi).data());
    out.root.dataset_c128("phi_mode_adj", {n, n}, pairs(phi_adj).data());
    return out.write(io_dir + "output_TRMM.h5", error);
}

}  // namespace mcbhost
