/* mcb200_host.h — C entry points of the host-side problem setup (libmcbhost.so).
 *
 * Replaces the reference's `Simulator::Simulator(const std::string io_dir)`
 * (src/simulator/setup.cpp:30-1069, include/simulator.h:144): input.xml +
 * xs_library text files -> the flattened `mcb_problem` that mcb_create() and the
 * CPU oracle consume.  No GPU is needed for anything in this header.
 */
#ifndef MCB200_HOST_H
#define MCB200_HOST_H

#include "mcb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mcbh_deck mcbh_deck;

enum { MCBH_IGNORE_TRMM = 1 }; /* accept <trmm> but build no TRMM tallies (transport + k only) */

/* io_dir: directory holding input.xml (a trailing '/' is added when missing, like Main.cpp:16);
 * xs_dir: directory holding <ZAID>.txt (reference: "./xs_library" relative to the CWD, setup.cpp:326).
 * Returns NULL on error; the message (the reference's own wording where it has one) is in mcbh_last_error(). */
mcbh_deck* mcbh_load_deck(const char* io_dir, const char* xs_dir, int flags);
/* same, from XML text */
mcbh_deck* mcbh_load_deck_string(const char* xml_text, const char* xs_dir, int flags);
void mcbh_free_deck(mcbh_deck* d);
const char* mcbh_last_error(void);

/* the flattened problem; pointers stay valid until mcbh_free_deck. n_sample / n_cycle / n_passive / seed may be
 * overridden through mcbh_set_run (0 keeps the deck's value) — the reference has no such knobs */
const mcb_problem* mcbh_problem(mcbh_deck* d);
void mcbh_set_run(mcbh_deck* d, uint64_t n_sample, uint64_t n_cycle, uint64_t n_passive, uint64_t seed);

/* scalar facts about the problem: out[0..15] = n_sample, n_cycle, n_passive, ksearch, n_nuclides, n_materials,
 * n_surfaces, n_cells, n_estimators, n_tallies, n_sources, entropy_on, n_xs_rows, n_scores, n_filters, trmm_present */
void mcbh_info(mcbh_deck* d, int64_t out[16]);

/* names for reporting (Estimator::report): kind 0 nuclide, 1 material, 2 surface, 3 cell, 4 estimator, 5 score
 * (global score index); NULL when out of range */
const char* mcbh_name(const mcbh_deck* d, int kind, int index);
/* tally layout for reporting (Estimator.cpp:280-295,368-422): out = attach, score_begin, n_scores, filter_begin,
 * n_filters, tally_begin, n_tallies, simulate / out = type, grid_begin, grid_n, size; -1 when out of range */
int mcbh_estimator_info(const mcbh_deck* d, int estimator, int64_t out[8]);
int mcbh_filter_info(const mcbh_deck* d, int filter, int64_t out[4]);
const double* mcbh_filter_grid(const mcbh_deck* d);
const char* mcbh_mode(const mcbh_deck* d);            /* "fixed source" | "k-eigenvalue" */
const char* mcbh_simulation_name(const mcbh_deck* d);
int mcbh_search_cell(const mcbh_deck* d, double x, double y, double z); /* general.cpp:26-34; -1 = lost */

/* Simulator::report (src/simulator/report.cpp:9-52), Estimator::report (src/Estimator.cpp:368-422) and
 * EstimatorK::report (:562-594): writes the reference's output.h5 tree to `path` with the library's own
 * HDF5 writer (host/h5lite.cpp; there is no libhdf5 in the image).  k_cycle / H_cycle: n_cycle values; k_avg / k_uncer:
 * n_active running values (ksearch decks only; pass NULL / 0 otherwise); tallies in the flat order of mcb_get_tallies.
 * Returns 0, or -1 with mcbh_last_error(). */
int mcbh_write_output(mcbh_deck* d, const char* path, uint64_t n_track, const double* k_cycle, const double* H_cycle,
                      int32_t n_cycle, const double* k_avg, const double* k_uncer, int32_t n_active,
                      const double* tally_mean, const double* tally_uncer, int64_t n_tallies);

/* TRM assembly of the TRMM tally set (report.cpp:53-157) from the flat tally means: TRM (G+6)^2 row-major,
 * inverse_speed G, C_initial 6, psi_initial G.  Returns G, or -1 when the deck has no TRMM set. */
int mcbh_trm_assemble(mcbh_deck* d, const double* tally_mean, double* TRM, double* inverse_speed, double* C_initial,
                      double* psi_initial);

/* census times of a time-dependent deck (<tdmc>, setup.cpp:133-169) and the interval lengths as the reference computes
 * them; returns their number (0: not a time-dependent deck).  Either output may be NULL. */
int mcbh_tdmc(const mcbh_deck* d, double* time_out, double* interval_out);

/* The reference's TRMM.exe (TRMM.cpp:10-81), a post-processing step on the host: reads "TRM" and "inverse_speed" from
 * the run's output file `output_h5`, solves the eigen-problems of TRM and of the adjoint matrix and writes alpha,
 * alpha_adj (N x 1), phi_mode, phi_mode_adj (N x N, column = mode) as complex {r, i} datasets to output_TRMM.h5 in the
 * same directory.  Eigenvalues come sorted by descending real part, eigenvectors with unit norm (Eigen::EigenSolver's
 * order and phase are artefacts of its iteration; consumers sort).  Returns 0, or -1 with mcbh_last_error(). */
int mcbh_trmm_postprocess(const char* output_h5);

/* the solver behind it: A n x n row-major; w_pairs 2n doubles (re, im); v_pairs 2 n n doubles, row-major, column j the
 * eigenvector of eigenvalue j.  Returns 0, or -1 with mcbh_last_error(). */
int mcbh_eigen_general(int32_t n, const double* A, double* w_pairs, double* v_pairs);

/* self-check of the device lookup structure (union grid + map + hash, built by the same code mcb_create uses):
 * idx_out[i*Nn + k] = row index the device lookup uses for nuclide k of `material` at E[i]; must equal the
 * reference's binary_search(E, n_E) = #{n_E < E} - 1 (Algorithm.cpp:46-64).  Returns Nn; stats = nU, n_hash,
 * shift, largest hash bin. */
int mcbh_union_indices(mcbh_deck* d, int material, const double* E, int64_t n, int32_t* idx_out, int64_t stats[4]);
/* The crossing shortcut of the walk kernel (test / inspection): out[2 s + (side > 0)] = the cell that search_cell
 * (general.cpp:26-34) is bound to return for any point strictly on that side of surface s, or -1 where it has to search.
 * Returns the number of entries (2 * surfaces), -1 on error. */
int mcbh_cross_neighbors(mcbh_deck* d, int32_t* out, int32_t max_n);


#ifdef __cplusplus
}
#endif
#endif /* MCB200_HOST_H */
