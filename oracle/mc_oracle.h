/* mc_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the reference's particle-history transport loop
 * (ilhamv/MC-old, `Simulator::start()` and below).  It exists to CHECK the CUDA
 * path; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it.  The product never links or calls it.
 *
 * Parity status: PINNED.  In rng_mode 0 the oracle reproduces the compiled
 * reference (oracle/_ref/MC_ref, built from /root/reference by
 * oracle/build_ref.py) bit for bit — k per cycle, entropy, tallies, Ntrack — on
 * the decks tests/test_oracle_golden.py runs; golden outputs of the reference
 * are committed under tests/golden/.
 */
#ifndef MC_ORACLE_H
#define MC_ORACLE_H

#include <stdint.h>

#include "mcb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mco_ctx mco_ctx;

enum {
    MCO_RNG_GLOBAL = 0,  /* the reference as it is: ONE global LCG stream, seed 1 (Random.cpp:105) */
    MCO_RNG_HISTORY = 1  /* per-history streams (the reference's unused RN_init_particle design,
                            Random.cpp:196-204): history nps = cycle*Nsample + h starts at skip(seed0, nps*152917) */
};
enum {
    MCO_PICK_CDF = 0,    /* SourceBank::set_up/get_source: p[i]=p[i-1]+1/N, binary_search (Source.cpp:42-46,65-72) */
    MCO_PICK_FLOOR = 1   /* j = floor(xi*N): the rule the GPU uses (no O(N) serial prefix of rounding errors) */
};

typedef struct mco_cycle_result {
    double k_cycle, k_avg, k_uncer, H;
    double k_sum_C, k_sum_TL, k_sq_C, k_sq_TL, H_sum;
    uint64_t n_sites, n_tracks, n_collisions, n_histories, n_draws;
} mco_cycle_result;

mco_ctx* mco_create(const mcb_problem* problem, int rng_mode, int pick_mode);
void mco_destroy(mco_ctx* c);
/* histories [begin, begin+count) of every cycle are run by this ctx (default: all); MCO_RNG_HISTORY only */
void mco_set_shard(mco_ctx* c, uint64_t begin, uint64_t count);

/* one cycle of handler.cpp:14-44 on one rank */
int mco_run_cycle(mco_ctx* c, mco_cycle_result* out);
/* split phases for sharded runs: transport the owned histories, exchange, close the cycle with global sums */
int mco_transport_cycle(mco_ctx* c);
void mco_get_partials(const mco_ctx* c, double* sums5, uint64_t* counts4); /* {kC,kTL,kC2,kTL2,H}, {sites,tracks,coll,hist} */
/* EstimatorK::k_C / k_TL of every owned history of the last cycle at end_history (Estimator.cpp:514-525) */
int64_t mco_get_history_k(const mco_ctx* c, double* kC, double* kTL);
int64_t mco_bank_size(const mco_ctx* c);
void mco_get_bank(const mco_ctx* c, double* sites8, int32_t* cells);       /* x,y,z,u,v,w,E,t per site, banking order */
void mco_set_source_bank(mco_ctx* c, const double* sites8, const int32_t* cells, int64_t n);
void mco_get_tally_partials(const mco_ctx* c, double* sum, double* squared);
void mco_close_cycle(mco_ctx* c, const double* sums5, const uint64_t* counts4, const double* tally_sum,
                     const double* tally_squared, mco_cycle_result* out);
/* handler.cpp:47 + read-back */
void mco_end_simulation(mco_ctx* c);
void mco_get_tallies(const mco_ctx* c, double* mean, double* uncer);
double mco_get_k(const mco_ctx* c);
uint64_t mco_get_seed(const mco_ctx* c);

/* function-level entry points (same arithmetic the loop uses) */
/* Material::Sigma{T,S,C,F}, nuSigmaF (Material.cpp:18-65) */
void mco_xs_lookup(const mcb_problem* p, int material, const double* E, int64_t n, double* out5);
/* Material::nuclide_scatter (0) / nuclide_nufission (1) (Material.cpp:106-125); global nuclide index or -1 */
void mco_select_channel(const mcb_problem* p, int material, int kind, const double* E, const double* xi, int64_t n,
                        int32_t* nuclide);
/* Nuclide::beta (Nuclide.cpp:74-77) */
void mco_beta(const mcb_problem* p, int nuclide, const double* E, int64_t n, double* out);
int mco_binary_search(double x, const double* v, int n);                     /* Algorithm.cpp:46-64 */
double mco_interpolate(double x, double x1, double x2, double y1, double y2); /* Algorithm.cpp:103-105 */
double mco_geometry_quad(double a, double b, double c);                       /* Algorithm.cpp:16-38 */
void mco_scatter_direction(const double* dir, double mu0, double xi, double* out); /* Algorithm.cpp:67-101 */
uint64_t mco_lcg_next(uint64_t seed);                                         /* Random.cpp:121-126 */
uint64_t mco_lcg_skip(uint64_t seed, uint64_t n);                             /* Random.cpp:130-149 */
double mco_surface_eval(const mcb_problem* p, int s, const double* pos);      /* Geometry.cpp:29-69 */
double mco_surface_distance(const mcb_problem* p, int s, const double* pos, const double* dir); /* :76-188 */
void mco_surface_reflect(const mcb_problem* p, int s, double* dir);           /* :195-222 */
int mco_search_cell(const mcb_problem* p, const double* pos);                 /* general.cpp:13-34 */
/* surface_intersect (general.cpp:54-67): returns surface index or -1, distance through *dist */
int mco_surface_intersect(const mcb_problem* p, int cell, const double* pos, const double* dir, double* dist);
/* ReactionScatter::sample (Reaction.cpp:27-118) drawing from the LCG state *seed; io = {u,v,w,E} -> {u,v,w,E,speed} */
void mco_scatter_sample(const mcb_problem* p, int nuclide, uint64_t* seed, double* io5);
/* DistributionWatt::sample (Distribution.cpp:34-73) */
double mco_watt_sample(const mcb_problem* p, int nuclide, uint64_t* seed, double E);
double mco_speed_of_energy(double E);                                         /* Particle.cpp:42-48 */
double mco_energy_of_speed(double v);                                         /* Particle.cpp:49-56 */

#ifdef __cplusplus
}
#endif
#endif
