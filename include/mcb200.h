/* mcb200.h — C-ABI of the B200-native particle-history transport loop.
 *
 * Drop-in boundary for ilhamv/MC-old's `Simulator::start()` hot path (reference
 * src/simulator/handler.cpp:11-48 and everything below it).  The
 * reference has no FFI: its seam is the C++ class `Simulator(dir)` / `start()` /
 * `report(dir)` (include/simulator.h:144-149).  `setup.cpp` (host) builds an
 * object graph, `start()` consumes it, `report.cpp` (host) drains the tallies.
 * Here the object graph is flattened by the host into the POD `mcb_problem`
 * below, `start()` becomes `mcb_run_cycle()` on the GPU, and the tallies come
 * back through `mcb_get_tallies()` / `mcb_cycle_result`.
 *
 * Everything is `extern "C"`, plain pointers and sizes; no C++ or torch types.
 * All functions return 0 on success, a negative `mcb_status` otherwise;
 * `mcb_last_error()` gives the message (the reference prints a message and
 * calls exit(EXIT_FAILURE), e.g. src/simulator/general.cpp:31-33 — the host
 * wrapper keeps that behaviour, the library itself never exits the process).
 *
 * The same `mcb_problem` is consumed by the CPU oracle (oracle/mc_oracle.c,
 * test infrastructure only) so that both sides see identical inputs.
 */
#ifndef MCB200_H
#define MCB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCB_ABI_VERSION 4
#define MCB_MAX_MAT_NUCLIDES 8 /* nuclides per material handled in registers */
#define MCB_XS_ROW 6           /* doubles per xs row: E, sigma_s, sigma_c, sigma_f, nu, beta */

typedef enum mcb_status {
    MCB_OK = 0,
    MCB_ERR_ARG = -1,      /* bad argument / unsupported deck feature */
    MCB_ERR_CUDA = -2,     /* CUDA runtime error or no device */
    MCB_ERR_LOST = -3,     /* a particle was lost (reference: "[WARNING] A particle is lost", general.cpp:31) */
    MCB_ERR_CAPACITY = -4, /* a device bank overflowed */
    MCB_ERR_COMM = -5      /* NCCL error */
} mcb_status;

/* ---- geometry (reference include/Geometry.h:44-223, src/Geometry.cpp) ---- */
enum { MCB_SURF_PLANE_X = 0, MCB_SURF_PLANE_Y, MCB_SURF_PLANE_Z, MCB_SURF_PLANE,
       MCB_SURF_SPHERE, MCB_SURF_CYL_X, MCB_SURF_CYL_Y, MCB_SURF_CYL_Z };
enum { MCB_BC_TRANSMISSION = 0, MCB_BC_REFLECTIVE = 1, MCB_BC_VACUUM = -1 };

typedef struct mcb_surface {
    int32_t type, bc;
    /* plane_x/y/z: p[0]=location.  plane: p[0..3]=a,b,c,d, p[4..6]=2a/|n|^2,2b/|n|^2,2c/|n|^2 (Geometry.cpp:14-23).
     * sphere: p[0..2]=centre, p[3]=r, p[4]=r*r.  cylinder_x: p[0]=y0,p[1]=z0,p[2]=r,p[3]=r*r (y: x0,z0; z: x0,y0). */
    double p[8];
} mcb_surface;

typedef struct mcb_cell {
    int32_t surf_begin, surf_end; /* range in cell_surface[] / cell_sense[] */
    int32_t material;             /* index or -1 for void (Geometry.h:195-220) */
    int32_t reserved;
    double importance;
} mcb_cell;

/* ---- nuclear data (reference src/simulator/setup.cpp:310-467, xs_library/README.txt) ---- */
typedef struct mcb_nuclide {
    int64_t row_begin;  /* first row in xs_rows (rows of MCB_XS_ROW doubles) */
    int32_t n_rows;     /* 1 for a constant-xs (user-defined) nuclide */
    int32_t has_delayed;/* <ZAID>D.txt was loaded (first sigma_f != 0, setup.cpp:377) */
    double A;
    double watt_a[3], watt_b[3], watt_g[3]; /* g = sqrt(C*C-1)+C, C = 1+a*b/8 (Distribution.cpp:14-23) */
    double lambda[6], fraction[6];
    /* delayed-neutron spectra (setup.cpp:377-412): energies then 6 truncated CDFs, all in delayed_data[] */
    int32_t chid_E_begin, chid_E_n;
    int32_t chid_cdf_begin[6], chid_cdf_n[6];
} mcb_nuclide;

/* ---- sources and distributions (setup.cpp:174-305,1020-1066; src/Distribution.cpp) ---- */
enum { MCB_DIST_DELTA = 0, MCB_DIST_UNIFORM = 1, MCB_DIST_WATT = 2 };
enum { MCB_DIR_DELTA = 0, MCB_DIR_ISOTROPIC = 1, MCB_DIR_XYZ = 2 };
/* source shapes.  POINT: SourcePoint (Source.cpp:20-24).  DISK_Z: the <disk_z x y z r> element of
 * examples/sphere_detection/input.xml:106 — the deck behind the reference's MCNP6 integral test
 * (test/test_integral_Simulator.cpp:10-19) — which the reference's own loader rejects (setup.cpp:1051-1063):
 * positions uniform on the disk of radius r around (x, y, z) in the plane z = const: after energy and direction,
 * rho = r sqrt(xi1), phi = 2 pi xi2, (x + rho cos phi, y + rho sin phi, z); the cell is searched per particle. */
enum { MCB_SRC_POINT = 0, MCB_SRC_DISK_Z = 1 };
typedef struct mcb_dist1 {
    int32_t kind, reserved;
    double a, b; /* delta: a=value; uniform: a, b */
    double watt_a[3], watt_b[3], watt_g[3];
} mcb_dist1;
typedef struct mcb_source {
    double pos[3];
    int32_t cell;
    int32_t dir_kind;
    double dir[3];        /* MCB_DIR_DELTA */
    mcb_dist1 dir_xyz[3]; /* MCB_DIR_XYZ */
    mcb_dist1 energy;
    double prob;          /* parsed and ignored at sampling, like Source.cpp:42-46 */
    int32_t kind, reserved; /* MCB_SRC_*; cell = the cell of the centre for a disk */
    double radius;        /* MCB_SRC_DISK_Z */
} mcb_source;

/* ---- estimators (include/Estimator.h, src/Estimator.cpp, setup.cpp:637-805) ---- */
enum { MCB_KERNEL_NEUTRON = 0,  /* w            (Estimator.cpp:17-20) */
       MCB_KERNEL_TRACK = 1,    /* w*l          (:22-25) */
       MCB_KERNEL_COLLISION = 2,/* w/SigmaT(E)  (:27-30) */
       MCB_KERNEL_VELOCITY = 3, /* w*v          (:32-35) */
       MCB_KERNEL_TRACK_VELOCITY = 4 /* w*l*v   (:37-41) */ };
enum { MCB_SCORE_FLUX = 0, MCB_SCORE_ABSORPTION, MCB_SCORE_SCATTER, MCB_SCORE_CAPTURE, MCB_SCORE_FISSION,
       MCB_SCORE_NU_FISSION, MCB_SCORE_TOTAL, MCB_SCORE_INVERSE_VELOCITY,
       MCB_SCORE_SCATTER_OLD, MCB_SCORE_NU_FISSION_OLD, MCB_SCORE_NU_FISSION_PROMPT_OLD,
       MCB_SCORE_NU_FISSION_DELAYED_OLD, MCB_SCORE_NU_FISSION_DELAYED_DECAY_OLD };
enum { MCB_FILTER_SURFACE = 0, MCB_FILTER_CELL, MCB_FILTER_ENERGY, MCB_FILTER_ENERGY_OLD, MCB_FILTER_TIME,
       MCB_FILTER_TDMC /* FilterTDMC (Estimator.cpp:247-263): bin = the census the particle has just reached, else none */ };
enum { MCB_ATTACH_SURFACE = 0, MCB_ATTACH_CELL_TL = 1, MCB_ATTACH_CELL_C = 2 };
/* simulate-then-score estimators of the TRMM tally set (Estimator.cpp:441-482): a copy of the particle scatters /
 * fissions before the generic scoring.  MCB_SIM_FISSION_DELAYED + g = delayed group g (0..5) */
enum { MCB_SIM_NONE = 0, MCB_SIM_SCATTER = 1, MCB_SIM_FISSION = 2, MCB_SIM_FISSION_PROMPT = 3, MCB_SIM_FISSION_DELAYED = 4 };

typedef struct mcb_score {
    int32_t score, kernel, group, reserved;
    char name[48];
} mcb_score;
typedef struct mcb_filter {
    int32_t type;
    int32_t grid_begin, grid_n; /* range in filter_grid[] */
    int32_t size;               /* bins: grid_n for surface/cell, grid_n-1 for energy/time (Estimator.h:297) */
} mcb_filter;
typedef struct mcb_estimator {
    int32_t attach;
    int32_t score_begin, n_scores;
    int32_t filter_begin, n_filters; /* filter 0 is the surface/cell id filter */
    int32_t tally_begin, n_tallies;  /* range in the global tally vector; layout [score][f1][f2].. row-major */
    int32_t simulate;                /* MCB_SIM_* */
    char name[64];
} mcb_estimator;

/* ---- the flattened problem ---- */
typedef struct mcb_problem {
    int32_t abi_version;
    /* run control (setup.cpp:42-128, simulator.h:90-91,100-112) */
    int32_t ksearch;
    uint64_t n_sample, n_cycle, n_passive;
    double wr, ws;
    uint64_t seed;   /* base seed of the 63-bit LCG; the reference has no seed input and uses 1 (Random.cpp:97) */
    /* nuclear data */
    int32_t n_nuclides;
    int32_t n_materials;
    const mcb_nuclide* nuclides;
    const double* xs_rows;       /* n_xs_rows x MCB_XS_ROW */
    int64_t n_xs_rows;
    const double* delayed_data;
    int64_t n_delayed_data;
    const int32_t* mat_begin;    /* n_materials+1 offsets into mat_nuclide/mat_density (deck order = summation order) */
    const int32_t* mat_nuclide;
    const double* mat_density;
    /* geometry */
    int32_t n_surfaces, n_cells;
    const mcb_surface* surfaces;
    const mcb_cell* cells;
    const int32_t* cell_surface;
    const int32_t* cell_sense;
    int32_t n_cell_surface;
    /* sources */
    int32_t n_sources;
    const mcb_source* sources;
    /* estimators */
    int32_t n_estimators, n_scores, n_filters, n_filter_grid;
    const mcb_estimator* estimators;
    const mcb_score* scores;
    const mcb_filter* filters;
    const double* filter_grid;
    int64_t n_tallies;
    /* Shannon entropy mesh (setup.cpp:82-115; include/Entropy.h) ; entropy_on = 0 -> EntropyNone */
    int32_t entropy_on;
    int32_t entropy_n[3];        /* number of grid points per axis (bins + 1) */
    const double* entropy_grid;  /* x grid, then y grid, then z grid */
    /* particle comb (setup.cpp:57-67, population_control.cpp:55-84): after every random walk, a history whose particle
     * bank holds bank_max or more waiting particles is combed down to `teeth` particles of equal weight */
    int32_t comb_on, comb_bank_max, comb_teeth, reserved2;
    /* time-dependent mode (<tdmc>, setup.cpp:133-169; time_dependent.cpp; general.cpp:187-195; fixed_source.cpp:25-40):
     * particles stop at the census times tdmc_time[0..n_tdmc) (a particle carries the index of the next one) and die at
     * the last; delayed neutrons are forced to decay once in every remaining interval.  tdmc_interval[j] is the length
     * the reference uses for interval j — it is computed from the previous INTERVAL, not the previous time
     * (setup.cpp:146), and is passed through as the reference computes it */
    int32_t tdmc_on, n_tdmc;
    const double* tdmc_time;
    const double* tdmc_interval;
} mcb_problem;

/* ---- per-process device context ---- */
typedef struct mcb_ctx mcb_ctx;

typedef struct mcb_config {
    int32_t device;            /* CUDA device ordinal */
    int32_t rank, world;       /* history sharding: this process owns histories [n*rank/world, n*(rank+1)/world) */
    int32_t reserved;          /* flags: 1 = time every stage launch with CUDA events; 2 = event-queue mode, one kernel per event type */
    int64_t bank_capacity;     /* particle slots in flight on this GPU (0 = choose) */
    int64_t site_capacity;     /* fission sites this GPU can bank per cycle (0 = choose) */
    void* stream;              /* cudaStream_t to launch on (NULL = library-owned stream) */
} mcb_config;

typedef struct mcb_cycle_result {
    double k_cycle;            /* (mean_C + mean_TL)/2            (Estimator.cpp:526-535) */
    double k_avg, k_uncer;     /* running active-cycle values      (:539-549); 0 on passive cycles */
    double H;                  /* H_sum / Nsample (per-history entropy) */
    double H_cycle_conventional; /* extra: entropy of the whole cycle's fission source */
    double k_sum_C, k_sum_TL, k_sq_C, k_sq_TL; /* global sums over histories */
    uint64_t n_sites;          /* global fission-bank size produced by this cycle */
    uint64_t n_histories, n_tracks, n_collisions, n_lookups, n_crossings; /* global counts for this cycle */
    double ms_transport, ms_exchange; /* device time of the transport loop / the bank exchange+reductions */
    int32_t n_iterations;      /* event-loop iterations */
    int32_t lost;              /* particles lost on this rank */
    uint64_t n_kernel_launches;/* kernels this library launched on this rank during the cycle */
} mcb_cycle_result;

typedef struct mcb_stage_times { /* accumulated CUDA-event time per kernel class since mcb_reset_stage_times */
    double ms_source, ms_lookup, ms_flight, ms_cross, ms_collide, ms_closeout, ms_bank;
    uint64_t n_source, n_lookup, n_flight, n_cross, n_collide, n_closeout, n_bank; /* launches */
    uint64_t units_lookup;     /* particles looked up (for the xs roofline) */
    double ms_finish;          /* tail kernel: the last few particles of a batch followed to the end in registers */
    uint64_t n_finish;
    double ms_step;            /* history-walk kernel (k_walk): all events of a particle chained in registers */
    uint64_t n_step;
} mcb_stage_times;

/* lifecycle; replaces the object graph hand-over setup.cpp -> start() */
int mcb_create(const mcb_problem* problem, const mcb_config* config, mcb_ctx** out);
void mcb_destroy(mcb_ctx* ctx);
const char* mcb_last_error(const mcb_ctx* ctx); /* ctx may be NULL for a failed mcb_create */
int mcb_device_count(void);
/* Starts the CUDA driver and the primary context of `device` (what the first mcb_create would otherwise wait for: 0.6-1.6 s
 * on the boxes measured).  A host program calls it on a second thread while it parses the deck; MCB.exe does. */
int mcb_warm_up(int device);

/* multi-GPU plumbing: one process per GPU; id is an opaque 128-byte NCCL unique id made on rank 0 */
int mcb_comm_unique_id(char id[128]);
int mcb_comm_init(mcb_ctx* ctx, const char id[128]);

/* Simulator::start() cycle body (handler.cpp:14-44): source resampling, histories, close-outs, k update */
int mcb_run_cycle(mcb_ctx* ctx, mcb_cycle_result* out);
/* end_simulation (handler.cpp:47) + read-back in the reference's flat order (Estimator.cpp:361-367,423-426) */
int mcb_get_tallies(mcb_ctx* ctx, double* mean, double* uncer, int64_t n);
double mcb_get_k(const mcb_ctx* ctx);
void mcb_set_k(mcb_ctx* ctx, double k);
int mcb_get_stage_times(mcb_ctx* ctx, mcb_stage_times* out);
void mcb_reset_stage_times(mcb_ctx* ctx);
/* per-launch CUDA-event timing of the stage kernels on/off (off by default: the events cost a few percent) */
void mcb_set_stage_timing(mcb_ctx* ctx, int on);
/* fission bank of the last cycle on this rank, canonical order: out = n x 8 doubles (x,y,z,u,v,w,E,t), cells = n */
int64_t mcb_get_fission_bank(mcb_ctx* ctx, double* out, int32_t* cells, int64_t max_n);

/* the global source bank the next cycle will sample (rank-order concatenation of the ranks' banks, gathered on
 * demand), same layout; identical on every rank */
int64_t mcb_get_source_bank(mcb_ctx* ctx, double* out, int32_t* cells, int64_t max_n);
/* source bank of the next cycle from HOST memory (n x 8 doubles x,y,z,u,v,w,E,t + n cells), replacing the bank the
 * last cycle produced: the host-buffer form of `Sbank = Fbank` (handler.cpp:16).  With world > 1 every rank passes
 * the same global bank. */
int mcb_set_source_bank(mcb_ctx* ctx, const double* sites8, const int32_t* cells, int64_t n);
/* One generation with the source bank in HOST buffers on both sides: equivalent to mcb_set_source_bank(in) +
 * mcb_run_cycle + mcb_get_source_bank(out), the host-buffer form of one pass of the cycle body (handler.cpp:14-44 with
 * Sbank / Fbank living in host RAM like the reference's).  On one GPU the three steps are pipelined: the draws are
 * sorted by site index, the bank comes in over PCIe in pieces on a copy stream and the histories that drew from the
 * pieces already there are walked meanwhile; the new bank goes back piece by piece.  Pinned buffers make the copies
 * asynchronous.  *n_out = sites written (at most max_out).
 * Several GPUs (world > 1, banks peer-mapped): the host bank is one array of which rank r owns the slice
 * [n r / W, n (r + 1) / W): every rank passes pointers to the WHOLE arrays (same layout on every rank) but reads only its
 * slice of `in` and writes only slice [n' r / W, n' (r + 1) / W) of the new bank (n' sites) into `out`, at its global
 * offset; *n_out = n'.  Per rank 1/W of the bytes cross the host boundary. */
int mcb_run_cycle_host(mcb_ctx* ctx, const double* in_sites8, const int32_t* in_cells, int64_t n_in, double* out_sites8,
                       int32_t* out_cells, int64_t max_out, int64_t* n_out, mcb_cycle_result* out);
/* per-history k scores of the last cycle on this rank, shard-local history order (EstimatorK::k_C / k_TL at
 * end_history, Estimator.cpp:514-525): parity tests compare them with the oracle history by history */
int64_t mcb_get_history_k(mcb_ctx* ctx, double* kC, double* kTL, int64_t max_n);

/* parity / bench entry points on HOST buffers (H2D + kernel + D2H inside) */
/* Material::Sigma{T,S,C,F}, nuSigmaF (Material.cpp:18-65): out5 = n x {SigmaT,SigmaS,SigmaC,SigmaF,nuSigmaF} */
int mcb_xs_lookup_batch(mcb_ctx* ctx, int32_t material, const double* E, int64_t n, double* out5);
/* same kernel on DEVICE buffers already resident; returns device ms through *ms if non-NULL */
int mcb_xs_lookup_device(mcb_ctx* ctx, int32_t material, const double* dE, int64_t n, double* dout5, float* ms);
/* Material::nuclide_scatter (kind 0) / nuclide_nufission (kind 1) (Material.cpp:106-125): nuclide index or -1 */
int mcb_select_channel_batch(mcb_ctx* ctx, int32_t material, int32_t kind, const double* E, const double* xi,
                             int64_t n, int32_t* nuclide);
/* Nuclide::beta (Nuclide.cpp:74-77) of the material's local_nuclide-th nuclide */
int mcb_beta_batch(mcb_ctx* ctx, int32_t material, int32_t local_nuclide, const double* E, int64_t n, double* out);
/* Urand stream of history nps (Random.cpp:121-149,196-204): seeds_out[n*ndraw] raw 63-bit states after each draw */
int mcb_rng_batch(mcb_ctx* ctx, const uint64_t* nps, int64_t n, int32_t ndraw, uint64_t* seeds_out);
/* surface_intersect + Surface::eval (general.cpp:54-67, Geometry.cpp): per particle in cell[i]:
 * out3 = {distance, (double)surface index or -1, eval of that surface at pos} */
int mcb_geometry_batch(mcb_ctx* ctx, const int32_t* cell, const double* pos3, const double* dir3, int64_t n,
                       double* out3);
/* search_cell (general.cpp:13-34): cell index or -1 */
int mcb_search_cell_batch(mcb_ctx* ctx, const double* pos3, int64_t n, int32_t* cell);
/* ReactionScatter::sample (Reaction.cpp:27-118) with the xi stream of history nps[i]:
 * io5 = {u,v,w,E,-} in, {u,v,w,E,speed} out */
int mcb_scatter_batch(mcb_ctx* ctx, int32_t nuclide, const uint64_t* nps, int64_t n, double* io5);
/* DistributionWatt::sample (Distribution.cpp:34-73) for nuclide at incident E with the stream of nps[i] */
int mcb_watt_batch(mcb_ctx* ctx, int32_t nuclide, const uint64_t* nps, const double* E, int64_t n, double* Eout);

/* The kernels divide through a reciprocal shared between quotients of one denominator (interpolation weights, vector
 * normalisations; csrc/mcb_physics.h): out_shared[i] = that form of a[i] / b[i], out_plain[i] = the compiler's IEEE
 * division.  Cross sections are bit-exact against the reference's x86 divisions only if the two agree bit for bit. */
int mcb_division_batch(mcb_ctx* ctx, const double* a, const double* b, int64_t n, double* out_shared, double* out_plain);

/* launch shape of the walk kernel (the dominant kernel) as it runs on this device: out4 = {registers per thread, grid,
 * block, dynamic shared memory}; scoring = 1 for the instance of tally-scoring cycles.  bench.py checks the committed
 * ncu capture against it before quoting the capture's DRAM traffic. */
int mcb_walk_launch_info(mcb_ctx* ctx, int32_t scoring, int32_t out4[4]);

/* history sharding rule shared by every rank: first history and count owned by `rank` */
void mcb_shard_range(uint64_t n, int32_t rank, int32_t world, uint64_t* begin, uint64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* MCB200_H */
