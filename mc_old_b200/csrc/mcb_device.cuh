// mcb_device.cuh — device-side data layout and the per-particle physics of the transport loop.
//
// Everything here is compiled with -fmad=false: the x86-64 reference build has no fused
// multiply-add, and cross sections / channel selection are bit-exact against it.
#ifndef MCB_DEVICE_CUH
#define MCB_DEVICE_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "mcb200.h"
#include "mcb_physics.h"
#include "mcb_tables.h"

// ---------------------------------------------------------------------------------------------
// flattened problem in device memory (built once by mcb_create; read-only afterwards)
// ---------------------------------------------------------------------------------------------
struct DevMaterial {
    int32_t nuc_begin, n_nuc;    // range in mat_nuclide / mat_density (deck order = summation order)
    int32_t nU, n_hash, shift, hstride;
    int64_t key_min;
    const double* U;             // union grid
    const int32_t* map;          // nU x n_nuc
    const int32_t* hash;         // n_hash + 1
    const int32_t* hrec;         // (n_hash + 2) bin records of hstride ints (mcb_tables.h)
};
struct DevNuclide {
    const double* rows;          // n_rows x {E, sigma_s, sigma_c, sigma_f, nu, beta}; 48-byte rows, 16-byte aligned
    int32_t n_rows, has_delayed;
    double A;
    double fg_beta;              // sqrt(2.0659834e-11 * A): free-gas target speed parameter (Reaction.cpp:33), a per-nuclide constant
    double watt_a[3], watt_b[3], watt_g[3];
    // delayed neutrons (setup.cpp:377-412): decay constants, group fractions, tabulated emission spectra
    double lambda[6], fraction[6];
    const double* chid_E;
    const double* chid_cdf[6];
    int32_t chid_cdf_n[6];
};
struct DevProblem {
    int32_t ksearch, n_materials, n_nuclides, n_surfaces, n_cells, n_sources, n_estimators, entropy_on;
    int32_t shared_histories;    // several particles of one history can be in flight (secondaries / splitting)
    int32_t track_old;           // some estimator reads Particle::energy_old (TRMM tally set): Bank::Eold is maintained
    int32_t track_time;          // some estimator has a time filter: Bank::told (Particle::time_old) is maintained
    int32_t n_tdmc;              // census times of the time-dependent mode (general.cpp:187-195), 0 = off
    const double* tdmc_time;     // n_tdmc
    const double* tdmc_interval; // n_tdmc, as the reference computes them (setup.cpp:146)
    int32_t comb_teeth;          // particle comb (population_control.cpp:55-84): 0 = off
    int32_t comb_bank_max, pad;
    double wr, ws;
    uint64_t seed0, n_sample;
    const DevMaterial* materials;
    const DevNuclide* nuclides;
    const int32_t* mat_nuclide;
    const double* mat_density;
    const mcb_surface* surfaces;
    const mcb_cell* cells;
    const int32_t* cell_surface;
    const int32_t* cell_sense;
    const int32_t* cross_neighbor;        // [2 s + (side > 0)]: the cell on that side of surface s when search_cell's answer is provably that cell, else -1 (mcb_api.cu)
    const mcb_source* sources;
    const mcb_estimator* estimators;
    const mcb_score* scores;
    const mcb_filter* filters;
    const double* filter_grid;
    // estimators attached to surface s / cell c (TL) / cell c (C): CSR lists in deck order
    const int32_t* attach_begin[3];
    const int32_t* attach_list[3];
    int32_t entropy_n[3];
    int32_t entropy_bins;
    const double* entropy_grid;
};

// SoA particle bank: one slot per particle in flight
struct Bank {
    double *x, *y, *z, *u, *v, *w, *E, *speed, *wgt, *t;
    uint64_t* rng;
    int32_t *cell, *hist;                 // hist = history index local to this rank's shard
    double* Eold;                         // Particle::energy_old (Particle.cpp:42-56), only when DevProblem::track_old
    double* told;                         // Particle::time_old (Particle.cpp:66-76), only when DevProblem::track_time
    double *St, *Ss, *Sc, *Sf, *nSf;      // macroscopic xs of the cell's material at E (stage: xs_lookup)
    int32_t* uidx;                        // union-grid index of E in that material
    int32_t* surf;                        // surface hit by the last flight (stage: flight)
};

// fission site / source site (AoS: sites are written and read by random index).
// 64 bytes = two 32-byte sectors, fetched with four 128-bit loads: it is drawn from local HBM or from a peer's HBM over
// NVLink, where every sector is a request on the wire.  The isotropic emission direction is kept as the two numbers
// it is made of (DistributionIsotropicDirection::sample, Distribution.cpp:78-92: mu = 2 xi1 - 1 and the azimuth draw
// xi2) and rebuilt by site_direction() with the very expressions that sample it, so nothing is lost.  A bank handed
// in from the host carries explicit directions instead, in a side array (SourceBankView::dir_x).
struct alignas(16) Site {
    double x, y, z, E, t;
    double mu, xi;                        // direction = (mu, cos(2 pi xi) sqrt(1-mu^2), sin(2 pi xi) sqrt(1-mu^2))
    int32_t cell, pad;
};
static_assert(sizeof(Site) == 64, "Site is two 32-byte sectors");
__device__ __forceinline__ Site load_site(const Site* p)
{
    const double2* q = reinterpret_cast<const double2*>(p);
    const double2 a = q[0], b = q[1], c = q[2];
    const int4 d = *reinterpret_cast<const int4*>(q + 3);
    Site s;
    s.x = a.x; s.y = a.y; s.z = b.x; s.E = b.y; s.t = c.x; s.mu = c.y;
    s.xi = __hiloint2double(d.y, d.x); s.cell = d.z; s.pad = 0;
    return s;
}
__device__ __forceinline__ void store_site(Site* p, const Site& s)
{
    double2* q = reinterpret_cast<double2*>(p);
    q[0] = make_double2(s.x, s.y); q[1] = make_double2(s.z, s.E); q[2] = make_double2(s.t, s.mu);
    *reinterpret_cast<int4*>(q + 3) = make_int4(__double2loint(s.xi), __double2hiint(s.xi), s.cell, 0);
}
// direction of an isotropic emission from its two draws (Distribution.cpp:78-92: x is the polar axis)
__device__ __forceinline__ void direction_from_draws(double mu, double xi, double& dx, double& dy, double& dz)
{
    const double c = sqrt(1.0 - mu * mu);
    double sa, ca;
#ifdef MCB_FAST_TRIG
    sincospi(2.0 * xi, &sa, &ca);
#else
    const double azi = MCB_PI_2 * xi;
    sincos(azi, &sa, &ca);
#endif
    dy = ca * c;
    dz = sa * c;
    dx = mu;
}

// sin / cos of the azimuth 2 pi xi, the way direction_from_draws forms them
__device__ __forceinline__ void circle_from_draw(double xi, double& sa, double& ca)
{
#ifdef MCB_FAST_TRIG
    sincospi(2.0 * xi, &sa, &ca);
#else
    const double azi = MCB_PI_2 * xi;
    sincos(azi, &sa, &ca);
#endif
}

// The source bank a generation samples from.  Single GPU / a bank handed in from the host: one flat array.
// Multi-GPU: the global bank is the rank-order concatenation of every rank's canonical slice; the slices stay where
// they were written and are read in place over NVLink through peer pointers (CUDA IPC), so the bank is never
// gathered: seg[r] = rank r's slice, prefix[r] = global index of its first site.
#define MCB_MAX_WORLD 16
struct SourceBankView {
    const Site* flat;
    const double* dir_x;                  // flat banks handed in from the host: n x 3 explicit directions (else nullptr)
    const Site* seg[MCB_MAX_WORLD];
    const double* seg_dir[MCB_MAX_WORLD]; // slices handed in from the host, one per rank: explicit directions (else nullptr)
    unsigned long long prefix[MCB_MAX_WORLD + 1];
    int32_t n_seg, pad;
    unsigned long long n;                 // total sites; 0 with flat == nullptr and n_seg == 0 -> sample the deck's sources
};
__device__ __forceinline__ int source_bank_locate(const SourceBankView& V, unsigned long long j, unsigned long long& local)
{
    if (V.flat) { local = j; return -1; }
    int r = 0;
    while (r + 1 < V.n_seg && j >= V.prefix[r + 1]) r++;
    local = j - V.prefix[r];
    return r;
}
__device__ __forceinline__ Site source_bank_site(const SourceBankView& V, unsigned long long j)
{
    unsigned long long local;
    const int r = source_bank_locate(V, j, local);
    return load_site((r < 0 ? V.flat : V.seg[r]) + local);
}
__device__ __forceinline__ void source_bank_direction(const SourceBankView& V, unsigned long long j, const Site& s, double& u, double& v,
                                                      double& w)
{
    if (V.dir_x) { u = V.dir_x[3 * j]; v = V.dir_x[3 * j + 1]; w = V.dir_x[3 * j + 2]; return; }
    if (!V.flat) {
        unsigned long long local;
        const int r = source_bank_locate(V, j, local);
        const double* d = V.seg_dir[r];
        if (d) { u = d[3 * local]; v = d[3 * local + 1]; w = d[3 * local + 2]; return; }
    }
    direction_from_draws(s.mu, s.xi, u, v, w);
}

// a fission site as the collision leaves it: where, from what, and the stream it will be sampled from.  The
// outgoing energy and direction are drawn later, in one dense pass over the whole generation's requests.
struct SiteReq {
    double x, y, z, t, E_in;
    uint64_t seed;
    int32_t cell, seq, hist, nuclide;
};

struct Counters {
    unsigned long long n_tracks, n_collisions, n_lookups, n_crossings, n_histories;
    unsigned long long site_cursor;       // unordered fission sites banked this cycle
    unsigned long long slot_cursor;       // bank slots in use (primaries + secondaries of the batch)
    unsigned long long walk_head;         // history walk: slots of the current pass handed out so far
    unsigned long long src_ready;         // bank positions k_source has filled so far (sweep running beside the walk)
    unsigned long long n_active[3];       // lengths of the particle queue: iteration i reads [i%3], fills [(i+1)%3], clears [(i+2)%3]
    unsigned long long q_collide, q_cross;// lengths of the two halves of the event queue
    int lost, overflow_sites, overflow_slots, overflow_fixed;
    int overflow_tally, overflow_stack;   // walk kernel: a history's tally table / secondary stack is full
    int hang, pad_i;                      // walk kernel: a bounded wait ran out (never expected; reported as an error)
    long long live;                       // walk kernel: work units alive on this GPU (histories + donated secondaries)
    unsigned long long n_donated, n_shared_hist, n_donate_refused, n_idle_waits;  // work-sharing statistics
    double lost_pos[3];
    // exact (fixed-point, two-limb) sums over histories: k_C, k_TL, k_C^2, k_TL^2, H
    unsigned long long fx_lo[5], fx_hi[5];
};

// per-history accumulators, indexed by the shard-local history index
struct HistoryAcc {
    double *kC, *kTL;                     // EstimatorK::k_C / k_TL of the history (Estimator.cpp:503-512)
    int32_t* nsite;                       // fission sites banked by the history
};

// per-history tally accumulators.
// Event-queue kernels: dense rows acc[tally][history of the batch] (acc != nullptr), reduced per batch by k_tally_*.
// Walk kernel: one table per history context (a history is followed by one lane at a time): direct ({position ->
// value}, every tally its own position) when the problem has at most 8192 tallies, else open-addressed {tally index
// + 1 -> value}; flushed when the history ends as sum += v, squared += v * v (Estimator.cpp:339-346) with
// reductions into `sum` / `squared` (through block-private shared-memory bins when the problem has few tallies).
struct TallyAcc {
    double* acc;
    int64_t stride;                       // histories per batch (row length)
    int32_t first_hist;                   // shard-local index of the batch's first history
    int32_t on;                           // tallies are being scored (active cycle, handler.cpp:15)
    uint32_t* tab_key;                    // n_contexts x (tab_mask + 1)
    double* tab_val;
    uint16_t* tab_list;                   // positions in use, in order of first touch
    uint32_t tab_mask;                    // table size - 1 (a power of two >= the number of tallies when that fits)
    int32_t n_tallies;
    double *sum, *squared;                // Tally::sum / squared of the cycle on this rank
    // histories whose secondaries were handed to other lanes (long fission chains): their units merge into one dense
    // row each; the unit that finishes last (pending reaches 0) turns the row into sum / squared
    double* dense;                        // dense_rows x n_tallies
    int32_t* dense_pending;
    int32_t* dense_cursor;                // rows handed out in this launch
    int32_t dense_rows;
    int32_t direct;                       // tables hold every tally at its own position (no keys), see tally_slot
};


// a same-history secondary waiting on its history's LIFO stack (the reference's Pbank, handler.cpp:20-29)
struct alignas(16) StackRec {
    double x, y, z, u, v, w, E, speed, wgt, t, Eold;
    uint64_t rng;
    int32_t cell, pad0;
    double pad1;
};
static_assert(sizeof(StackRec) == 112, "StackRec is seven 16-byte words");

// secondaries handed over between lanes (work sharing for long fixed-source fission chains): a bounded multi-producer
// multi-consumer ring in global memory, one sequence number per cell (free for ticket t when seq = t, filled when t + 1)
struct DonationQueue {
    unsigned long long head, tail;
    int avail, count;
    unsigned long long* seq;
    StackRec* recs;
    uint32_t cap_mask, pad;
};

#define MCB_FX_SCALE 17592186044416.0     /* 2^44: fixed-point scale of the exact history sums */

// ---------------------------------------------------------------------------------------------
// cross sections
// ---------------------------------------------------------------------------------------------
struct MicroXS { double s, c, f, t, nu, beta; };

__device__ __forceinline__ double2 ld_row2(const double* p)
{
    return __ldg(reinterpret_cast<const double2*>(p));
}

// XSTable::xs for all tables of one nuclide at E (XSec.cpp:9-40, Algorithm.cpp:103-105); idx = #{n_E < E} - 1.
// sigma_t / sigma_a are interpolated from their per-grid-point values (setup.cpp:371-372), not summed afterwards.
static __device__ __noinline__ int row_bisect_cold(const double* rows, int n, double E) { return mcb_row_bisect(rows, n, E); }
__device__ __forceinline__ void micro_xs(const DevNuclide& N, int idx, double E, MicroXS& m)
{
    if (idx == MCB_MAP_BISECT) idx = row_bisect_cold(N.rows, N.n_rows, E);
    if (idx < 0 || idx >= N.n_rows - 1) {
        const double* r = N.rows + (size_t)(idx < 0 ? 0 : N.n_rows - 1) * MCB_XS_ROW;
        const double2 a = ld_row2(r), b = ld_row2(r + 2), c = ld_row2(r + 4);
        m.s = a.y; m.c = b.x; m.f = b.y; m.nu = c.x; m.beta = c.y;
        m.t = a.y + b.x + b.y;
        return;
    }
    const double* r = N.rows + (size_t)idx * MCB_XS_ROW;
    const double2 a1 = ld_row2(r), b1 = ld_row2(r + 2), c1 = ld_row2(r + 4);
    const double2 a2 = ld_row2(r + 6), b2 = ld_row2(r + 8), c2 = ld_row2(r + 10);
    const double E1 = a1.x, E2 = a2.x;
    // interpolate (Algorithm.cpp:103-105): E2 - E1 = -(E1 - E2) exactly and x / (-d) = -(x / d) exactly, so one
    // reciprocal serves both weights (mcb_div_shared is bit-identical to the IEEE division)
    const double d = E1 - E2, rd = mcb_rcp_shared(d);
    const double f1 = mcb_div_shared(E - E2, d, rd);
    const double f2 = -mcb_div_shared(E - E1, d, rd);
    m.s = f1 * a1.y + f2 * a2.y;
    m.c = f1 * b1.x + f2 * b2.x;
    m.f = f1 * b1.y + f2 * b2.y;
    m.nu = f1 * c1.x + f2 * c2.x;
    m.beta = f1 * c1.y + f2 * c2.y;  // Nuclide::beta (Nuclide.cpp:74-77): the same interpolate() expression
    m.t = f1 * (a1.y + b1.x + b1.y) + f2 * (a2.y + b2.x + b2.y);
}
// one derived column: 0 sigma_a (= sigma_c + sigma_f per grid point), 1 beta
__device__ __forceinline__ double micro_col(const DevNuclide& N, int idx, double E, int col)
{
    if (idx == MCB_MAP_BISECT) idx = row_bisect_cold(N.rows, N.n_rows, E);
    if (idx < 0 || idx >= N.n_rows - 1) {
        const double* r = N.rows + (size_t)(idx < 0 ? 0 : N.n_rows - 1) * MCB_XS_ROW;
        return col == 0 ? r[2] + r[3] : r[5];
    }
    const double* r1 = N.rows + (size_t)idx * MCB_XS_ROW;
    const double* r2 = r1 + MCB_XS_ROW;
    const double y1 = col == 0 ? r1[2] + r1[3] : r1[5];
    const double y2 = col == 0 ? r2[2] + r2[3] : r2[5];
    return mcb_interpolate(E, r1[0], r2[0], y1, y2);
}

// where an energy sits in the material's union grid: u = #{U < E} - 1 and the per-nuclide row indices, read from
// the bin record when the energy lies below every grid point of its bin (mcb_union_lookup), else from map[u]
struct UnionPos {
    int u;
    const int32_t* rec;   // bin record (indices at rec[2 + n])
    const int32_t* row;   // map row of u when the record does not apply, else nullptr
};
__device__ __forceinline__ UnionPos union_pos(const DevMaterial& M, double E)
{
    UnionPos q;
#ifndef MCB_HREC_LOOKUP  // measured: the record form changes nothing in the walk kernel (5.74 ms either way) and costs the
                         // lookup microbench 10 % (its table is four times the hash), so the plain hash stays the default
    q.u = mcb_union_count_less(M.U, M.hash, M.key_min, M.n_hash, M.shift, M.nU, E) - 1;
    q.rec = M.hrec + (size_t)(M.n_hash + 1) * M.hstride;  // the "below the grid" record: indices -1
    q.row = q.u < 0 ? nullptr : M.map + (size_t)q.u * M.n_nuc;
#else
    bool from_rec;
    const int lo = mcb_union_lookup(M.U, M.hrec, M.hstride, M.key_min, M.n_hash, M.shift, E, &q.rec, &from_rec);
    q.u = lo - 1;
    q.row = from_rec ? nullptr : M.map + (size_t)q.u * M.n_nuc;
#endif
    return q;
}
__device__ __forceinline__ int nuclide_index(const UnionPos& q, int n)
{
    const int ir = __ldg(q.rec + 2 + n);  // issued whatever the outcome of the search in the bin
    return q.row ? __ldg(q.row + n) : ir;
}
__device__ __forceinline__ int union_index(const DevMaterial& M, double E)
{
    return mcb_union_count_less(M.U, M.hash, M.key_min, M.n_hash, M.shift, M.nU, E) - 1;
}
__device__ __forceinline__ int nuclide_index(const DevMaterial& M, int u, int n)
{
    return u < 0 ? -1 : __ldg(&M.map[(size_t)u * M.n_nuc + n]);
}

struct MacroXS { double t, s, c, f, nf; };

// per-nuclide by-products of one macroscopic lookup.  The running sums ARE the partial sums that
// Material::nuclide_scatter / nuclide_nufission (Material.cpp:106-125) rebuild, so a kernel that keeps them can
// pick the reaction nuclide without evaluating the tables again.  Detail types: set(n, ..) while the lookup runs,
// cum_s(n) / cum_nf(n) / beta(n) afterwards; XSDetail keeps them in local memory, the walk kernel's SlotDetail in the
// particle's shared-memory slot, NoDetail drops them.
struct XSDetail {
    static constexpr bool present = true;
    double cs[MCB_MAX_MAT_NUCLIDES];      // Sigma_s after nuclides 0..n
    double cnf[MCB_MAX_MAT_NUCLIDES];     // nuSigma_f after nuclides 0..n
    double b[MCB_MAX_MAT_NUCLIDES];       // beta_n(E)
    __device__ __forceinline__ void set(int n, double s, double nf, double be) { cs[n] = s; cnf[n] = nf; b[n] = be; }
    __device__ __forceinline__ double cum_s(int n) const { return cs[n]; }
    __device__ __forceinline__ double cum_nf(int n) const { return cnf[n]; }
    __device__ __forceinline__ double beta(int n) const { return b[n]; }
};
struct NoDetail {
    static constexpr bool present = false;
    __device__ __forceinline__ void set(int, double, double, double) {}
    __device__ __forceinline__ double cum_s(int) const { return 0.0; }
    __device__ __forceinline__ double cum_nf(int) const { return 0.0; }
    __device__ __forceinline__ double beta(int) const { return 0.0; }
};

// Material::SigmaT/S/C/F, nuSigmaF (Material.cpp:18-65): sums over nuclides in deck order, starting from 0.0
template <class DET>
__device__ __forceinline__ void macro_xs_impl(const DevProblem& P, const DevMaterial& M, const UnionPos& q, double E, MacroXS& X, DET& D)
{
    X.t = 0.0; X.s = 0.0; X.c = 0.0; X.f = 0.0; X.nf = 0.0;
    for (int n = 0; n < M.n_nuc; n++) {
        const int gn = __ldg(&P.mat_nuclide[M.nuc_begin + n]);
        const double dens = __ldg(&P.mat_density[M.nuc_begin + n]);
        MicroXS m;
        micro_xs(P.nuclides[gn], nuclide_index(q, n), E, m);
        X.t += m.t * dens;
        X.s += m.s * dens;
        X.c += m.c * dens;
        X.f += m.f * dens;
        X.nf += (m.f * m.nu) * dens;   // Nuclide::nusigmaF = sigmaF*nu (Nuclide.cpp:57-61)
        if (DET::present) D.set(n, X.s, X.nf, m.beta);
    }
}
__device__ __forceinline__ void macro_xs(const DevProblem& P, const DevMaterial& M, const UnionPos& q, double E, MacroXS& X)
{
    NoDetail nd;
    macro_xs_impl(P, M, q, E, X, nd);
}
// nuclide pick from kept partial sums: first n with cum[n] > total*xi (Material.cpp:106-125); KIND 0 scatter, 1 nu-fission
template <int KIND, class DET>
__device__ __forceinline__ int select_from_detail(const DevProblem& P, const DevMaterial& M, const DET& D, double total,
                                                  double xi, int* local_n)
{
    const double thr = total * xi;
    int sel = -1;
    for (int n = M.n_nuc - 1; n >= 0; n--)   // single exit; the last hit is the first nuclide with cum > thr
        if ((KIND == 0 ? D.cum_s(n) : D.cum_nf(n)) > thr) sel = n;
    if (sel < 0) return -1;
    *local_n = sel;
    return __ldg(&P.mat_nuclide[M.nuc_begin + sel]);
}
// Material::SigmaA (Material.cpp:34-41)
__device__ __forceinline__ double macro_sigma_a(const DevProblem& P, const DevMaterial& M, int u, double E)
{
    double sum = 0.0;
    for (int n = 0; n < M.n_nuc; n++) {
        const int gn = __ldg(&P.mat_nuclide[M.nuc_begin + n]);
        sum += micro_col(P.nuclides[gn], nuclide_index(M, u, n), E, 0) * __ldg(&P.mat_density[M.nuc_begin + n]);
    }
    return sum;
}
// Material::nuclide_scatter (kind 0) / nuclide_nufission (kind 1) (Material.cpp:106-125): global nuclide index or -1.
// `total` is the macroscopic xs the partial sums are compared against (SigmaS or nuSigmaF at the same E).
// The loop has a single exit (no early return): lanes that pick different nuclides must leave it together, or the
// warp runs everything that follows once per picked nuclide.
__device__ __forceinline__ int select_nuclide(const DevProblem& P, const DevMaterial& M, int u, double E, int kind,
                                              double total, double xi, int* local_n)
{
    const double thr = total * xi;
    double s = 0.0;
    int sel = -1;
    for (int n = 0; n < M.n_nuc; n++) {
        const int gn = __ldg(&P.mat_nuclide[M.nuc_begin + n]);
        MicroXS m;
        micro_xs(P.nuclides[gn], nuclide_index(M, u, n), E, m);
        s += (kind == 0 ? m.s : m.f * m.nu) * __ldg(&P.mat_density[M.nuc_begin + n]);
        if (sel < 0 && s > thr) sel = n;
    }
    if (sel < 0) return -1;
    if (local_n) *local_n = sel;
    return __ldg(&P.mat_nuclide[M.nuc_begin + sel]);
}

// Reaction channels of the TRMM tally set: 0 scatter, 1 nu-fission, 2 prompt nu-fission, 3+g delayed nu-fission of
// precursor group g (Nuclide.cpp:57-73): sigma_s | sigma_f*nu | (1-beta)*sigma_f*nu | beta*fraction_g*sigma_f*nu
__device__ __forceinline__ double micro_channel(const DevNuclide& N, const MicroXS& m, int kind)
{
    if (kind == 0) return m.s;
    if (kind == 1) return m.f * m.nu;
    if (kind == 2) return (1.0 - m.beta) * m.f * m.nu;
    return m.beta * N.fraction[kind - 3] * m.f * m.nu;
}
// The per-nuclide microscopic data of one material at one energy, kept while one event is scored: the TRMM tally set
// asks for a dozen different channel sums at the same two energies (the particle's energy and its energy_old).
struct ChannelCache {
    double E[2];
    int32_t mat[2], used;
    uint32_t valid[2];                       // channel sums already formed at E[i]: bit kind, bit 9 + g for the decay sums
    MicroXS m[2][MCB_MAX_MAT_NUCLIDES];
    double sums[2][15];
    // bins of the last two (energy filter grid, energy) pairs searched, most recent first: the estimators of one event
    // look up the same energy in the same grid again and again
    double fb_val[2];
    int32_t fb_grid[2], fb_idx[2];
};
__device__ __forceinline__ void channel_cache_reset(ChannelCache& C) { C.mat[0] = C.mat[1] = -1; C.used = 0; C.fb_grid[0] = C.fb_grid[1] = -1; }
__device__ __noinline__ static int channel_cache_get(const DevProblem& P, int material, double E, ChannelCache& C)
{
    for (int i = 0; i < 2; i++) if (C.mat[i] == material && C.E[i] == E) return i;
    const int i = C.used & 1;
    C.used++;
    const DevMaterial& M = P.materials[material];
    const UnionPos up = union_pos(M, E);
    for (int n = 0; n < M.n_nuc; n++) {
        const int gn = __ldg(&P.mat_nuclide[M.nuc_begin + n]);
        micro_xs(P.nuclides[gn], nuclide_index(up, n), E, C.m[i][n]);
    }
    C.E[i] = E; C.mat[i] = material; C.valid[i] = 0u;
    return i;
}
// mcb_binary_search(x, grid, n) of a filter grid through the two-entry cache
__device__ __forceinline__ int filter_bin(const DevProblem& P, int grid_begin, int n, double x, ChannelCache& C)
{
    if (C.fb_grid[0] == grid_begin && C.fb_val[0] == x) return C.fb_idx[0];
    int idx;
    if (C.fb_grid[1] == grid_begin && C.fb_val[1] == x) idx = C.fb_idx[1];
    else idx = mcb_binary_search(x, P.filter_grid + grid_begin, n);
    C.fb_grid[1] = C.fb_grid[0]; C.fb_val[1] = C.fb_val[0]; C.fb_idx[1] = C.fb_idx[0];
    C.fb_grid[0] = grid_begin; C.fb_val[0] = x; C.fb_idx[0] = idx;
    return idx;
}
// Material::SigmaS / nuSigmaF / nuSigmaF_prompt / nuSigmaF_delayed (Material.cpp:26-82) at any energy; with `picked`
// also Material::nuclide_scatter / _nufission / _nufission_prompt / _nufission_delayed (Material.cpp:106-146): *picked =
// global nuclide index or -1.  decay = true: nuSigmaF_delayed_decay (Material.cpp:83-91), each term divided by lambda_g.
__device__ __noinline__ static double macro_channel(const DevProblem& P, int material, double E, int kind, bool decay, double xi,
                                                    int* picked, ChannelCache& C)
{
    const DevMaterial& M = P.materials[material];
    const int ci = channel_cache_get(P, material, E, C);
    const MicroXS* m = C.m[ci];
    const int slot = decay ? kind + 6 : kind;
    double sum;
    if (C.valid[ci] >> slot & 1u) sum = C.sums[ci][slot];
    else {
        sum = 0.0;
        for (int n = 0; n < M.n_nuc; n++) {
            const DevNuclide& N = P.nuclides[__ldg(&P.mat_nuclide[M.nuc_begin + n])];
            const double v = micro_channel(N, m[n], kind);
            sum += (decay ? v / N.lambda[kind - 3] : v) * __ldg(&P.mat_density[M.nuc_begin + n]);
        }
        C.sums[ci][slot] = sum;
        C.valid[ci] |= 1u << slot;
    }
    if (picked) {
        const double thr = sum * xi;
        double s = 0.0;
        int sel = -1;
        for (int n = 0; n < M.n_nuc; n++) {
            const int gn = __ldg(&P.mat_nuclide[M.nuc_begin + n]);
            s += micro_channel(P.nuclides[gn], m[n], kind) * __ldg(&P.mat_density[M.nuc_begin + n]);
            if (sel < 0 && s > thr) sel = gn;
        }
        *picked = sel;
    }
    return sum;
}

// ---------------------------------------------------------------------------------------------
// distributions / reactions
// ---------------------------------------------------------------------------------------------
// ReactionFission::ChiD -> DistributionDelayedNeutron::sample (Distribution.cpp:97-102): tabulated CDF, lin-lin
__device__ __forceinline__ double chid_sample(const DevNuclide& N, int g, uint64_t& rng)
{
    const double xi = mcb_urand(rng);
    const double* cdf = N.chid_cdf[g];
    const int idx = mcb_binary_search(xi, cdf, N.chid_cdf_n[g]);
    return mcb_interpolate(xi, cdf[idx], cdf[idx + 1], N.chid_E[idx], N.chid_E[idx + 1]);
}
// DistributionWatt::sample (Distribution.cpp:34-73); returns eV
__device__ __forceinline__ double watt_sample(const double* va, const double* vb, const double* vg, double E,
                                              uint64_t& rng)
{
    double a, b, g;
    if (E <= 1.0) { a = va[0]; b = vb[0]; g = vg[0]; }
    else if (E <= 1.0e6) {
        a = mcb_interpolate(E, 1.0, 1.0e6, va[0], va[1]);
        b = mcb_interpolate(E, 1.0, 1.0e6, vb[0], vb[1]);
        g = mcb_interpolate(E, 1.0, 1.0e6, vg[0], vg[1]);
    } else {
        a = mcb_interpolate(E, 1.0e6, 14.0e6, va[1], va[2]);
        b = mcb_interpolate(E, 1.0e6, 14.0e6, vb[1], vb[2]);
        g = mcb_interpolate(E, 1.0e6, 14.0e6, vg[1], vg[2]);
    }
    double Eout, C;
    do {
        const double lx = mcb_log(mcb_urand(rng));
        Eout = -a * g * lx;
        const double l2 = mcb_log(mcb_urand(rng));
        C = (1.0 - g) * (1.0 - lx) - l2;
    } while (C * C > b * Eout);
    return Eout * 1.0e6;
}
// DistributionIsotropicDirection::sample (Distribution.cpp:78-92): x is the polar axis
__device__ __forceinline__ void isotropic_direction(uint64_t& rng, double& dx, double& dy, double& dz)
{
    const double mu = 2.0 * mcb_urand(rng) - 1.0;
    const double xi = mcb_urand(rng);
    direction_from_draws(mu, xi, dx, dy, dz);
}
// Delta / Uniform (Distribution.cpp:30-33) / Watt at incident energy 0
__device__ __forceinline__ double dist1_sample(const mcb_dist1& d, uint64_t& rng)
{
    switch (d.kind) {
    case MCB_DIST_DELTA: return d.a;
    case MCB_DIST_UNIFORM: return d.a + mcb_urand(rng) * (d.b - d.a);
    default: return watt_sample(d.watt_a, d.watt_b, d.watt_g, 0.0, rng);
    }
}

// ReactionScatter::sample (Reaction.cpp:27-118): elastic scatter off a free-gas target at 293.6 K, isotropic in
// the centre of mass.  In/out: direction, energy and speed of the neutron.  Each three-component normalisation
// divides by one number: the reciprocal is refined once (mcb_div_shared, bit-identical to three IEEE divisions).
__device__ __forceinline__ void scatter_sample(const DevNuclide& N, double& dx, double& dy, double& dz, double& E, double& speed,
                                               uint64_t& rng)
{
    const double A = N.A;
    const double mu0 = 2.0 * mcb_urand(rng) - 1.0;  // DistributionIsotropicScatter (Distribution.cpp:74-77)
    const double beta = N.fg_beta;                  // sqrt(2.0659834e-11 * A)
    const double y = beta * speed;
    double V_tilda, mu_tilda, accept;
    do {
        double x;
        if (mcb_urand(rng) < 2.0 / (2.0 + MCB_PI_SQRT * y)) {
            const double r1 = mcb_urand(rng), r2 = mcb_urand(rng);
            x = sqrt(-mcb_log(r1 * r2));
        } else {
#ifdef MCB_FAST_TRIG
            const double cos_val = cospi(0.5 * mcb_urand(rng));
#else
            const double cos_val = cos(MCB_PI_HALF * mcb_urand(rng));
#endif
            const double l1 = mcb_log(mcb_urand(rng));
            const double l2 = mcb_log(mcb_urand(rng));
            x = sqrt(-l1 - l2 * cos_val * cos_val);
        }
        V_tilda = x / beta;
        mu_tilda = 2.0 * mcb_urand(rng) - 1.0;
        accept = mcb_urand(rng);
    } while (accept > sqrt(speed * speed + V_tilda * V_tilda - 2.0 * speed * V_tilda * mu_tilda) / (speed + V_tilda));
    double nx, ny, nz;
    mcb_scatter_direction(dx, dy, dz, mu_tilda, mcb_urand(rng), nx, ny, nz);
    const double Vx = nx * V_tilda, Vy = ny * V_tilda, Vz = nz * V_tilda;
    double vx = speed * dx, vy = speed * dy, vz = speed * dz;
    const double A1 = 1.0 + A, rA1 = mcb_rcp_shared(A1);
    const double ux = mcb_div_shared(vx + A * Vx, A1, rA1);
    const double uy = mcb_div_shared(vy + A * Vy, A1, rA1);
    const double uz = mcb_div_shared(vz + A * Vz, A1, rA1);
    double cx = vx - ux, cy = vy - uy, cz = vz - uz;
    const double speed_c = sqrt(cx * cx + cy * cy + cz * cz);
    const double rsc = mcb_rcp_shared(speed_c);
    const double dcx = mcb_div_shared(cx, speed_c, rsc), dcy = mcb_div_shared(cy, speed_c, rsc), dcz = mcb_div_shared(cz, speed_c, rsc);
    double ex, ey, ez;
    mcb_scatter_direction(dcx, dcy, dcz, mu0, mcb_urand(rng), ex, ey, ez);
    cx = speed_c * ex; cy = speed_c * ey; cz = speed_c * ez;
    vx = cx + ux; vy = cy + uy; vz = cz + uz;
    speed = sqrt(vx * vx + vy * vy + vz * vz);  // Particle::set_speed (Particle.cpp:49-56)
    E = mcb_energy_of_speed(speed);
    const double rsp = mcb_rcp_shared(speed);
    dx = mcb_div_shared(vx, speed, rsp); dy = mcb_div_shared(vy, speed, rsp); dz = mcb_div_shared(vz, speed, rsp);
}

// surface_intersect (general.cpp:54-67): nearest surface of the cell along the flight direction
__device__ __forceinline__ int surface_intersect(const DevProblem& P, int cell, double x, double y, double z,
                                                 double u, double v, double w, double& dist_out)
{
    const mcb_cell C = P.cells[cell];
    double dist = MCB_MAX_FLOAT;
    int S = -1;
    for (int i = C.surf_begin; i < C.surf_end; i++) {
        const int s = __ldg(&P.cell_surface[i]);
        const double d = mcb_surf_distance(P.surfaces[s], x, y, z, u, v, w);
        if (d < dist) { dist = d; S = s; }
    }
    dist_out = dist;
    return S;
}

#endif
