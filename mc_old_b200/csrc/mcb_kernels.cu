// mcb_kernels.cu — sm_100a kernels of the particle-history transport loop.
//
// One generation (reference: one pass of the cycle body of Simulator::start(), handler.cpp:14-44) is
//   k_source -> k_walk (one launch per pass over the bank) -> close-out kernels.
// k_source fills an SoA bank (Bank) with the generation's source particles.  k_walk follows every particle through
// its whole chain of events { xs_lookup -> flight -> collide | cross } in registers; warps are autonomous, draw bank
// slots in chunks from a device-side head counter and refill a lane as soon as its particle has ended, so the lanes
// stay busy until the bank runs dry.  Fission sites leave the kernel as requests (one warp-aggregated cursor
// reservation per iteration) and are sampled and put into canonical order by k_bank_sample_order.
//
// The event-queue formulation of the same loop (k_xs_stage / k_flight / k_collide / k_cross over index queues
// rebuilt by block-level stream compaction, k_finish for the tail) is kept as a cross-check (MCB_MODE=split): both
// must produce bit-identical generations (tests/test_gpu_transport.py).
//
// Nothing here is GEMM-shaped: the loop is FP64 scalar work, L2-resident table gathers and HBM streams of the
// bank, so tensor cores are unused on purpose.
#include "mcb_kernels.h"

#include <algorithm>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace {

constexpr unsigned FULL = 0xffffffffu;
#ifndef MCB_BLOCK
#define MCB_BLOCK 128  // measured (tools/sweep.sh): 128 x 6 blocks per SM (85 registers) beats 128 x 5, 128 x 7, 256 x 2, 192 x 3 and 96 x 7 for the walk kernel
#endif
constexpr int BLOCK = MCB_BLOCK;
constexpr int WARPS = BLOCK / 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// ---------------------------------------------------------------------------------------------
// block-level stream compaction: every thread asks for n[c] consecutive positions behind cursor c; one
// atomicAdd per block and cursor (same-address atomics serialise in L2, so per-warp cursors were the bottleneck).
// Must be reached by all threads of the block.  `S` is per-tile scratch; callers alternate two of them.
// ---------------------------------------------------------------------------------------------
template <int NC>
struct BlockScratch {
    unsigned warp_tot[NC][WARPS];
    unsigned long long base[NC];
};
template <int NC>
__device__ __forceinline__ void block_reserve(BlockScratch<NC>& S, const unsigned (&n)[NC], unsigned long long* const (&cursor)[NC],
                                              unsigned long long (&pos)[NC])
{
    unsigned incl[NC];
#pragma unroll
    for (int c = 0; c < NC; c++) {
        unsigned v = n[c];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(FULL, v, d);
            if (lane_id() >= d) v += t;
        }
        incl[c] = v;
        if (lane_id() == 31) S.warp_tot[c][warp_id()] = v;
    }
    __syncthreads();
    if (threadIdx.x < NC) {
        const int c = threadIdx.x;
        unsigned run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; w++) { const unsigned t = S.warp_tot[c][w]; S.warp_tot[c][w] = run; run += t; }
        S.base[c] = run ? atomicAdd(cursor[c], (unsigned long long)run) : 0ull;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < NC; c++) pos[c] = S.base[c] + S.warp_tot[c][warp_id()] + (incl[c] - n[c]);
}
// block-wide event counter: one atomic per block
__device__ __forceinline__ void block_count(unsigned long long* dst, bool pred)
{
    const int c = __syncthreads_count(pred);
    if (threadIdx.x == 0 && c) atomicAdd(dst, (unsigned long long)c);
}

// per-history accumulators are bumped with reductions at L2 (RED, no return value): a read-modify-write in the
// thread would stall the warp for a DRAM round trip on every event
__device__ __forceinline__ void hist_add(double* p, double v) { atomicAdd(p, v); }

// ---------------------------------------------------------------------------------------------
// tally scoring (Estimator::score, Estimator.cpp:298-336, for filters that yield one bin: surface, cell, energy)
// ---------------------------------------------------------------------------------------------
struct ScoreState {   // the particle as Score / Filter / the simulating estimators see it
    double w, E, E_old, speed, t, t_old;
    double u, v, wd;  // direction
    int cell, surface_old, material, uidx;
    MacroXS X;        // macroscopic xs of `material` at E
};

__device__ __forceinline__ double kernel_value(int kernel, const ScoreState& s, double l)  // Estimator.cpp:17-41
{
    switch (kernel) {
    case MCB_KERNEL_NEUTRON: return s.w;
    case MCB_KERNEL_TRACK: return s.w * l;
    case MCB_KERNEL_COLLISION: return s.w / s.X.t;
    case MCB_KERNEL_VELOCITY: return s.w * s.speed;
    default: return s.w * l * s.speed;
    }
}
__device__ __forceinline__ double score_value(const DevProblem& P, const mcb_score& S, const ScoreState& s, double l, ChannelCache& CC)
{  // Estimator.cpp:48-124
    const double kv = kernel_value(S.kernel, s, l);
    if (S.score == MCB_SCORE_FLUX) return kv;
    if (S.score == MCB_SCORE_INVERSE_VELOCITY) return kv / s.speed;
    if (s.material < 0) return 0.0;
    const DevMaterial& M = P.materials[s.material];
    switch (S.score) {
    case MCB_SCORE_ABSORPTION: return macro_sigma_a(P, M, s.uidx, s.E) * kv;
    case MCB_SCORE_SCATTER: return s.X.s * kv;
    case MCB_SCORE_CAPTURE: return s.X.c * kv;
    case MCB_SCORE_FISSION: return s.X.f * kv;
    case MCB_SCORE_NU_FISSION: return s.X.nf * kv;
    case MCB_SCORE_TOTAL: return s.X.t * kv;
    // the "Old" scores of the TRMM tally set evaluate at Particle::energy_old (Estimator.cpp:96-118)
    case MCB_SCORE_SCATTER_OLD: return macro_channel(P, s.material, s.E_old, 0, false, 0.0, nullptr, CC) * kv;
    case MCB_SCORE_NU_FISSION_OLD: return macro_channel(P, s.material, s.E_old, 1, false, 0.0, nullptr, CC) * kv;
    case MCB_SCORE_NU_FISSION_PROMPT_OLD: return macro_channel(P, s.material, s.E_old, 2, false, 0.0, nullptr, CC) * kv;
    case MCB_SCORE_NU_FISSION_DELAYED_OLD: return macro_channel(P, s.material, s.E_old, 3 + S.group, false, 0.0, nullptr, CC) * kv;
    case MCB_SCORE_NU_FISSION_DELAYED_DECAY_OLD: return macro_channel(P, s.material, s.E_old, 3 + S.group, true, 0.0, nullptr, CC) * kv;
    default: return 0.0;
    }
}
// Estimator::score (Estimator.cpp:298-336).  Surface / cell / energy / energy_old filters yield one bin; a time
// filter (Estimator.cpp:199-246) splits the track [t_old, t] over the bins it spans, piece by piece; the loop scores
// the shortest remaining piece of all filters, subtracts it everywhere and advances the exhausted ones, like the
// reference's.  The time pieces are generated on demand.
struct FilterCursor {
    int idx;         // current bin
    double l;        // remaining length of the current piece
    int k, n;        // piece number, pieces in all
    int loc1, loc2;  // time filter: bins of t_old and t
    bool first;      // time filter: the piece of t_old's bin exists (loc1 >= 0)
};
__device__ __forceinline__ void time_piece(const mcb_filter& F, const double* g, const ScoreState& s, FilterCursor& c)
{
    const int Nbin = F.grid_n - 1;
    int i = c.k;  // 0 = the piece in loc1 (when it exists), then the full bins, then the piece in loc2
    if (c.first) {
        if (i == 0) { c.idx = c.loc1; c.l = (g[c.loc1 + 1] - s.t_old) * s.speed; return; }
        i--;
    }
    const int num_bin = c.loc2 - c.loc1 - 1;
    if (i < num_bin) { c.idx = c.loc1 + i + 1; c.l = (g[c.loc1 + i + 2] - g[c.loc1 + i + 1]) * s.speed; return; }
    (void)Nbin;
    c.idx = c.loc2; c.l = (s.t - g[c.loc2]) * s.speed;
}
__device__ __forceinline__ void estimator_score_plain(const DevProblem& P, const TallyAcc& T, const mcb_estimator& E, const ScoreState& s,
                                                      double l_in, int hist, ChannelCache& CC)
{
    constexpr int MAXF = 4;
    FilterCursor cur[MAXF];
    int64_t factor[MAXF + 1];  // idx_factor (Estimator.cpp:288-295)
    const int nf = E.n_filters < MAXF ? E.n_filters : MAXF;
    factor[nf] = 1;
    for (int i = nf - 1; i >= 0; i--) factor[i] = factor[i + 1] * P.filters[E.filter_begin + i].size;
    for (int i = 0; i < nf; i++) {
        const mcb_filter F = P.filters[E.filter_begin + i];
        const double* g = P.filter_grid + F.grid_begin;
        FilterCursor& c = cur[i];
        c.k = 0; c.n = 1; c.l = l_in; c.first = false; c.loc1 = c.loc2 = 0;
        switch (F.type) {
        case MCB_FILTER_SURFACE: c.idx = mcb_binary_search((double)s.surface_old, g, F.grid_n) + 1; break;  // Estimator.cpp:133-140
        case MCB_FILTER_CELL: c.idx = mcb_binary_search((double)s.cell, g, F.grid_n) + 1; break;            // :141-148
        case MCB_FILTER_ENERGY:
        case MCB_FILTER_ENERGY_OLD:                                                                          // :149-179
            c.idx = mcb_binary_search(F.type == MCB_FILTER_ENERGY ? s.E : s.E_old, g, F.grid_n);
            if (c.idx < 0 || c.idx >= F.grid_n - 1) return;
            break;
        default: {                                                                                           // time :199-246
            const int Nbin = F.grid_n - 1;
            c.loc1 = mcb_binary_search(s.t_old, g, F.grid_n);
            c.loc2 = mcb_binary_search(s.t, g, F.grid_n);
            if (c.loc1 == c.loc2) {
                if (c.loc1 < 0 || c.loc1 >= Nbin) return;
                c.idx = c.loc1;
            } else {
                c.first = c.loc1 >= 0;
                c.n = (c.first ? 1 : 0) + (c.loc2 - c.loc1 - 1) + (c.loc2 < Nbin ? 1 : 0);
                if (c.n == 0) return;
                time_piece(F, g, s, c);
            }
        }
        }
    }
    double* acc = T.acc + (int64_t)(hist - T.first_hist);
    for (;;) {
        double l = MCB_MAX_FLOAT;
        int64_t idx_1D = 0;
        for (int i = 0; i < nf; i++) { l = fmin(l, cur[i].l); idx_1D += (int64_t)cur[i].idx * factor[i + 1]; }
        for (int k = 0; k < E.n_scores; k++) {
            const double v = score_value(P, P.scores[E.score_begin + k], s, l, CC);
            const int64_t t = E.tally_begin + idx_1D + (int64_t)k * factor[0];
            if (t >= E.tally_begin && t < E.tally_begin + E.n_tallies) atomicAdd(acc + t * T.stride, v);
        }
        for (int i = 0; i < nf; i++) {
            FilterCursor& c = cur[i];
            c.l -= l;
            if (c.l < MCB_EPSILON_FLOAT) {
                if (c.k == c.n - 1) return;
                c.k++;
                const mcb_filter F = P.filters[E.filter_begin + i];
                time_piece(F, P.filter_grid + F.grid_begin, s, c);
            }
        }
        if (nf == 0) return;
    }
}
// One estimator scores one event.  The TRMM estimators first let a COPY of the particle scatter / fission, drawing
// from the particle's own stream like the reference draws from its global one (EstimatorScatter / FissionPrompt /
// FissionDelayed::score, Estimator.cpp:441-482): energy_old = incident energy, energy and speed = outgoing.
__device__ __forceinline__ void estimator_score(const DevProblem& P, const TallyAcc& T, int e, const ScoreState& s, uint64_t& rng, double l,
                                             int hist, ChannelCache& CC)
{
    const mcb_estimator E = P.estimators[e];
    if (E.simulate == MCB_SIM_NONE) { estimator_score_plain(P, T, E, s, l, hist, CC); return; }
    if (s.material < 0) return;
    const DevMaterial& M = P.materials[s.material];
    ScoreState q = s;
    int n = -1;
    if (E.simulate == MCB_SIM_SCATTER) {
        (void)macro_channel(P, s.material, s.E, 0, false, mcb_urand(rng), &n, CC);
        if (n < 0) return;  // the reference dereferences a null nuclide here
        scatter_sample(P.nuclides[n].A, q.u, q.v, q.wd, q.E, q.speed, rng);
    } else if (E.simulate == MCB_SIM_FISSION || E.simulate == MCB_SIM_FISSION_PROMPT) {
        (void)macro_channel(P, s.material, s.E, E.simulate == MCB_SIM_FISSION ? 1 : 2, false, mcb_urand(rng), &n, CC);
        if (n < 0) return;
        const DevNuclide& N = P.nuclides[n];
        q.E = watt_sample(N.watt_a, N.watt_b, N.watt_g, s.E, rng);
        q.speed = mcb_speed_of_energy(q.E);
    } else {
        const int g = E.simulate - MCB_SIM_FISSION_DELAYED;
        (void)macro_channel(P, s.material, s.E, 3 + g, false, mcb_urand(rng), &n, CC);
        if (n < 0) return;
        q.E = chid_sample(P.nuclides[n], g, rng);
        q.speed = mcb_speed_of_energy(q.E);
    }
    q.E_old = s.E;  // Particle::set_energy / set_speed (Particle.cpp:42-56)
    // cross sections at the outgoing energy only when a score or kernel of this estimator reads them
    bool needs_X = false;
    for (int k = 0; k < E.n_scores; k++) {
        const mcb_score& S = P.scores[E.score_begin + k];
        if (S.kernel == MCB_KERNEL_COLLISION || (S.score >= MCB_SCORE_ABSORPTION && S.score <= MCB_SCORE_TOTAL)) needs_X = true;
    }
    if (needs_X) { q.uidx = union_index(M, q.E); macro_xs(P, M, q.uidx, q.E, q.X); }
    estimator_score_plain(P, T, E, q, l, hist, CC);
}
__device__ __forceinline__ void score_attached(const DevProblem& P, const TallyAcc& T, int kind, int id,
                                               const ScoreState& s, uint64_t& rng, double l, int hist)
{
    const int b = P.attach_begin[kind][id], e = P.attach_begin[kind][id + 1];
    ChannelCache CC;  // microscopic data at the (at most two) energies the estimators of this event ask about
    channel_cache_reset(CC);
    for (int i = b; i < e; i++) estimator_score(P, T, P.attach_list[kind][i], s, rng, l, hist, CC);
}
__device__ __forceinline__ bool has_attached(const DevProblem& P, int kind, int id)
{
    return P.attach_begin[kind][id + 1] > P.attach_begin[kind][id];
}

// ---------------------------------------------------------------------------------------------
// the events of one particle, on registers.  Stage kernels load/store the fields an event needs; k_finish keeps
// the whole particle in registers and chains the events.
// ---------------------------------------------------------------------------------------------
struct Particle {
    double x, y, z, u, v, w, E, speed, wgt, t;
    uint64_t rng;
    int cell, hist;
    int slot;  // bank slot the particle was loaded from (addresses Bank::Eold)
};
// all estimators attached to surface / cell `id` score one event of particle p.  Cold and out of line, with every
// input BY VALUE (no address of a register-resident particle escapes), so that the transport kernels' register
// allocation is not shaped by it; returns the particle's stream state (the simulating estimators draw from it).
// have_X = false (surface estimators): the cross sections of the cell the particle is now in are looked up here.
__device__ __noinline__ uint64_t score_event(const DevProblem& P, const Bank& B, const TallyAcc& T, int kind, int id, double w, double E,
                                             double speed, double t, double du, double dv, double dw, int cell, int hist, int slot, uint64_t rng,
                                             int material, int uidx, bool have_X, double Xt, double Xs, double Xc, double Xf, double Xnf,
                                             int surface_old, double l)
{
    ScoreState s;
    s.w = w; s.E = E; s.speed = speed; s.u = du; s.v = dv; s.wd = dw;
    s.E_old = P.track_old ? B.Eold[slot] : E;
    s.t = t; s.t_old = P.track_time ? B.told[slot] : t;
    s.cell = cell; s.surface_old = surface_old; s.material = material; s.uidx = uidx;
    s.X = MacroXS{Xt, Xs, Xc, Xf, Xnf};
    if (!have_X) {
        s.uidx = -1;
        if (material >= 0) { s.uidx = union_index(P.materials[material], E); macro_xs(P, P.materials[material], s.uidx, E, s.X); }
    }
    score_attached(P, T, kind, id, s, rng, l, hist);
    return rng;
}
#define MCB_SCORE_EVENT(kind, id, p, material, uidx, have_X, X, surface_old, l)                                              \
    (p).rng = score_event(P, B, T, kind, id, (p).wgt, (p).E, (p).speed, (p).t, (p).u, (p).v, (p).w, (p).cell, (p).hist, (p).slot, \
                          (p).rng, material, uidx, have_X, (X).t, (X).s, (X).c, (X).f, (X).nf, surface_old, l)

// xs_lookup event
template <bool DETAIL>
__device__ __forceinline__ bool ev_lookup(const DevProblem& P, const Particle& p, MacroXS& X, int& uidx, XSDetail* D)
{
    const int m = P.cells[p.cell].material;
    if (m < 0) return false;
    const DevMaterial M = P.materials[m];
    uidx = union_index(M, p.E);
    macro_xs_impl<DETAIL>(P, M, uidx, p.E, X, D);
    return true;
}

// flight event: surface_intersect + collision_distance + move_particle (general.cpp:40-83,177-207).
// Returns true when the flight ends on a surface (S_hit), false when it ends in a collision.
// A history whose particles never share it with others in flight (k-eigenvalue without splitting) can keep its
// EstimatorK scores and its site count in registers (HistLocal) and store them once when it ends; otherwise they
// are bumped in memory with reductions.
struct HistLocal { double kC, kTL; int nsite; };

template <bool TALLY>
__device__ __forceinline__ bool ev_flight(const DevProblem& P, const Bank& B, Particle& p, const MacroXS& X, int uidx, const HistoryAcc& H,
                                          const TallyAcc& T, int& S_hit, HistLocal* L = nullptr)
{
    const int m = P.cells[p.cell].material;
    double dsurf;
    S_hit = surface_intersect(P, p.cell, p.x, p.y, p.z, p.u, p.v, p.w, dsurf);
    double dcol;
    if (m >= 0) dcol = -log(mcb_urand(p.rng)) / X.t;   // exponential_sample (Algorithm.cpp:123-126)
    else dcol = MCB_MAX_FLOAT_LESS;                      // vacuum (general.cpp:44-46)
    const bool to_cross = dcol > dsurf;
    const double l = to_cross ? dsurf : dcol;
    // Particle::move (Particle.cpp:66-76)
    p.x += p.u * l; p.y += p.v * l; p.z += p.w * l;
    if (TALLY && P.track_time) B.told[p.slot] = p.t;
    p.t += l / p.speed;
    if (P.ksearch && m >= 0) {  // estimate_TL (Estimator.cpp:509-512)
        if (L) L->kTL += X.nf * p.wgt * l;
        else hist_add(&H.kTL[p.hist], X.nf * p.wgt * l);
    }
    if (TALLY && T.on && has_attached(P, MCB_ATTACH_CELL_TL, p.cell)) {
        MCB_SCORE_EVENT(MCB_ATTACH_CELL_TL, p.cell, p, m, uidx, true, X, -1, l);
    }
    return to_cross;
}

// collide event, first half: Simulator::collision up to the fission dispatch (general.cpp:121-150): collision
// tallies, bank_nu, fissioning nuclide, prompt/delayed draw.  Tells how many fission sites (k-eigenvalue) or
// same-history secondaries (fixed source) the second half will write.
struct CollideCtx {
    int m, N_fission;
    unsigned n_sites, n_second;
};
template <bool TALLY>
__device__ __forceinline__ bool ev_collide_pre(const DevProblem& P, const Bank& B, Particle& p, const MacroXS& X, int uidx, const XSDetail* D,
                                               const TallyAcc& T, double k_eff, CollideCtx& c)
{
    c.m = P.cells[p.cell].material;
    c.N_fission = -1; c.n_sites = 0; c.n_second = 0;
    if (c.m < 0) { p.wgt = 0.0; return false; }  // vacuum: kill (general.cpp:124-128)
    if (TALLY && T.on && has_attached(P, MCB_ATTACH_CELL_C, p.cell)) {
        MCB_SCORE_EVENT(MCB_ATTACH_CELL_C, p.cell, p, c.m, uidx, true, X, -1, 0.0);
    }
    // floor( w/k * nuSigmaF / SigmaT + xi ) (general.cpp:135-136)
    const double a = p.wgt / k_eff * X.nf / X.t;
    const double bn = floor(a + mcb_urand(p.rng));
    const unsigned bank_nu = bn > 0.0 ? (unsigned)bn : 0u;
    const DevMaterial& M = P.materials[c.m];
    int ln = 0;
    const double xi_f = mcb_urand(p.rng);
    c.N_fission = D ? select_from_detail(P, M, D->cum_nf, X.nf, xi_f, &ln)
                    : select_nuclide(P, M, uidx, p.E, 1, X.nf, xi_f, &ln);  // Material.cpp:116-125
    if (c.N_fission >= 0) {
        // prompt or delayed (ksearch.cpp:24-38, fixed_source.cpp:12,41-52)
        const double beta = D ? D->beta[ln] : micro_col(P.nuclides[c.N_fission], nuclide_index(M, uidx, ln), p.E, 1);
        const bool prompt = mcb_urand(p.rng) > beta;
        if (P.ksearch) {
            if (!prompt) (void)mcb_urand(p.rng);  // precursor group pick, result unused
            c.n_sites = bank_nu;
        } else if (prompt) {
            c.n_second = bank_nu;
        } else {
            // delayed, non-TDMC branch: draws are consumed, no neutron is banked (fixed_source.cpp:41-63)
            (void)mcb_urand(p.rng);
            for (unsigned b = 0; b < bank_nu; b++) { (void)mcb_urand(p.rng); (void)mcb_urand(p.rng); }
        }
    }
    return true;
}
// collide event, banking part.  k-eigenvalue (ksearch.cpp:39-46): one request per fission site goes to
// [site0, ..) of the request buffer; its Watt energy and isotropic direction are sampled by k_bank_sample_order
// from the request's own stream.  Fixed source (fixed_source.cpp:12-22): same-history secondaries are written to
// bank slots [slot0, ..), each sampled from, and continuing on, its own stream.  The parent's stream does not
// advance.  Only lanes that bank anything call this; callers reconverge the warp afterwards.
template <bool TALLY>
__device__ __forceinline__ void ev_collide_bank(const DevProblem& P, const Bank& B, const Particle& p, const CollideCtx& c,
                                                const HistoryAcc& H, Counters* C, SiteReq* reqs, uint64_t site_cap,
                                                uint32_t n_slots, unsigned long long site0, unsigned long long slot0,
                                                unsigned& n_second_ok, HistLocal* L = nullptr, unsigned long long ring_begin = 0)
{
    n_second_ok = 0;
    uint64_t seed = p.rng;
    if (c.n_sites) {
        int seq0;
        if (L) { seq0 = L->nsite; L->nsite += (int)c.n_sites; }
        else seq0 = atomicAdd(&H.nsite[p.hist], (int)c.n_sites);
        for (unsigned b = 0; b < c.n_sites; b++) {
            seed = (seed * MCB_RN_JUMP40) & MCB_RN_MASK;
            SiteReq r;
            r.x = p.x; r.y = p.y; r.z = p.z; r.t = p.t; r.E_in = p.E; r.seed = seed;
            r.cell = p.cell; r.seq = seq0 + (int)b; r.hist = p.hist; r.nuclide = c.N_fission;
            if (site0 + b < site_cap) reqs[site0 + b] = r;
            else C->overflow_sites = 1;
        }
    }
    if (c.n_second) {
        const DevNuclide& N = P.nuclides[c.N_fission];
        for (unsigned b = 0; b < c.n_second; b++) {
            seed = (seed * MCB_RN_JUMP40) & MCB_RN_MASK;
            uint64_t rs = seed;
            const double Es = watt_sample(N.watt_a, N.watt_b, N.watt_g, p.E, rs);   // energy first, then direction (App. D-4)
            double du, dv, dw;
            isotropic_direction(rs, du, dv, dw);
            // the bank is a ring: positions grow without bound, slots behind the running pass (ring_begin) are reused
            const unsigned long long pos = slot0 + b;
            if (pos - ring_begin < n_slots) {
                const uint32_t j = (uint32_t)(pos % n_slots);
                B.x[j] = p.x; B.y[j] = p.y; B.z[j] = p.z; B.u[j] = du; B.v[j] = dv; B.w[j] = dw;
                B.E[j] = Es; B.speed[j] = mcb_speed_of_energy(Es); B.wgt[j] = 1.0; B.t[j] = p.t;
                B.rng[j] = rs; B.cell[j] = p.cell; B.hist[j] = p.hist;
                if (TALLY && P.track_old) B.Eold[j] = Es;
                n_second_ok++;
            } else C->overflow_slots = 1;
        }
    }
}
// collide event, last part: k_C, implicit capture, scatter, weight_roulette (general.cpp:146-163,
// population_control.cpp:9-15).  Returns whether the particle survives.
template <bool TALLY>
__device__ __forceinline__ bool ev_collide_scatter(const DevProblem& P, const Bank& B, Particle& p, const MacroXS& X, int uidx, const XSDetail* D,
                                                   const CollideCtx& c, const HistoryAcc& H, HistLocal* L = nullptr)
{
    if (P.ksearch && c.N_fission >= 0) {  // estimate_C (Estimator.cpp:503-507)
        if (L) L->kC += X.nf * p.wgt / X.t;
        else hist_add(&H.kC[p.hist], X.nf * p.wgt / X.t);
    }
    // implicit absorption (general.cpp:154-156)
    const double implicit = X.c + X.f;
    p.wgt = p.wgt * (X.t - implicit) / X.t;
    const double xi_s = mcb_urand(p.rng);
    int ln_s = 0;
    const int N_scatter = D ? select_from_detail(P, P.materials[c.m], D->cum_s, X.s, xi_s, &ln_s)
                            : select_nuclide(P, P.materials[c.m], uidx, p.E, 0, X.s, xi_s, &ln_s);  // Material.cpp:106-115
    if (N_scatter >= 0) {
        if (TALLY && P.track_old) B.Eold[p.slot] = p.E;  // Particle::set_speed keeps the pre-collision energy (Particle.cpp:49-56)
        scatter_sample(P.nuclides[N_scatter].A, p.u, p.v, p.w, p.E, p.speed, p.rng);
    }
    // weight_roulette (population_control.cpp:9-15)
    if (p.wgt < P.wr) {
        if (mcb_urand(p.rng) < p.wgt / P.ws) p.wgt = P.ws;
        else { p.wgt = 0.0; return false; }
    }
    return true;
}

// cross event, first half: surface_hit + cell_importance up to the split (general.cpp:89-115,
// population_control.cpp:21-43).  n_copy = split copies the second half will write.
template <bool TALLY>
__device__ __forceinline__ bool ev_cross_pre(const DevProblem& P, const Bank& B, Particle& p, int S, const TallyAcc& T, Counters* C, unsigned& n_copy)
{
    n_copy = 0;
    if (S < 0) { p.wgt = 0.0; return false; }  // no surface ahead: cannot happen in a closed geometry
    const mcb_surface& Sf = P.surfaces[S];
    const int cell_old = p.cell;
    bool alive = true;
    if (Sf.bc == MCB_BC_TRANSMISSION) {
        p.x += p.u * MCB_EPSILON_FLOAT; p.y += p.v * MCB_EPSILON_FLOAT; p.z += p.w * MCB_EPSILON_FLOAT;
        if (TALLY && P.track_time) B.told[p.slot] = p.t;
        p.t += MCB_EPSILON_FLOAT / p.speed;
        const int cn = mcb_search_cell(P.cells, P.n_cells, P.surfaces, P.cell_surface, P.cell_sense, p.x, p.y, p.z);
        if (cn < 0) {  // "[WARNING] A particle is lost" (general.cpp:31-33)
            if (atomicExch(&C->lost, 1) == 0) { C->lost_pos[0] = p.x; C->lost_pos[1] = p.y; C->lost_pos[2] = p.z; }
            alive = false; p.wgt = 0.0;
        } else p.cell = cn;
    } else if (Sf.bc == MCB_BC_VACUUM) {
        alive = false; p.wgt = 0.0;
    } else {
        mcb_surf_reflect(Sf, p.u, p.v, p.w);
        p.x += p.u * MCB_EPSILON_FLOAT; p.y += p.v * MCB_EPSILON_FLOAT; p.z += p.w * MCB_EPSILON_FLOAT;
        if (TALLY && P.track_time) B.told[p.slot] = p.t;
        p.t += MCB_EPSILON_FLOAT / p.speed;
    }
    if (TALLY && T.on && has_attached(P, MCB_ATTACH_SURFACE, S)) {
        const MacroXS X0 = {0, 0, 0, 0, 0};
        MCB_SCORE_EVENT(MCB_ATTACH_SURFACE, S, p, P.cells[p.cell].material, -1, false, X0, S, 0.0);
    }
    const double Iold = P.cells[cell_old].importance, Inew = P.cells[p.cell].importance;
    if (Inew != Iold) {
        const double rat = Inew / Iold;
        if (rat < 1.0) {
            if (mcb_urand(p.rng) < rat) p.wgt = p.wgt / rat;
            else { alive = false; p.wgt = 0.0; }
        } else {
            const int ns = (int)floor(rat + mcb_urand(p.rng));
            p.wgt = p.wgt / (double)ns;
            n_copy = ns > 1 ? (unsigned)(ns - 1) : 0u;
        }
    }
    return alive;
}
// cross event, second half: the split copies (population_control.cpp:44-48) and weight_roulette, which also
// draws for a particle that was just killed (w = 0 < wr), like the reference
template <bool TALLY>
__device__ __forceinline__ bool ev_cross_post(const DevProblem& P, const Bank& B, Particle& p, bool alive, unsigned n_copy,
                                              unsigned long long slot0, uint32_t n_slots, Counters* C, unsigned& n_copy_ok,
                                              unsigned long long ring_begin = 0)
{
    n_copy_ok = 0;
    for (unsigned b = 0; b < n_copy; b++) {
        const unsigned long long pos = slot0 + b;
        if (pos - ring_begin < n_slots) {
            const uint32_t j = (uint32_t)(pos % n_slots);
            B.x[j] = p.x; B.y[j] = p.y; B.z[j] = p.z; B.u[j] = p.u; B.v[j] = p.v; B.w[j] = p.w;
            B.E[j] = p.E; B.speed[j] = p.speed; B.wgt[j] = p.wgt; B.t[j] = p.t;
            B.rng[j] = mcb_rn_child_seed(p.rng, b); B.cell[j] = p.cell; B.hist[j] = p.hist;
            if (TALLY && P.track_old) B.Eold[j] = B.Eold[p.slot];
            n_copy_ok++;
        } else C->overflow_slots = 1;
    }
    if (p.wgt < P.wr) {
        if (mcb_urand(p.rng) < p.wgt / P.ws) p.wgt = P.ws;
        else { p.wgt = 0.0; alive = false; }
    }
    return alive;
}

// ---------------------------------------------------------------------------------------------
// stage kernels
// ---------------------------------------------------------------------------------------------
// source: SourceBank::get_source (Source.cpp:42-46) with j = floor(xi*N), SourcePoint / SourceDelta (Source.cpp:16-24)
// History h (shard-local) of this cycle gets the stream of nps = cycle*Nsample + (shard_begin + h)
// (RN_init_particle, Random.cpp:196-204).
// Sorted sourcing (a bank spread over several GPUs): drawing sites by random index from peers' HBM thrashes the
// address translation of the peer mappings once the banks exceed ~1 GB (measured: 188 ms instead of 7 ms for 4e7
// draws from a 3.2 GB peer bank).  So the draws are made first (k_pick), sorted by site index (cub radix sort), and
// k_source then walks the bank in ascending order: slot q gets the history whose draw is the q-th smallest, every
// peer page is visited once, neighbouring threads read neighbouring (or the same) sites.
__global__ void __launch_bounds__(BLOCK)
k_pick(uint64_t seed0, uint64_t nps0, int32_t first_hist, uint32_t count, unsigned long long n_bank, unsigned long long rot,
       unsigned long long* key, uint32_t* val, uint64_t* rng_after)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ uint64_t s_base;  // stream of the block's first history; the others are a short skip away
    if (threadIdx.x == 0) s_base = mcb_rn_history_seed(seed0, nps0 + (uint64_t)(first_hist + (int32_t)(blockIdx.x * blockDim.x)));
    __syncthreads();
    if (q >= count) return;
    uint64_t rng = mcb_rn_history_seed_from(s_base, threadIdx.x);
    const double xi = mcb_urand(rng);
    unsigned long long j = (unsigned long long)(xi * (double)n_bank);
    if (j >= n_bank) j = n_bank - 1;
    // the sweep of rank r starts at its own slice (rot = global index of its first site) and wraps around: at any
    // moment the ranks read from different peers instead of all queueing at rank 0's HBM
    key[q] = j >= rot ? j - rot : j + n_bank - rot; val[q] = q; rng_after[q] = rng;
}

__global__ void __launch_bounds__(BLOCK)
k_source(const DevProblem P, Bank B, uint32_t* active, int32_t first_hist, uint32_t count, uint64_t nps0,
         const SourceBankView V, Counters* C, const unsigned long long* __restrict__ sorted_key,
         const uint32_t* __restrict__ sorted_val, const uint64_t* __restrict__ rng_after, unsigned long long rot, uint32_t q0)
{
    // q0 > 0: one chunk [q0, q0 + count) of a sorted sweep that is launched piece by piece (streamed host bank)
    const uint32_t q = q0 + blockIdx.x * blockDim.x + threadIdx.x;
    count += q0;
    if (q == 0) {  // queue state of the batch: `count` primaries in queue 0, slots behind them are free
        C->n_active[0] = count; C->n_active[1] = 0; C->n_active[2] = 0; C->q_collide = 0; C->q_cross = 0; C->slot_cursor = count;
    }
    __shared__ uint64_t s_base;  // stream of the block's first history; the others are a short skip away
    if (!sorted_key) {
        if (threadIdx.x == 0) s_base = mcb_rn_history_seed(P.seed0, nps0 + (uint64_t)(first_hist + (int32_t)(q0 + blockIdx.x * blockDim.x)));
        __syncthreads();
    }
    if (q >= count) return;
    int32_t h;
    uint64_t rng;
    double xi = 0.0;
    if (sorted_key) { h = first_hist + (int32_t)sorted_val[q]; rng = rng_after[sorted_val[q]]; }
    else { h = first_hist + (int32_t)q; rng = mcb_rn_history_seed_from(s_base, threadIdx.x); xi = mcb_urand(rng); }
    double x, y, z, u, v, w, E, t;
    int cell;
    if (V.n) {
        uint64_t j;
        if (sorted_key) { j = sorted_key[q] + rot; if (j >= V.n) j -= V.n; }
        else { j = (uint64_t)(xi * (double)V.n); if (j >= V.n) j = V.n - 1; }
        const Site s = source_bank_site(V, j);  // local HBM, or a peer's HBM over NVLink
        x = s.x; y = s.y; z = s.z; E = s.E; t = s.t; cell = s.cell;
        source_bank_direction(V, j, s, u, v, w);
    } else {
        int j = (int)(xi * (double)P.n_sources);
        if (j >= P.n_sources) j = P.n_sources - 1;
        const mcb_source& S = P.sources[j];
        // energy before direction: g++ evaluates the constructor arguments right to left
        E = dist1_sample(S.energy, rng);
        if (S.dir_kind == MCB_DIR_DELTA) { u = S.dir[0]; v = S.dir[1]; w = S.dir[2]; }
        else if (S.dir_kind == MCB_DIR_ISOTROPIC) isotropic_direction(rng, u, v, w);
        else { w = dist1_sample(S.dir_xyz[2], rng); v = dist1_sample(S.dir_xyz[1], rng); u = dist1_sample(S.dir_xyz[0], rng); }
        x = S.pos[0]; y = S.pos[1]; z = S.pos[2]; t = 0.0; cell = S.cell;
    }
    B.x[q] = x; B.y[q] = y; B.z[q] = z; B.u[q] = u; B.v[q] = v; B.w[q] = w;
    B.E[q] = E; B.speed[q] = mcb_speed_of_energy(E); B.wgt[q] = 1.0; B.t[q] = t;
    B.rng[q] = rng; B.cell[q] = cell; B.hist[q] = h;
    if (P.track_old) B.Eold[q] = E;  // the reference leaves energy_old uninitialised at birth; defined as E here
    active[q] = q;
}

// xs_lookup stage: macroscopic cross sections of every queued particle at its energy in its cell's material.
// As the first kernel of an iteration it also clears the queue lengths the iteration is going to fill.
__global__ void __launch_bounds__(BLOCK, 4)
k_xs_stage(const DevProblem P, Bank B, const uint32_t* __restrict__ active, int cur, Counters* C)
{
    const unsigned long long n = C->n_active[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) { C->n_active[(cur + 2) % 3] = 0; C->q_collide = 0; C->q_cross = 0; }
    unsigned looked = 0;
    for (unsigned long long tile = blockIdx.x; tile * BLOCK < n; tile += gridDim.x) {
        const unsigned long long q = tile * BLOCK + threadIdx.x;
        if (q < n) {
            const uint32_t i = active[q];
            Particle p;
            p.cell = B.cell[i]; p.E = B.E[i];
            MacroXS X;
            int u;
            if (ev_lookup<false>(P, p, X, u, nullptr)) {
                B.St[i] = X.t; B.Ss[i] = X.s; B.Sc[i] = X.c; B.Sf[i] = X.f; B.nSf[i] = X.nf;
                B.uidx[i] = u;
                looked++;
            }
        }
    }
    // one counter atomic per block
    for (int d = 16; d; d >>= 1) looked += __shfl_xor_sync(FULL, looked, d);
    __shared__ unsigned s_cnt[WARPS];
    if (lane_id() == 0) s_cnt[warp_id()] = looked;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
        for (int w = 0; w < WARPS; w++) tot += s_cnt[w];
        if (tot) atomicAdd(&C->n_lookups, (unsigned long long)tot);
    }
}

// flight stage.  The event queue is split in place: collisions from the front, surface hits from the back.
__global__ void __launch_bounds__(BLOCK, 4)
k_flight(const DevProblem P, Bank B, const uint32_t* __restrict__ active, int cur, uint32_t* evq, Counters* C,
         HistoryAcc H, TallyAcc T)
{
    __shared__ BlockScratch<2> scratch[2];
    const unsigned long long n = C->n_active[cur];
    unsigned tracks = 0;
    int it = 0;
    for (unsigned long long tile = blockIdx.x; tile * BLOCK < n; tile += gridDim.x, it++) {
        const unsigned long long q = tile * BLOCK + threadIdx.x;
        const bool valid = q < n;
        bool to_cross = false;
        uint32_t i = 0;
        if (valid) {
            i = active[q];
            Particle p;
            p.cell = B.cell[i]; p.hist = B.hist[i]; p.slot = (int)i;
            p.x = B.x[i]; p.y = B.y[i]; p.z = B.z[i]; p.u = B.u[i]; p.v = B.v[i]; p.w = B.w[i];
            p.E = B.E[i]; p.speed = B.speed[i]; p.wgt = B.wgt[i]; p.t = B.t[i]; p.rng = B.rng[i];
            MacroXS X;
            X.t = B.St[i]; X.nf = B.nSf[i];
            int uidx = 0;
            if (T.on) { X.s = B.Ss[i]; X.c = B.Sc[i]; X.f = B.Sf[i]; uidx = B.uidx[i]; }
            int S;
            to_cross = ev_flight<true>(P, B, p, X, uidx, H, T, S);
            B.x[i] = p.x; B.y[i] = p.y; B.z[i] = p.z; B.t[i] = p.t; B.rng[i] = p.rng; B.surf[i] = S;
            tracks++;
        }
        const unsigned cnt[2] = {valid && !to_cross ? 1u : 0u, valid && to_cross ? 1u : 0u};
        unsigned long long* const cur[2] = {&C->q_collide, &C->q_cross};
        unsigned long long pos[2];
        block_reserve<2>(scratch[it & 1], cnt, cur, pos);
        if (cnt[0]) evq[pos[0]] = i;
        if (cnt[1]) evq[n - 1 - pos[1]] = i;
    }
    for (int d = 16; d; d >>= 1) tracks += __shfl_xor_sync(FULL, tracks, d);
    if (lane_id() == 0 && tracks) atomicAdd(&C->n_tracks, (unsigned long long)tracks);
}

// collide stage
__global__ void __launch_bounds__(BLOCK, 4)
k_collide(const DevProblem P, Bank B, const uint32_t* __restrict__ evq, int cur, Counters* C, uint32_t* next, HistoryAcc H,
          TallyAcc T, SiteReq* reqs, uint64_t site_cap, uint32_t n_slots, double k_eff)
{
    __shared__ BlockScratch<2> scratchA[2];
    __shared__ BlockScratch<1> scratchB[2];
    const unsigned long long n = C->q_collide;
    unsigned long long* const next_len = &C->n_active[(cur + 1) % 3];
    unsigned collisions = 0;
    int it = 0;
    for (unsigned long long tile = blockIdx.x; tile * BLOCK < n; tile += gridDim.x, it++) {
        const unsigned long long q = tile * BLOCK + threadIdx.x;
        const bool valid = q < n;
        uint32_t i = 0;
        Particle p;
        MacroXS X = {0, 0, 0, 0, 0};
        CollideCtx c = {-1, -1, 0, 0};
        int uidx = -1;
        bool in_material = false;
        if (valid) {
            i = evq[q];
            p.cell = B.cell[i]; p.hist = B.hist[i]; p.slot = (int)i;
            p.x = B.x[i]; p.y = B.y[i]; p.z = B.z[i]; p.u = B.u[i]; p.v = B.v[i]; p.w = B.w[i];
            p.E = B.E[i]; p.speed = B.speed[i]; p.wgt = B.wgt[i]; p.t = B.t[i]; p.rng = B.rng[i];
            X.t = B.St[i]; X.s = B.Ss[i]; X.c = B.Sc[i]; X.f = B.Sf[i]; X.nf = B.nSf[i]; uidx = B.uidx[i];
            in_material = ev_collide_pre<true>(P, B, p, X, uidx, nullptr, T, k_eff, c);
            if (in_material) collisions++;
        }
        const unsigned cntA[2] = {c.n_sites, c.n_second};
        unsigned long long* const curA[2] = {&C->site_cursor, &C->slot_cursor};
        unsigned long long posA[2];
        block_reserve<2>(scratchA[it & 1], cntA, curA, posA);
        bool alive = false;
        unsigned n_second_ok = 0;
        if (c.n_sites | c.n_second) ev_collide_bank<true>(P, B, p, c, H, C, reqs, site_cap, n_slots, posA[0], posA[1], n_second_ok);
        __syncwarp();  // the banking lanes rejoin the warp before the scatter kinematics
        if (in_material) alive = ev_collide_scatter<true>(P, B, p, X, uidx, nullptr, c, H);
        __syncwarp();
        if (valid) {
            B.u[i] = p.u; B.v[i] = p.v; B.w[i] = p.w; B.E[i] = p.E; B.speed[i] = p.speed; B.wgt[i] = p.wgt; B.rng[i] = p.rng;
        }
        // survivors and their secondaries go to the next iteration's queue
        const unsigned cntB[1] = {(alive ? 1u : 0u) + n_second_ok};
        unsigned long long* const curB[1] = {next_len};
        unsigned long long posB[1];
        block_reserve<1>(scratchB[it & 1], cntB, curB, posB);
        unsigned long long o = posB[0];
        if (alive) next[o++] = i;
        for (unsigned b = 0; b < n_second_ok; b++) next[o++] = (uint32_t)(posA[1] + b);
    }
    for (int d = 16; d; d >>= 1) collisions += __shfl_xor_sync(FULL, collisions, d);
    if (lane_id() == 0 && collisions) atomicAdd(&C->n_collisions, (unsigned long long)collisions);
}

// cross stage
__global__ void __launch_bounds__(BLOCK, 4)
k_cross(const DevProblem P, Bank B, const uint32_t* __restrict__ evq, int cur, Counters* C, uint32_t* next, TallyAcc T,
        uint32_t n_slots)
{
    __shared__ BlockScratch<1> scratchA[2];
    __shared__ BlockScratch<1> scratchB[2];
    const unsigned long long n = C->q_cross;
    const unsigned long long n_active = C->n_active[cur];
    unsigned long long* const next_len = &C->n_active[(cur + 1) % 3];
    unsigned crossings = 0;
    int it = 0;
    for (unsigned long long tile = blockIdx.x; tile * BLOCK < n; tile += gridDim.x, it++) {
        const unsigned long long q = tile * BLOCK + threadIdx.x;
        const bool valid = q < n;
        uint32_t i = 0;
        Particle p;
        bool alive = false;
        unsigned n_copy = 0;
        if (valid) {
            i = evq[n_active - 1 - q];
            p.cell = B.cell[i]; p.hist = B.hist[i]; p.slot = (int)i;
            p.x = B.x[i]; p.y = B.y[i]; p.z = B.z[i]; p.u = B.u[i]; p.v = B.v[i]; p.w = B.w[i];
            p.E = B.E[i]; p.speed = B.speed[i]; p.wgt = B.wgt[i]; p.t = B.t[i]; p.rng = B.rng[i];
            alive = ev_cross_pre<true>(P, B, p, B.surf[i], T, C, n_copy);
            crossings++;
        }
        unsigned long long slot0 = 0;
        if (P.shared_histories) {  // uniform: only problems with splitting reserve slots here
            const unsigned cntA[1] = {n_copy};
            unsigned long long* const curA[1] = {&C->slot_cursor};
            unsigned long long posA[1];
            block_reserve<1>(scratchA[it & 1], cntA, curA, posA);
            slot0 = posA[0];
        }
        unsigned n_copy_ok = 0;
        if (valid) {
            alive = ev_cross_post<true>(P, B, p, alive, n_copy, slot0, n_slots, C, n_copy_ok);
            if (alive) {
                B.x[i] = p.x; B.y[i] = p.y; B.z[i] = p.z; B.u[i] = p.u; B.v[i] = p.v; B.w[i] = p.w; B.t[i] = p.t;
                B.wgt[i] = p.wgt; B.rng[i] = p.rng; B.cell[i] = p.cell;
            }
        }
        const unsigned cntB[1] = {(alive ? 1u : 0u) + n_copy_ok};
        unsigned long long* const curB[1] = {next_len};
        unsigned long long posB[1];
        block_reserve<1>(scratchB[it & 1], cntB, curB, posB);
        unsigned long long o = posB[0];
        if (alive) next[o++] = i;
        for (unsigned b = 0; b < n_copy_ok; b++) next[o++] = (uint32_t)(slot0 + b);
    }
    for (int d = 16; d; d >>= 1) crossings += __shfl_xor_sync(FULL, crossings, d);
    if (lane_id() == 0 && crossings) atomicAdd(&C->n_crossings, (unsigned long long)crossings);
}

#ifndef MCB_STEP_MINB
#define MCB_STEP_MINB 6
#endif

// history walk: the whole event chain of a particle in registers, one launch per pass over bank slots
// [begin, end).  Every warp is autonomous (no block barriers): it draws slot indices in private chunks from a
// global head counter and, at every iteration, refills the lanes whose particle has just ended with the next slots,
// so lanes stay busy until the pass runs dry; fission-site requests are reserved with one cursor atomic per warp and
// iteration.  Nothing but the site requests (and the rare secondaries, which go to slots >= end and are walked by
// the next pass) is written back: the particle record is read once.  With one particle per history the EstimatorK
// scores live in registers and are stored once when the history ends.
template <bool TALLY, bool SHARED>
__global__ void __launch_bounds__(BLOCK, MCB_STEP_MINB)
k_walk(const DevProblem P, Bank B, unsigned long long begin, unsigned long long end, uint32_t chunk, Counters* C, HistoryAcc H, TallyAcc T,
       SiteReq* reqs, uint64_t site_cap, uint32_t n_slots, double k_eff)
{
    const unsigned lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    constexpr bool local_acc = !SHARED;  // SHARED = DevProblem::shared_histories: secondaries / split copies can exist
    unsigned tracks = 0, collisions = 0, crossings = 0, lookups = 0;
    bool alive = false, exhausted = false;
    Particle p;
    HistLocal L = {0.0, 0.0, 0};
    unsigned long long chunk_next = 0, chunk_end = 0;  // warp-uniform: this warp's private range of bank positions
    for (;;) {
        // ---- refill the idle lanes
        unsigned idle = __ballot_sync(FULL, !alive);
        while (idle && !exhausted) {
            if (chunk_next == chunk_end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&C->walk_head, (unsigned long long)chunk);
                base = __shfl_sync(FULL, base, 0);
                const unsigned long long b = begin + base;
                chunk_next = b < end ? b : end;
                chunk_end = b + chunk < end ? b + chunk : end;
                if (chunk_next == chunk_end) { exhausted = true; break; }
            }
            const unsigned take = min((unsigned)__popc(idle), (unsigned)(chunk_end - chunk_next));
            const unsigned rank = __popc(idle & lt_mask);
            if (!alive && rank < take) {
                // ring: position -> slot (without secondaries positions never leave [0, n_slots))
                const uint32_t j = SHARED ? (uint32_t)((chunk_next + rank) % n_slots) : (uint32_t)(chunk_next + rank);
                p.cell = B.cell[j]; p.hist = B.hist[j]; p.slot = (int)j;
                p.x = B.x[j]; p.y = B.y[j]; p.z = B.z[j]; p.u = B.u[j]; p.v = B.v[j]; p.w = B.w[j];
                p.E = B.E[j]; p.speed = B.speed[j]; p.wgt = B.wgt[j]; p.t = B.t[j]; p.rng = B.rng[j];
                L.kC = 0.0; L.kTL = 0.0; L.nsite = 0;
                alive = true;
            }
            chunk_next += take;
            idle = __ballot_sync(FULL, !alive);
        }
#ifdef MCB_WALK_SYNC
        // keep the warps of a block in phase: they then run the same stretch of this (large) loop body and share
        // its instruction-cache lines instead of evicting each other's
        if (!__syncthreads_or(idle != FULL)) break;
#else
        if (idle == FULL) break;  // nothing in flight and nothing left to draw
#endif
        // ---- one event per live lane
        MacroXS X = {0, 0, 0, 0, 0};
        XSDetail D;
        CollideCtx c = {-1, -1, 0, 0};
        int uidx = -1;
        unsigned n_copy = 0;
        bool to_cross = false, in_material = false;
        const bool was_alive = alive;
        if (alive) {
            int S;
            if (ev_lookup<true>(P, p, X, uidx, &D)) lookups++;
            to_cross = ev_flight<TALLY>(P, B, p, X, uidx, H, T, S, local_acc ? &L : nullptr);
            tracks++;
            if (to_cross) { alive = ev_cross_pre<TALLY>(P, B, p, S, T, C, n_copy); crossings++; }
            else { in_material = ev_collide_pre<TALLY>(P, B, p, X, uidx, &D, T, k_eff, c); if (in_material) collisions++; else alive = false; }
        }
        __syncwarp();
        // fission-site requests: one reservation per warp
        unsigned long long site0 = 0;
        {
            unsigned v = c.n_sites;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(FULL, v, d); if (lane >= (unsigned)d) v += t; }
            const unsigned total = __shfl_sync(FULL, v, 31);
            if (total) {
                unsigned long long base = 0;
                if (lane == 31) base = atomicAdd(&C->site_cursor, (unsigned long long)total);
                site0 = __shfl_sync(FULL, base, 31) + (v - c.n_sites);
            }
        }
        // secondaries (fixed-source fission neutrons, split copies) are rare: one slot reservation per lane
        unsigned long long slot0 = 0;
        if (!SHARED) { c.n_second = 0; n_copy = 0; }  // k-eigenvalue without splitting: nothing is ever born in flight
        else if (c.n_second + n_copy) slot0 = atomicAdd(&C->slot_cursor, (unsigned long long)(c.n_second + n_copy));
        unsigned n_new = 0;
        if (c.n_sites | c.n_second) ev_collide_bank<TALLY>(P, B, p, c, H, C, reqs, site_cap, n_slots, site0, slot0, n_new, local_acc ? &L : nullptr, begin);
        __syncwarp();  // the banking lanes rejoin the warp before the scatter kinematics
        if (to_cross) alive = ev_cross_post<TALLY>(P, B, p, alive, n_copy, slot0, n_slots, C, n_new, begin);
        __syncwarp();
        if (in_material) alive = ev_collide_scatter<TALLY>(P, B, p, X, uidx, &D, c, H, local_acc ? &L : nullptr);
        __syncwarp();
        if (local_acc && was_alive && !alive) {  // end of the history: EstimatorK::end_history inputs (Estimator.cpp:514-525)
            H.kC[p.hist] = L.kC; H.kTL[p.hist] = L.kTL; H.nsite[p.hist] = L.nsite;
        }
    }
    for (int d = 16; d; d >>= 1) {
        tracks += __shfl_xor_sync(FULL, tracks, d); collisions += __shfl_xor_sync(FULL, collisions, d);
        crossings += __shfl_xor_sync(FULL, crossings, d); lookups += __shfl_xor_sync(FULL, lookups, d);
    }
    if (lane == 0) {
        if (tracks) atomicAdd(&C->n_tracks, (unsigned long long)tracks);
        if (collisions) atomicAdd(&C->n_collisions, (unsigned long long)collisions);
        if (crossings) atomicAdd(&C->n_crossings, (unsigned long long)crossings);
        if (lookups) atomicAdd(&C->n_lookups, (unsigned long long)lookups);
    }
}

// tail of a batch: every queued particle is followed to the end of its history in registers (the same events,
// chained).  Secondaries born here are queued for another pass.  Cursors are bumped per thread: with a few
// thousand particles left there is no contention to aggregate away.
__global__ void __launch_bounds__(BLOCK, 4)
k_finish(const DevProblem P, Bank B, const uint32_t* __restrict__ active, int cur, Counters* C, uint32_t* next, HistoryAcc H,
         TallyAcc T, SiteReq* reqs, uint64_t site_cap, uint32_t n_slots, double k_eff)
{
    const unsigned long long n = C->n_active[cur];
    unsigned long long* const next_len = &C->n_active[(cur + 1) % 3];
    unsigned long long tracks = 0, collisions = 0, crossings = 0, lookups = 0;
    for (unsigned long long q = (unsigned long long)blockIdx.x * BLOCK + threadIdx.x; q < n; q += (unsigned long long)gridDim.x * BLOCK) {
        const uint32_t i = active[q];
        Particle p;
        p.cell = B.cell[i]; p.hist = B.hist[i]; p.slot = (int)i;
        p.x = B.x[i]; p.y = B.y[i]; p.z = B.z[i]; p.u = B.u[i]; p.v = B.v[i]; p.w = B.w[i];
        p.E = B.E[i]; p.speed = B.speed[i]; p.wgt = B.wgt[i]; p.t = B.t[i]; p.rng = B.rng[i];
        bool alive = true;
        while (alive) {
            MacroXS X = {0, 0, 0, 0, 0};
            XSDetail D;
            int uidx = -1, S;
            if (ev_lookup<true>(P, p, X, uidx, &D)) lookups++;
            const bool to_cross = ev_flight<true>(P, B, p, X, uidx, H, T, S);
            tracks++;
            unsigned n_new = 0;
            unsigned long long slot0 = 0;
            if (to_cross) {
                unsigned n_copy;
                alive = ev_cross_pre<true>(P, B, p, S, T, C, n_copy);
                crossings++;
                if (n_copy) slot0 = atomicAdd(&C->slot_cursor, (unsigned long long)n_copy);
                alive = ev_cross_post<true>(P, B, p, alive, n_copy, slot0, n_slots, C, n_new);
            } else {
                CollideCtx c;
                alive = ev_collide_pre<true>(P, B, p, X, uidx, &D, T, k_eff, c);
                if (alive) {
                    collisions++;
                    unsigned long long site0 = 0;
                    if (c.n_sites) site0 = atomicAdd(&C->site_cursor, (unsigned long long)c.n_sites);
                    if (c.n_second) slot0 = atomicAdd(&C->slot_cursor, (unsigned long long)c.n_second);
                    if (c.n_sites | c.n_second) ev_collide_bank<true>(P, B, p, c, H, C, reqs, site_cap, n_slots, site0, slot0, n_new);
                    alive = ev_collide_scatter<true>(P, B, p, X, uidx, &D, c, H);
                }
            }
            if (n_new) {
                const unsigned long long o = atomicAdd(next_len, (unsigned long long)n_new);
                for (unsigned b = 0; b < n_new; b++) next[o + b] = (uint32_t)(slot0 + b);
            }
        }
    }
    if (tracks) atomicAdd(&C->n_tracks, tracks);
    if (collisions) atomicAdd(&C->n_collisions, collisions);
    if (crossings) atomicAdd(&C->n_crossings, crossings);
    if (lookups) atomicAdd(&C->n_lookups, lookups);
}

// ---------------------------------------------------------------------------------------------
// generation close-out
// ---------------------------------------------------------------------------------------------
// fission bank in canonical order (parent history, banking order): position = offset[hist] + seq.  The site's
// energy (Watt spectrum of the fissioning nuclide at the incident energy) and isotropic direction are sampled
// here, one thread per site from the site's own stream (ksearch.cpp:41-46: energy first, then direction).
__global__ void __launch_bounds__(256)
k_bank_sample_order(const DevProblem P, const SiteReq* __restrict__ reqs, uint64_t n, const uint32_t* __restrict__ offset, Site* out)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const SiteReq r = reqs[q];
    const DevNuclide& N = P.nuclides[r.nuclide];
    uint64_t rng = r.seed;
    Site s;
    s.E = watt_sample(N.watt_a, N.watt_b, N.watt_g, r.E_in, rng);
    s.mu = 2.0 * mcb_urand(rng) - 1.0;  // the two draws of the isotropic direction (rebuilt when the site is used)
    s.xi = mcb_urand(rng);
    s.x = r.x; s.y = r.y; s.z = r.z; s.t = r.t; s.cell = r.cell; s.pad = 0;
    store_site(out + ((uint64_t)offset[r.hist] + (uint64_t)r.seq), s);
}

__device__ __forceinline__ int entropy_bin(const DevProblem& P, double x, double y, double z)  // Entropy.cpp:27-35
{
    const double* gx = P.entropy_grid;
    const double* gy = gx + P.entropy_n[0];
    const double* gz = gy + P.entropy_n[1];
    const int Iy = P.entropy_n[1] - 1, Iz = P.entropy_n[2] - 1;
    const int ix = mcb_binary_search(x, gx, P.entropy_n[0]);
    const int iy = mcb_binary_search(y, gy, P.entropy_n[1]);
    const int iz = mcb_binary_search(z, gz, P.entropy_n[2]);
    const int idx = ix * (Iz * Iy) + iy * Iz + iz;
    return (idx < 0 || idx >= P.entropy_bins) ? -1 : idx;
}

// exact accumulation of a non-negative double into a two-limb fixed-point sum
__device__ __forceinline__ void fx_split(double v, unsigned long long& lo, unsigned long long& hi, int* overflow)
{
    const double s = v * MCB_FX_SCALE;
    if (!(s < 4.0e18) || s < 0.0) { *overflow = 1; return; }
    const unsigned long long f = (unsigned long long)__double2ll_rn(s);
    lo += f & 0xffffffffull;
    hi += f >> 32;
}
__device__ __forceinline__ void fx_block_add(unsigned long long lo, unsigned long long hi, unsigned long long* dlo,
                                             unsigned long long* dhi)
{
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        lo += __shfl_xor_sync(FULL, lo, d);
        hi += __shfl_xor_sync(FULL, hi, d);
    }
    if (lane_id() == 0 && (lo | hi)) { atomicAdd(dlo, lo); atomicAdd(dhi, hi); }
}

// EstimatorK::end_history (Estimator.cpp:514-525) for all histories of the shard: sums and squares of k_C, k_TL
__global__ void __launch_bounds__(256)
k_reduce_k(const double* __restrict__ kC, const double* __restrict__ kTL, uint32_t n, Counters* C)
{
    unsigned long long lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    int ovf = 0;
    for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n; h += gridDim.x * blockDim.x) {
        const double c = kC[h], t = kTL[h];
        fx_split(c, lo[0], hi[0], &ovf);
        fx_split(t, lo[1], hi[1], &ovf);
        fx_split(c * c, lo[2], hi[2], &ovf);
        fx_split(t * t, lo[3], hi[3], &ovf);
    }
    for (int j = 0; j < 4; j++) fx_block_add(lo[j], hi[j], &C->fx_lo[j], &C->fx_hi[j]);
    if (ovf) C->overflow_fixed = 1;
}

// ShannonEntropy per history (Estimator.cpp:514-518 + Entropy.cpp:43-62): the bins a history touched
// are recovered from its slice of the canonical fission bank (bank_nu sites per banking collision)
__global__ void __launch_bounds__(256)
k_entropy_history(const DevProblem P, const Site* __restrict__ bank, const uint32_t* __restrict__ offset,
                  const int32_t* __restrict__ nsite, uint32_t n_hist, Counters* C)
{
    unsigned long long lo = 0, hi = 0;
    int ovf = 0;
    for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < n_hist; h += gridDim.x * blockDim.x) {
        const int n = nsite[h];
        if (n <= 0) continue;
        const Site* s = bank + offset[h];
        int total = 0;
        for (int a = 0; a < n; a++) if (entropy_bin(P, s[a].x, s[a].y, s[a].z) >= 0) total++;
        if (total == 0) continue;
        double H = 0.0;
        for (int a = 0; a < n; a++) {
            const int ba = entropy_bin(P, s[a].x, s[a].y, s[a].z);
            if (ba < 0) continue;
            bool first = true;
            int c = 0;
            for (int b = 0; b < n; b++) {
                const int bb = entropy_bin(P, s[b].x, s[b].y, s[b].z);
                if (bb == ba) { if (b < a) { first = false; break; } c++; }
            }
            if (!first) continue;
            const double p = (double)c / (double)total;
            H -= p * log2(p);
        }
        if (H > 0.0) fx_split(H, lo, hi, &ovf);
    }
    fx_block_add(lo, hi, &C->fx_lo[4], &C->fx_hi[4]);
    if (ovf) C->overflow_fixed = 1;
}
// whole-generation source histogram over the entropy mesh (conventional Shannon entropy, extra output)
__global__ void __launch_bounds__(256)
k_entropy_histogram(const DevProblem P, const Site* __restrict__ bank, uint64_t n, unsigned long long* bins)
{
    for (uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (uint64_t)gridDim.x * blockDim.x) {
        const int b = entropy_bin(P, bank[q].x, bank[q].y, bank[q].z);
        if (b >= 0) atomicAdd(&bins[b], 1ull);
    }
}

// Estimator::end_history (Estimator.cpp:339-346) over a batch: per tally bin, sum and sum of squares of the
// per-history scores, in two deterministic passes; the accumulator rows are cleared for the next batch
constexpr int TALLY_CHUNK = 8192;
__global__ void __launch_bounds__(256)
k_tally_partial(double* acc, int64_t stride, uint32_t n_hist, double* partial /* [tally][chunk][2] */, int n_chunks)
{
    __shared__ double ss[8], sq[8];
    const int tally = blockIdx.y, chunk = blockIdx.x;
    double* row = acc + (int64_t)tally * stride;
    double s = 0.0, q = 0.0;
    const uint32_t b = (uint32_t)chunk * TALLY_CHUNK;
    const uint32_t e = min(b + (uint32_t)TALLY_CHUNK, n_hist);
    for (uint32_t h = b + threadIdx.x; h < e; h += blockDim.x) {
        const double v = row[h];
        if (v != 0.0) { s += v; q += v * v; row[h] = 0.0; }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) { s += __shfl_xor_sync(FULL, s, d); q += __shfl_xor_sync(FULL, q, d); }
    if (lane_id() == 0) { ss[threadIdx.x >> 5] = s; sq[threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; i++) { s += ss[i]; q += sq[i]; }
        partial[((int64_t)tally * n_chunks + chunk) * 2 + 0] = s;
        partial[((int64_t)tally * n_chunks + chunk) * 2 + 1] = q;
    }
}
__global__ void k_tally_final(const double* partial, int n_chunks, int64_t n_tallies, double* sum, double* squared)
{
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tallies) return;
    double s = 0.0, q = 0.0;
    for (int c = 0; c < n_chunks; c++) { s += partial[(t * n_chunks + c) * 2]; q += partial[(t * n_chunks + c) * 2 + 1]; }
    sum[t] += s;
    squared[t] += q;
}

// host-facing bank layout (n x 8 doubles + n cells) <-> Site records
__global__ void __launch_bounds__(256)
k_pack_sites(const double* __restrict__ s8, const int32_t* __restrict__ cells, uint64_t n, Site* out, double* dir_x)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const double* s = s8 + 8 * q;
    Site d;
    d.x = s[0]; d.y = s[1]; d.z = s[2]; d.E = s[6]; d.t = s[7]; d.mu = 0.0; d.xi = 0.0;
    d.cell = cells[q]; d.pad = 0;
    dir_x[3 * q] = s[3]; dir_x[3 * q + 1] = s[4]; dir_x[3 * q + 2] = s[5];  // explicit directions ride beside the sites
    store_site(out + q, d);
}
__global__ void __launch_bounds__(256)
k_unpack_sites(const Site* __restrict__ in, const double* __restrict__ dir_x, uint64_t n, double* s8, int32_t* cells)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const Site d = load_site(in + q);
    double* s = s8 + 8 * q;
    double u, v, w;
    if (dir_x) { u = dir_x[3 * q]; v = dir_x[3 * q + 1]; w = dir_x[3 * q + 2]; }
    else direction_from_draws(d.mu, d.xi, u, v, w);
    s[0] = d.x; s[1] = d.y; s[2] = d.z; s[3] = u; s[4] = v; s[5] = w; s[6] = d.E; s[7] = d.t;
    cells[q] = d.cell;
}

// the global bank of a multi-GPU run, materialised on demand (host read-back): out[q] = site q of the view
__global__ void __launch_bounds__(256)
k_gather_sites(const SourceBankView V, uint64_t n, Site* out)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) store_site(out + q, source_bank_site(V, q));
}

__global__ void k_iota(uint32_t* a, uint32_t n)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) a[q] = q;
}

// ---------------------------------------------------------------------------------------------
// parity / bench kernels on plain arrays
// ---------------------------------------------------------------------------------------------
// Material::Sigma{T,S,C,F}, nuSigmaF at n energies; out = n x 5 (row-major), staged through shared memory so
// that the 40-byte records leave the SM as full 128-byte lines
__global__ void __launch_bounds__(256)
k_xs_lookup(const DevProblem P, int material, const double* __restrict__ E, int64_t n, double* __restrict__ out5)
{
    __shared__ double stage[256 * 5];
    const int64_t base = (int64_t)blockIdx.x * 256;
    const int64_t q = base + threadIdx.x;
    if (q < n) {
        const double e = __ldg(&E[q]);
        const DevMaterial M = P.materials[material];
        const int u = union_index(M, e);
        MacroXS X;
        macro_xs(P, M, u, e, X);
        double* s = stage + threadIdx.x * 5;
        s[0] = X.t; s[1] = X.s; s[2] = X.c; s[3] = X.f; s[4] = X.nf;
    }
    __syncthreads();
    const int64_t left = n - base;
    const int cnt = (int)(left < 256 ? left : 256) * 5;
    for (int j = threadIdx.x; j < cnt; j += 256) out5[base * 5 + j] = stage[j];
}
__global__ void k_select_channel(const DevProblem P, int material, int kind, const double* E, const double* xi,
                                 int64_t n, int32_t* out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const DevMaterial M = P.materials[material];
    const int u = union_index(M, E[q]);
    MacroXS X;
    macro_xs(P, M, u, E[q], X);
    out[q] = select_nuclide(P, M, u, E[q], kind, kind == 0 ? X.s : X.nf, xi[q], nullptr);
}
__global__ void k_beta(const DevProblem P, int material, int local_n, const double* E, int64_t n, double* out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const DevMaterial M = P.materials[material];
    const int u = union_index(M, E[q]);
    const int gn = P.mat_nuclide[M.nuc_begin + local_n];
    out[q] = micro_col(P.nuclides[gn], nuclide_index(M, u, local_n), E[q], 1);
}
__global__ void k_rng(uint64_t seed0, const uint64_t* nps, int64_t n, int ndraw, uint64_t* out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint64_t s = mcb_rn_history_seed(seed0, nps[q]);
    for (int d = 0; d < ndraw; d++) { (void)mcb_urand(s); out[q * ndraw + d] = s; }
}
__global__ void k_geometry(const DevProblem P, const int32_t* cell, const double* pos, const double* dir, int64_t n,
                           double* out3)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    double d;
    const int S = surface_intersect(P, cell[q], pos[3 * q], pos[3 * q + 1], pos[3 * q + 2], dir[3 * q], dir[3 * q + 1],
                                    dir[3 * q + 2], d);
    out3[3 * q] = d;
    out3[3 * q + 1] = (double)S;
    out3[3 * q + 2] = S >= 0 ? mcb_surf_eval(P.surfaces[S], pos[3 * q], pos[3 * q + 1], pos[3 * q + 2]) : 0.0;
}
__global__ void k_search_cell(const DevProblem P, const double* pos, int64_t n, int32_t* out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    out[q] = mcb_search_cell(P.cells, P.n_cells, P.surfaces, P.cell_surface, P.cell_sense, pos[3 * q], pos[3 * q + 1],
                             pos[3 * q + 2]);
}
__global__ void k_scatter(const DevProblem P, int nuclide, const uint64_t* nps, int64_t n, double* io5)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint64_t rng = mcb_rn_history_seed(P.seed0, nps[q]);
    double u = io5[5 * q], v = io5[5 * q + 1], w = io5[5 * q + 2], E = io5[5 * q + 3];
    double speed = mcb_speed_of_energy(E);
    scatter_sample(P.nuclides[nuclide].A, u, v, w, E, speed, rng);
    io5[5 * q] = u; io5[5 * q + 1] = v; io5[5 * q + 2] = w; io5[5 * q + 3] = E; io5[5 * q + 4] = speed;
}
__global__ void k_watt(const DevProblem P, int nuclide, const uint64_t* nps, const double* E, int64_t n, double* out)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint64_t rng = mcb_rn_history_seed(P.seed0, nps[q]);
    const DevNuclide& N = P.nuclides[nuclide];
    out[q] = watt_sample(N.watt_a, N.watt_b, N.watt_g, E[q], rng);
}

inline unsigned blocks_for(uint64_t n, unsigned bs = 256) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
namespace mcbk {

static thread_local uint64_t g_launches = 0;
uint64_t launch_count() { return g_launches; }
#define MCB_LAUNCHED(k) (g_launches += (k))

static unsigned grid_for(uint64_t n_hint)
{
    // persistent tile loops: enough blocks to fill the machine a few times over, never more than the work
    const uint64_t need = (n_hint + BLOCK - 1) / BLOCK;
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(need, 148ull * 16ull * (256 / BLOCK)));
}
void source(cudaStream_t st, const DevProblem& P, const Bank& B, uint32_t* active, int32_t first_hist, uint32_t count,
            uint64_t nps0, const SourceBankView& V, Counters* C, const SortScratch* sort)
{
    if (sort && V.n) {
        // draws -> sorted by site index -> the bank is read in ascending order
        k_pick<<<std::max(1u, blocks_for(count, BLOCK)), BLOCK, 0, st>>>(P.seed0, nps0, first_hist, count, V.n, sort->rot, sort->key_in, sort->val_in, sort->rng_after);
        int bits = 1;
        while (bits < 64 && (V.n >> bits)) bits++;
        size_t tb = sort->temp_bytes;
        cub::DeviceRadixSort::SortPairs(sort->temp, tb, sort->key_in, sort->key_out, sort->val_in, sort->val_out, (int)count, 0, bits, st);
        k_source<<<std::max(1u, blocks_for(count, BLOCK)), BLOCK, 0, st>>>(P, B, active, first_hist, count, nps0, V, C, sort->key_out,
                                                                          sort->val_out, sort->rng_after, sort->rot, 0u);
        MCB_LAUNCHED(2 + (bits + 7) / 8 + 2);  // pick, source, the sort's histogram / onesweep passes
        return;
    }
    k_source<<<std::max(1u, blocks_for(count, BLOCK)), BLOCK, 0, st>>>(P, B, active, first_hist, count, nps0, V, C, nullptr, nullptr, nullptr, 0ull, 0u);
    MCB_LAUNCHED(1);
}
// the two halves of the sorted path on their own, for a bank that arrives from the host in chunks: draws + sort
// first, then one k_source launch per chunk of the sweep
void pick_sort(cudaStream_t st, const DevProblem& P, int32_t first_hist, uint32_t count, uint64_t nps0, uint64_t n_bank,
               const SortScratch* sort)
{
    k_pick<<<std::max(1u, blocks_for(count, BLOCK)), BLOCK, 0, st>>>(P.seed0, nps0, first_hist, count, n_bank, sort->rot, sort->key_in, sort->val_in, sort->rng_after);
    int bits = 1;
    while (bits < 64 && (n_bank >> bits)) bits++;
    size_t tb = sort->temp_bytes;
    cub::DeviceRadixSort::SortPairs(sort->temp, tb, sort->key_in, sort->key_out, sort->val_in, sort->val_out, (int)count, 0, bits, st);
    MCB_LAUNCHED(1 + (bits + 7) / 8 + 2);
}
__global__ void k_chunk_bounds(const unsigned long long* __restrict__ sorted_key, uint32_t n, const unsigned long long* __restrict__ lo,
                               int n_lo, uint32_t* pos)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_lo) return;
    uint32_t a = 0, b = n;  // first position whose key is >= lo[c]
    while (a < b) { const uint32_t m = (a + b) >> 1; if (sorted_key[m] < lo[c]) a = m + 1; else b = m; }
    pos[c] = a;
}
void chunk_bounds(cudaStream_t st, const SortScratch* sort, uint32_t n, const unsigned long long* lo, int n_lo, uint32_t* pos)
{
    k_chunk_bounds<<<1, 64, 0, st>>>(sort->key_out, n, lo, n_lo, pos);
    MCB_LAUNCHED(1);
}
void source_sorted_range(cudaStream_t st, const DevProblem& P, const Bank& B, uint32_t* active, int32_t first_hist, uint32_t q0,
                         uint32_t count, uint64_t nps0, const SourceBankView& V, Counters* C, const SortScratch* sort)
{
    if (!count) return;
    k_source<<<blocks_for(count, BLOCK), BLOCK, 0, st>>>(P, B, active, first_hist, count, nps0, V, C, sort->key_out, sort->val_out,
                                                         sort->rng_after, sort->rot, q0);
    MCB_LAUNCHED(1);
}
size_t sort_temp_bytes(uint32_t n)
{
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n, 0, 64, (cudaStream_t)0);
    return bytes;
}
void xs_stage(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* active, int cur, uint64_t n_hint, Counters* C)
{
    k_xs_stage<<<grid_for(n_hint), BLOCK, 0, st>>>(P, B, active, cur, C);
    MCB_LAUNCHED(1);
}
void flight(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* active, int cur, uint64_t n_hint,
            uint32_t* evq, Counters* C, const HistoryAcc& H, const TallyAcc& T)
{
    k_flight<<<grid_for(n_hint), BLOCK, 0, st>>>(P, B, active, cur, evq, C, H, T);
    MCB_LAUNCHED(1);
}
void collide(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* evq, int cur, uint64_t n_hint, Counters* C,
             uint32_t* next, const HistoryAcc& H, const TallyAcc& T, SiteReq* reqs,
             uint64_t site_cap, uint32_t n_slots, double k_eff)
{
    k_collide<<<grid_for(n_hint), BLOCK, 0, st>>>(P, B, evq, cur, C, next, H, T, reqs, site_cap, n_slots, k_eff);
    MCB_LAUNCHED(1);
}
void cross(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* evq, int cur, uint64_t n_hint, Counters* C,
           uint32_t* next, const TallyAcc& T, uint32_t n_slots)
{
    k_cross<<<grid_for(n_hint), BLOCK, 0, st>>>(P, B, evq, cur, C, next, T, n_slots);
    MCB_LAUNCHED(1);
}
void walk(cudaStream_t st, const DevProblem& P, const Bank& B, uint64_t begin, uint64_t end, Counters* C, const HistoryAcc& H,
          const TallyAcc& T, SiteReq* reqs, uint64_t site_cap, uint32_t n_slots, double k_eff)
{
    if (end <= begin) return;
    // persistent: every resident warp draws chunks of slots until the pass runs dry
    const uint64_t n = end - begin;
    const unsigned resident = 148u * MCB_STEP_MINB;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(((uint64_t)n + BLOCK - 1) / BLOCK, resident));
    const uint64_t warps = (uint64_t)grid * WARPS;
    const uint32_t chunk = (uint32_t)std::max<uint64_t>(32, std::min<uint64_t>(128, n / (warps * 8)));
    // two instances: the one for cycles that score nothing carries no estimator code (and no energy_old upkeep)
    // and one pair for problems where nothing is born in flight (k-eigenvalue without splitting)
#define MCB_WALK(TALLY, SHARED) k_walk<TALLY, SHARED><<<grid, BLOCK, 0, st>>>(P, B, (unsigned long long)begin, (unsigned long long)end, chunk, C, H, T, reqs, site_cap, n_slots, k_eff)
    if (T.on) { if (P.shared_histories) MCB_WALK(true, true); else MCB_WALK(true, false); }
    else { if (P.shared_histories) MCB_WALK(false, true); else MCB_WALK(false, false); }
#undef MCB_WALK
    MCB_LAUNCHED(1);
}
void finish(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* active, int cur, uint64_t n_hint, Counters* C,
            uint32_t* next, const HistoryAcc& H, const TallyAcc& T, SiteReq* reqs, uint64_t site_cap,
            uint32_t n_slots, double k_eff)
{
    k_finish<<<grid_for(n_hint), BLOCK, 0, st>>>(P, B, active, cur, C, next, H, T, reqs, site_cap, n_slots, k_eff);
    MCB_LAUNCHED(1);
}
void bank_sample_order(cudaStream_t st, const DevProblem& P, const SiteReq* reqs, uint64_t n, const uint32_t* offset, Site* out)
{
    if (n) { k_bank_sample_order<<<blocks_for(n), 256, 0, st>>>(P, reqs, n, offset, out); MCB_LAUNCHED(1); }
}
size_t scan_temp_bytes(uint32_t n)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return bytes;
}
void scan_sites(cudaStream_t st, void* temp, size_t temp_bytes, const int32_t* nsite, uint32_t* offset, uint32_t n)
{
    if (n) { cub::DeviceScan::ExclusiveSum(temp, temp_bytes, nsite, offset, (int)n, st); MCB_LAUNCHED(2); }  // init + scan kernels
}
void reduce_k(cudaStream_t st, const double* kC, const double* kTL, uint32_t n, Counters* C)
{
    if (n) { k_reduce_k<<<min(blocks_for(n), 148u * 8u), 256, 0, st>>>(kC, kTL, n, C); MCB_LAUNCHED(1); }
}
void entropy_history(cudaStream_t st, const DevProblem& P, const Site* bank, const uint32_t* offset,
                     const int32_t* nsite, uint32_t n_hist, Counters* C)
{
    if (n_hist) { k_entropy_history<<<min(blocks_for(n_hist), 148u * 8u), 256, 0, st>>>(P, bank, offset, nsite, n_hist, C); MCB_LAUNCHED(1); }
}
void entropy_histogram(cudaStream_t st, const DevProblem& P, const Site* bank, uint64_t n, unsigned long long* bins)
{
    if (n) { k_entropy_histogram<<<min(blocks_for(n), 148u * 8u), 256, 0, st>>>(P, bank, n, bins); MCB_LAUNCHED(1); }
}
int tally_chunks(uint32_t n_hist) { return (int)((n_hist + TALLY_CHUNK - 1) / TALLY_CHUNK); }
void tally_reduce(cudaStream_t st, double* acc, int64_t stride, uint32_t n_hist, int64_t n_tallies, double* partial,
                  double* sum, double* squared)
{
    if (!n_hist || !n_tallies) return;
    const int nc = tally_chunks(n_hist);
    k_tally_partial<<<dim3(nc, (unsigned)n_tallies), 256, 0, st>>>(acc, stride, n_hist, partial, nc);
    k_tally_final<<<blocks_for(n_tallies, 128), 128, 0, st>>>(partial, nc, n_tallies, sum, squared);
    MCB_LAUNCHED(2);
}
void pack_sites(cudaStream_t st, const double* s8, const int32_t* cells, uint64_t n, Site* out, double* dir_x)
{
    if (n) { k_pack_sites<<<blocks_for(n), 256, 0, st>>>(s8, cells, n, out, dir_x); MCB_LAUNCHED(1); }
}
void unpack_sites(cudaStream_t st, const Site* in, const double* dir_x, uint64_t n, double* s8, int32_t* cells)
{
    if (n) { k_unpack_sites<<<blocks_for(n), 256, 0, st>>>(in, dir_x, n, s8, cells); MCB_LAUNCHED(1); }
}
void gather_sites(cudaStream_t st, const SourceBankView& V, uint64_t n, Site* out)
{
    if (n) { k_gather_sites<<<blocks_for(n), 256, 0, st>>>(V, n, out); MCB_LAUNCHED(1); }
}
void iota(cudaStream_t st, uint32_t* a, uint32_t n)
{
    if (n) { k_iota<<<blocks_for(n), 256, 0, st>>>(a, n); MCB_LAUNCHED(1); }
}

void xs_lookup(cudaStream_t st, const DevProblem& P, int material, const double* E, int64_t n, double* out5)
{
    if (n) { k_xs_lookup<<<blocks_for(n), 256, 0, st>>>(P, material, E, n, out5); MCB_LAUNCHED(1); }
}
void select_channel(cudaStream_t st, const DevProblem& P, int material, int kind, const double* E, const double* xi,
                    int64_t n, int32_t* out)
{
    if (n) { k_select_channel<<<blocks_for(n), 256, 0, st>>>(P, material, kind, E, xi, n, out); MCB_LAUNCHED(1); }
}
void beta(cudaStream_t st, const DevProblem& P, int material, int local_n, const double* E, int64_t n, double* out)
{
    if (n) { k_beta<<<blocks_for(n), 256, 0, st>>>(P, material, local_n, E, n, out); MCB_LAUNCHED(1); }
}
void rng(cudaStream_t st, uint64_t seed0, const uint64_t* nps, int64_t n, int ndraw, uint64_t* out)
{
    if (n) { k_rng<<<blocks_for(n), 256, 0, st>>>(seed0, nps, n, ndraw, out); MCB_LAUNCHED(1); }
}
void geometry(cudaStream_t st, const DevProblem& P, const int32_t* cell, const double* pos, const double* dir,
              int64_t n, double* out3)
{
    if (n) { k_geometry<<<blocks_for(n), 256, 0, st>>>(P, cell, pos, dir, n, out3); MCB_LAUNCHED(1); }
}
void search_cell(cudaStream_t st, const DevProblem& P, const double* pos, int64_t n, int32_t* out)
{
    if (n) { k_search_cell<<<blocks_for(n), 256, 0, st>>>(P, pos, n, out); MCB_LAUNCHED(1); }
}
void scatter(cudaStream_t st, const DevProblem& P, int nuclide, const uint64_t* nps, int64_t n, double* io5)
{
    if (n) { k_scatter<<<blocks_for(n), 256, 0, st>>>(P, nuclide, nps, n, io5); MCB_LAUNCHED(1); }
}
void watt(cudaStream_t st, const DevProblem& P, int nuclide, const uint64_t* nps, const double* E, int64_t n, double* out)
{
    if (n) { k_watt<<<blocks_for(n), 256, 0, st>>>(P, nuclide, nps, E, n, out); MCB_LAUNCHED(1); }
}

}  // namespace mcbk
