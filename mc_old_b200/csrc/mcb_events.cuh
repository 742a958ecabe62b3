// mcb_events.cuh — the events of one particle on registers (lookup, flight, collide, cross) and tally scoring,
// shared by the walk kernel (mcb_walk.cu: block-level event queues in shared memory) and the global event-queue
// kernels (mcb_kernels.cu: the cross-check formulation).
//
// Every event function states the reference lines it follows (file:line under /root/reference).  Nothing here is
// GEMM-shaped: FP64 scalar work and L2-resident table gathers, so tensor cores are unused on purpose.
#ifndef MCB_EVENTS_CUH
#define MCB_EVENTS_CUH

#include "mcb_kernels.h"

namespace mcbe {

constexpr unsigned FULL = 0xffffffffu;
#ifndef MCB_BLOCK
#define MCB_BLOCK 128
#endif
constexpr int BLOCK = MCB_BLOCK;
constexpr int WARPS = BLOCK / 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// per-history accumulators are bumped with reductions at L2 (RED, no return value): a read-modify-write in the
// thread would stall the warp for a DRAM round trip on every event
__device__ __forceinline__ void hist_add(double* p, double v) { atomicAdd(p, v); }

// ---------------------------------------------------------------------------------------------
// per-history tally accumulation (Estimator::score adds into Tally::hist, Estimator.cpp:298-336; end_history turns
// hist into sum += hist, squared += hist^2, Estimator.cpp:339-346).
//  * walk kernel: a history is followed by one lane at a time, so its accumulator is a private open-addressed table
//    {tally index -> value} in global memory (row = the history's context), touched entries listed for the flush;
//  * event-queue kernels: particles of one history are spread over threads: dense rows acc[tally][history of batch].
// ---------------------------------------------------------------------------------------------
// A history's table is only ever touched from one SM (the lane, or the lanes of one block, that follow it), so plain
// cached accesses are coherent and most probes are L1 hits.
__device__ __forceinline__ void tally_add(const TallyAcc& T, Counters* C, int row, int& n_touched, int64_t t, double v)
{
    if (T.acc) { atomicAdd(T.acc + t * T.stride + (int64_t)(row - T.first_hist), v); return; }
    const uint32_t mask = T.tab_mask;
    double* vals = T.tab_val + (size_t)row * (mask + 1u);
    if (T.direct) {
        // every tally has its own entry (t is the table position, see tally_slot): the value itself says whether the
        // entry is in use, so an add is one read-modify-write of one sector and a zero score touches nothing
        if (v == 0.0) return;
#ifdef MCB_TALLY_CG
        const double old = __ldcg(vals + t);
        if (old == 0.0 && n_touched <= (int)mask) { __stcg(T.tab_list + (size_t)row * (mask + 1u) + n_touched, (uint16_t)t); n_touched++; }
        __stcg(vals + t, old + v);
#else
        const double old = vals[t];
        if (old == 0.0 && n_touched <= (int)mask) { T.tab_list[(size_t)row * (mask + 1u) + n_touched] = (uint16_t)t; n_touched++; }
        vals[t] = old + v;
#endif
        return;
    }
    uint32_t* keys = T.tab_key + (size_t)row * (mask + 1u);
    uint32_t h = (uint32_t)t & mask;
    for (uint32_t probe = 0; probe <= mask; probe++, h = (h + 1u) & mask) {
        const uint32_t k = keys[h];
        if (k == (uint32_t)t + 1u) { vals[h] += v; return; }
        if (k == 0u) {
            keys[h] = (uint32_t)t + 1u;
            vals[h] = v;
            T.tab_list[(size_t)row * (mask + 1u) + n_touched] = (uint16_t)h;
            n_touched++;
            return;
        }
    }
    C->overflow_tally = 1;  // a history touched more bins than a table holds
}
// Position of tally (estimator E, flat filter bin idx, score k) in a history's table.  Tally::sum is laid out
// [score][bin] (Estimator.cpp:288-295), but one event adds to all scores of ONE bin: direct tables keep an estimator's
// scores of a bin side by side ([bin][score]) so that those adds fall into one or two sectors instead of one each.
__device__ __forceinline__ int64_t tally_slot(const TallyAcc& T, const mcb_estimator& E, int64_t idx, int k, int64_t bins)
{
    return T.direct ? E.tally_begin + idx * E.n_scores + k : E.tally_begin + idx + (int64_t)k * bins;
}
// ... and back: the index into Tally::sum of the table position `pos` of a direct table
__device__ __forceinline__ uint32_t tally_of_slot(const DevProblem& P, uint32_t pos)
{
    int e = 0;
    while (e + 1 < P.n_estimators && (int64_t)pos >= P.estimators[e + 1].tally_begin) e++;
    const int64_t begin = P.estimators[e].tally_begin, S = P.estimators[e].n_scores;
    const int64_t local = (int64_t)pos - begin, bins = P.estimators[e].n_tallies / S;
    return (uint32_t)(begin + (local % S) * bins + local / S);
}
// takes entry i of a history's touched list out of its table: {index into Tally::sum, value}
__device__ __forceinline__ bool tally_take(const DevProblem& P, const TallyAcc& T, int row, int i, uint32_t& t, double& v)
{
    const uint32_t size = T.tab_mask + 1u;
    const uint32_t h = T.tab_list[(size_t)row * size + i];
    double* val = T.tab_val + (size_t)row * size + h;
    if (T.direct) {
        v = __longlong_as_double((long long)atomicExch((unsigned long long*)val, 0ull));  // a position listed twice is taken once
        if (v == 0.0) return false;
        t = tally_of_slot(P, h);
        return true;
    }
    uint32_t* key = T.tab_key + (size_t)row * size + h;
    t = *key - 1u;
    v = *val;
    *key = 0u;
    return true;
}

// ---------------------------------------------------------------------------------------------
// tally scoring (Estimator::score, Estimator.cpp:298-336, for filters that yield one bin: surface, cell, energy)
// ---------------------------------------------------------------------------------------------
struct ScoreState {   // the particle as Score / Filter / the simulating estimators see it
    double w, E, E_old, speed, t, t_old;
    double u, v, wd;  // direction
    int cell, surface_old, material, uidx;
    int tdmc;         // index of the next census time (time-dependent mode)
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
    int i = c.k;  // 0 = the piece in loc1 (when it exists), then the full bins, then the piece in loc2
    if (c.first) {
        if (i == 0) { c.idx = c.loc1; c.l = (g[c.loc1 + 1] - s.t_old) * s.speed; return; }
        i--;
    }
    const int num_bin = c.loc2 - c.loc1 - 1;
    if (i < num_bin) { c.idx = c.loc1 + i + 1; c.l = (g[c.loc1 + i + 2] - g[c.loc1 + i + 1]) * s.speed; return; }
    c.idx = c.loc2; c.l = (s.t - g[c.loc2]) * s.speed;
}
// the general form: a time filter splits the track into pieces
__device__ __noinline__ static int estimator_score_pieces(const DevProblem& P, const TallyAcc& T, Counters* C, const mcb_estimator& E,
                                                          const ScoreState& s, double l_in, int row, int n_touched, ChannelCache& CC)
{
    constexpr int MAXF = 4;
    const int nf = E.n_filters < MAXF ? E.n_filters : MAXF;
    FilterCursor cur[MAXF];
    int64_t factor[MAXF + 1];  // idx_factor (Estimator.cpp:288-295)
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
            if (c.idx < 0 || c.idx >= F.grid_n - 1) return n_touched;
            break;
        case MCB_FILTER_TDMC:
            if (s.t != g[s.tdmc]) return n_touched;
            c.idx = s.tdmc;
            break;
        default: {                                                                                           // time :199-246
            const int Nbin = F.grid_n - 1;
            c.loc1 = mcb_binary_search(s.t_old, g, F.grid_n);
            c.loc2 = mcb_binary_search(s.t, g, F.grid_n);
            if (c.loc1 == c.loc2) {
                if (c.loc1 < 0 || c.loc1 >= Nbin) return n_touched;
                c.idx = c.loc1;
            } else {
                c.first = c.loc1 >= 0;
                c.n = (c.first ? 1 : 0) + (c.loc2 - c.loc1 - 1) + (c.loc2 < Nbin ? 1 : 0);
                if (c.n == 0) return n_touched;
                time_piece(F, g, s, c);
            }
        }
        }
    }
    for (;;) {
        double l = MCB_MAX_FLOAT;
        int64_t idx_1D = 0;
        for (int i = 0; i < nf; i++) { l = fmin(l, cur[i].l); idx_1D += (int64_t)cur[i].idx * factor[i + 1]; }
        for (int k = 0; k < E.n_scores; k++) {
            const double v = score_value(P, P.scores[E.score_begin + k], s, l, CC);
            if (idx_1D >= 0 && idx_1D < factor[0]) tally_add(T, C, row, n_touched, tally_slot(T, E, idx_1D, k, factor[0]), v);
            else {  // a bin index past its filter lands in a neighbouring score's bins, as long as it stays inside e_tally
                const int64_t local = idx_1D + (int64_t)k * factor[0];
                if (local >= 0 && local < E.n_tallies)
                    tally_add(T, C, row, n_touched, tally_slot(T, E, local % factor[0], (int)(local / factor[0]), factor[0]), v);
            }
        }
        for (int i = 0; i < nf; i++) {
            FilterCursor& c = cur[i];
            c.l -= l;
            if (c.l < MCB_EPSILON_FLOAT) {
                if (c.k == c.n - 1) return n_touched;
                c.k++;
                const mcb_filter F = P.filters[E.filter_begin + i];
                time_piece(F, P.filter_grid + F.grid_begin, s, c);
            }
        }
        if (nf == 0) return n_touched;
    }
    return n_touched;
}
__device__ __forceinline__ void estimator_score_plain(const DevProblem& P, const TallyAcc& T, Counters* C, const mcb_estimator& E,
                                                      const ScoreState& s, double l_in, int row, int& n_touched, ChannelCache& CC)
{
    constexpr int MAXF = 4;
    const int nf = E.n_filters < MAXF ? E.n_filters : MAXF;
    bool pieces = false;
    for (int i = 0; i < nf; i++) pieces |= P.filters[E.filter_begin + i].type == MCB_FILTER_TIME;
    if (!pieces) {
        // no time filter: every filter yields one bin and the whole length, one pass of the loop below
        int64_t idx_1D = 0, bins = 1;
        for (int i = 0; i < nf; i++) {
            const mcb_filter F = P.filters[E.filter_begin + i];
            int idx;
            switch (F.type) {
            case MCB_FILTER_SURFACE: idx = mcb_binary_search((double)s.surface_old, P.filter_grid + F.grid_begin, F.grid_n) + 1; break;
            case MCB_FILTER_CELL: idx = mcb_binary_search((double)s.cell, P.filter_grid + F.grid_begin, F.grid_n) + 1; break;
            case MCB_FILTER_TDMC:  // only a particle that sits exactly on its census time scores (Estimator.cpp:247-263)
                if (s.t != P.filter_grid[F.grid_begin + s.tdmc]) return;
                idx = s.tdmc;
                break;
            default:
                idx = filter_bin(P, F.grid_begin, F.grid_n, F.type == MCB_FILTER_ENERGY ? s.E : s.E_old, CC);
                if (idx < 0 || idx >= F.grid_n - 1) return;
            }
            idx_1D = idx_1D * F.size + idx;
            bins *= F.size;
        }
        const double l = fmin(MCB_MAX_FLOAT, l_in);
        for (int k = 0; k < E.n_scores; k++) {
            const double v = score_value(P, P.scores[E.score_begin + k], s, l, CC);
            if (idx_1D >= 0 && idx_1D < bins) tally_add(T, C, row, n_touched, tally_slot(T, E, idx_1D, k, bins), v);
            else {
                const int64_t local = idx_1D + (int64_t)k * bins;
                if (local >= 0 && local < E.n_tallies) tally_add(T, C, row, n_touched, tally_slot(T, E, local % bins, (int)(local / bins), bins), v);
            }
        }
        return;
    }
    n_touched = estimator_score_pieces(P, T, C, E, s, l_in, row, n_touched, CC);
}
// One estimator scores one event.  The TRMM estimators first let a COPY of the particle scatter / fission, drawing
// from the particle's own stream like the reference draws from its global one (EstimatorScatter / FissionPrompt /
// FissionDelayed::score, Estimator.cpp:441-482): energy_old = incident energy, energy and speed = outgoing.
__device__ __forceinline__ void estimator_score(const DevProblem& P, const TallyAcc& T, Counters* C, int e, const ScoreState& s, uint64_t& rng,
                                                double l, int row, int& n_touched, ChannelCache& CC)
{
    const mcb_estimator E = P.estimators[e];
    if (E.simulate == MCB_SIM_NONE) { estimator_score_plain(P, T, C, E, s, l, row, n_touched, CC); return; }
    if (s.material < 0) return;
    const DevMaterial& M = P.materials[s.material];
    ScoreState q = s;
    int n = -1;
    if (E.simulate == MCB_SIM_SCATTER) {
        (void)macro_channel(P, s.material, s.E, 0, false, mcb_urand(rng), &n, CC);
        if (n < 0) return;  // the reference dereferences a null nuclide here
        scatter_sample(P.nuclides[n], q.u, q.v, q.wd, q.E, q.speed, rng);
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
    if (needs_X) { const UnionPos up = union_pos(M, q.E); q.uidx = up.u; macro_xs(P, M, up, q.E, q.X); }
    estimator_score_plain(P, T, C, E, q, l, row, n_touched, CC);
}
__device__ __forceinline__ bool has_attached(const DevProblem& P, int kind, int id)
{
    return P.attach_begin[kind][id + 1] > P.attach_begin[kind][id];
}

// ---------------------------------------------------------------------------------------------
// the particle on registers
// ---------------------------------------------------------------------------------------------
struct Particle {
    double x, y, z, u, v, w, E, speed, wgt, t;
    double Eold, told;  // Particle::energy_old / time_old (Particle.cpp:42-76); kept up by the scoring instances only
    uint64_t rng;
    int cell, hist;
    int row;        // tally accumulator row of the history: context (walk kernel) or shard-local history index (dense rows)
    int n_touched;  // entries of the history's tally table in use (walk kernel)
    int drow;       // dense tally row of a history that is shared between lanes, -1 while a history is followed by one lane
    int tdmc;       // time-dependent mode: index of the next census time (Particle::tdmc, Particle.cpp:29,82)
};
// all estimators attached to surface / cell `id` score one event of particle p.  Cold and out of line, with every
// input BY VALUE (no address of a register-resident particle escapes), so that the transport kernels' register
// allocation is not shaped by it; returns the particle's stream state (the simulating estimators draw from it) and,
// through the pointer, the table fill.  have_X = false (surface estimators): the cross sections of the cell the
// particle is now in are looked up here.
struct ScoreRet { uint64_t rng; int n_touched; };
static __device__ __noinline__ ScoreRet score_event(const DevProblem& P, const TallyAcc& T, Counters* C, int kind, int id, double w, double E, double E_old,
                                             double speed, double t, double t_old, double du, double dv, double dw, int cell, int row,
                                             int n_touched, uint64_t rng, int material, int uidx, bool have_X, double Xt, double Xs, double Xc,
                                             double Xf, double Xnf, int surface_old, double l, int tdmc)
{
    ScoreState s;
    s.tdmc = tdmc;
    s.w = w; s.E = E; s.speed = speed; s.u = du; s.v = dv; s.wd = dw;
    s.E_old = E_old;
    s.t = t; s.t_old = t_old;
    s.cell = cell; s.surface_old = surface_old; s.material = material; s.uidx = uidx;
    s.X = MacroXS{Xt, Xs, Xc, Xf, Xnf};
    if (!have_X) {
        s.uidx = -1;
        if (material >= 0) { const UnionPos up = union_pos(P.materials[material], E); s.uidx = up.u; macro_xs(P, P.materials[material], up, E, s.X); }
    }
    const int b = P.attach_begin[kind][id], e = P.attach_begin[kind][id + 1];
    ChannelCache CC;  // microscopic data at the (at most two) energies the estimators of this event ask about
    channel_cache_reset(CC);
    for (int i = b; i < e; i++) estimator_score(P, T, C, P.attach_list[kind][i], s, rng, l, row, n_touched, CC);
    return ScoreRet{rng, n_touched};
}
#define MCB_SCORE_EVENT(kind, id, p, material, uidx, have_X, X, surface_old, l)                                                              \
    do {                                                                                                                                   \
        const ScoreRet sr_ = score_event(P, T, C, kind, id, (p).wgt, (p).E, P.track_old ? (p).Eold : (p).E, (p).speed, (p).t,              \
                                         P.track_time ? (p).told : (p).t, (p).u, (p).v, (p).w, (p).cell, (p).row, (p).n_touched, (p).rng, \
                                         material, uidx, have_X, (X).t, (X).s, (X).c, (X).f, (X).nf, surface_old, l, P.n_tdmc ? (p).tdmc : 0); \
        (p).rng = sr_.rng; (p).n_touched = sr_.n_touched;                                                                                  \
    } while (0)

// where same-history secondaries go (fixed-source fission neutrons, split copies): bank slots behind the running
// batch (event-queue kernels) or the history's own LIFO stack (walk kernel); NoSink where none can be born
struct NoSink {
    __device__ __forceinline__ void push(const Particle&) {}
};

// xs_lookup event
template <class DET>
__device__ __forceinline__ bool ev_lookup(const DevProblem& P, const Particle& p, MacroXS& X, int& uidx, DET& D)
{
    const int m = P.cells[p.cell].material;
    if (m < 0) return false;
    const DevMaterial M = P.materials[m];
    const UnionPos up = union_pos(M, p.E);
    uidx = up.u;
    macro_xs_impl(P, M, up, p.E, X, D);
    return true;
}

// flight event: surface_intersect + collision_distance + move_particle (general.cpp:40-83,177-207).
// A history that is followed by one lane at a time keeps its EstimatorK scores and its site count on registers
// (HistLocal) and stores them once when it ends; otherwise they are bumped in memory with reductions.
struct HistLocal { double kC, kTL; int nsite; };

// Returns the event the flight ends in: 0 collision, 1 surface, 2 census (time-dependent mode, TD instances only: the
// particle stops at its next census time, general.cpp:187-195, and Simulator::time_hit moves it on to the next interval
// or, after the last census, kills it, time_dependent.cpp:51-55).
template <bool TALLY, bool TD = false>
__device__ __forceinline__ int ev_flight(const DevProblem& P, Particle& p, const MacroXS& X, int uidx, const HistoryAcc& H,
                                         const TallyAcc& T, Counters* C, int& S_hit, HistLocal* L = nullptr)
{
    const int m = P.cells[p.cell].material;
    double dsurf;
    S_hit = surface_intersect(P, p.cell, p.x, p.y, p.z, p.u, p.v, p.w, dsurf);
    double dcol;
    if (m >= 0) dcol = -mcb_log(mcb_urand(p.rng)) / X.t;   // exponential_sample (Algorithm.cpp:123-126)
    else dcol = MCB_MAX_FLOAT_LESS;                      // vacuum (general.cpp:44-46)
    int event = dcol > dsurf ? 1 : 0;
    double l = event ? dsurf : dcol;
    if (TD && P.n_tdmc) {
        const double dbound = (P.tdmc_time[p.tdmc] - p.t) * p.speed;
        if (l > dbound) { l = dbound; event = 2; }
    }
    // Particle::move (Particle.cpp:66-76)
    p.x += p.u * l; p.y += p.v * l; p.z += p.w * l;
    if (TALLY) p.told = p.t;
    p.t += l / p.speed;
    if (P.ksearch && m >= 0) {  // estimate_TL (Estimator.cpp:509-512)
        if (L) L->kTL += X.nf * p.wgt * l;
        else hist_add(&H.kTL[p.hist], X.nf * p.wgt * l);
    }
    if (TALLY && T.on && has_attached(P, MCB_ATTACH_CELL_TL, p.cell)) {
        MCB_SCORE_EVENT(MCB_ATTACH_CELL_TL, p.cell, p, m, uidx, true, X, -1, l);
    }
    if (TD && event == 2) {
        p.tdmc++;
        if (p.tdmc == P.n_tdmc) p.wgt = 0.0;  // Particle::kill
    }
    return event;
}

// collide event, first half: Simulator::collision up to the fission dispatch (general.cpp:121-150): collision
// tallies, bank_nu, fissioning nuclide, prompt/delayed draw.  Tells how many fission sites (k-eigenvalue) or
// same-history secondaries (fixed source) the second half will write.
struct CollideCtx {
    int m, N_fission;
    unsigned n_sites, n_second;
    unsigned n_forced;  // time-dependent mode, delayed fission: fission neutrons, each forced to decay once per remaining interval
    double rXt;  // refined reciprocal of SigmaT: it divides three times in a collision (mcb_div_shared)
};
template <bool TALLY, class DET>
__device__ __forceinline__ bool ev_collide_pre(const DevProblem& P, Particle& p, const MacroXS& X, int uidx, const DET& D,
                                               const TallyAcc& T, Counters* C, double k_eff, CollideCtx& c)
{
    c.m = P.cells[p.cell].material;
    c.N_fission = -1; c.n_sites = 0; c.n_second = 0; c.n_forced = 0;
    if (c.m < 0) { p.wgt = 0.0; return false; }  // vacuum: kill (general.cpp:124-128)
    if (TALLY && T.on && has_attached(P, MCB_ATTACH_CELL_C, p.cell)) {
        MCB_SCORE_EVENT(MCB_ATTACH_CELL_C, p.cell, p, c.m, uidx, true, X, -1, 0.0);
    }
    // floor( w/k * nuSigmaF / SigmaT + xi ) (general.cpp:135-136)
    c.rXt = mcb_rcp_shared(X.t);
    const double a = mcb_div_shared(p.wgt / k_eff * X.nf, X.t, c.rXt);
    const double bn = floor(a + mcb_urand(p.rng));
    const unsigned bank_nu = bn > 0.0 ? (unsigned)bn : 0u;
    const DevMaterial& M = P.materials[c.m];
    int ln = 0;
    const double xi_f = mcb_urand(p.rng);
    c.N_fission = DET::present ? select_from_detail<1>(P, M, D, X.nf, xi_f, &ln)
                               : select_nuclide(P, M, uidx, p.E, 1, X.nf, xi_f, &ln);  // Material.cpp:116-125
    if (c.N_fission >= 0) {
        // prompt or delayed (ksearch.cpp:24-38, fixed_source.cpp:12,41-52)
        const double beta = DET::present ? D.beta(ln) : micro_col(P.nuclides[c.N_fission], nuclide_index(M, uidx, ln), p.E, 1);
        const bool prompt = mcb_urand(p.rng) > beta;
        if (P.ksearch) {
            if (!prompt) (void)mcb_urand(p.rng);  // precursor group pick, result unused
            c.n_sites = bank_nu;
        } else if (prompt) {
            c.n_second = bank_nu;
        } else if (P.n_tdmc) {
            // delayed, time-dependent mode (fixed_source.cpp:25-40): room for one neutron per remaining interval
            c.n_forced = bank_nu;
            c.n_second = bank_nu * (unsigned)(P.n_tdmc - p.tdmc);
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
// from the request's own stream.  Fixed source (fixed_source.cpp:12-22): same-history secondaries go to the sink,
// each sampled from, and continuing on, its own stream.  The parent's stream does not advance.  Only lanes that
// bank anything call this; callers reconverge the warp afterwards.
template <class SINK>
__device__ __forceinline__ void ev_collide_bank(const DevProblem& P, const Particle& p, const CollideCtx& c, const HistoryAcc& H,
                                                Counters* C, SiteReq* reqs, uint64_t site_cap, unsigned long long site0, SINK& sink,
                                                HistLocal* L, double E_in, uint64_t seed_in);
template <class SINK>
__device__ __forceinline__ void ev_collide_bank(const DevProblem& P, const Particle& p, const CollideCtx& c, const HistoryAcc& H,
                                                Counters* C, SiteReq* reqs, uint64_t site_cap, unsigned long long site0, SINK& sink,
                                                HistLocal* L = nullptr)
{
    ev_collide_bank(P, p, c, H, C, reqs, site_cap, site0, sink, L, p.E, p.rng);
}
// E_in / seed_in: the particle's energy and stream state at the collision, for callers that bank after the scatter
// kinematics have moved both on (the walk kernel's lean instances: the slot reservation's round trip is hidden that way)
template <class SINK>
__device__ __forceinline__ void ev_collide_bank(const DevProblem& P, const Particle& p, const CollideCtx& c, const HistoryAcc& H,
                                                Counters* C, SiteReq* reqs, uint64_t site_cap, unsigned long long site0, SINK& sink,
                                                HistLocal* L, double E_in, uint64_t seed_in)
{
    uint64_t seed = seed_in;
    if (c.n_sites) {
        int seq0;
        if (L) { seq0 = L->nsite; L->nsite += (int)c.n_sites; }
        else seq0 = atomicAdd(&H.nsite[p.hist], (int)c.n_sites);
        for (unsigned b = 0; b < c.n_sites; b++) {
            seed = (seed * MCB_RN_JUMP40) & MCB_RN_MASK;
            SiteReq r;
            r.x = p.x; r.y = p.y; r.z = p.z; r.t = p.t; r.E_in = E_in; r.seed = seed;
            r.cell = p.cell; r.seq = seq0 + (int)b; r.hist = p.hist; r.nuclide = c.N_fission;
            if (site0 + b < site_cap) reqs[site0 + b] = r;
            else C->overflow_sites = 1;
        }
    }
    if (c.n_forced) {
        // Simulator::forced_decay (time_dependent.cpp:16-45) for the rest of the current interval, then for every later
        // one; each neutron is drawn from its own stream, its roulette (fixed_source.cpp:31,36) included
        const DevNuclide& N = P.nuclides[c.N_fission];
        for (unsigned i = 0; i < c.n_forced; i++) {
            for (int j = p.tdmc - 1; j < P.n_tdmc - 1; j++) {
                const bool rest = j < p.tdmc;
                const double initial = rest ? p.t : P.tdmc_time[j];
                const double interval = rest ? P.tdmc_time[p.tdmc] - p.t : P.tdmc_interval[j + 1];
                seed = (seed * MCB_RN_JUMP40) & MCB_RN_MASK;
                Particle q = p;
                q.rng = seed;
                q.tdmc = rest ? p.tdmc : j + 1;
                q.t = initial + mcb_urand(q.rng) * interval;
                double prob[6], w = 0.0;
#pragma unroll
                for (int k = 0; k < 6; k++) {
                    prob[k] = N.fraction[k] * N.lambda[k] * exp(-N.lambda[k] * (q.t - p.t));  // f_lambda = fraction * lambda
                    w += prob[k];
                }
                const double xi = mcb_urand(q.rng) * w;  // std::accumulate adds in the same order
                int cg = 0;
                double sum = 0.0;
                bool found = false;
#pragma unroll
                for (int k = 0; k < 6; k++) { sum += prob[k]; if (!found && sum > xi) { cg = k; found = true; } }
                q.E = chid_sample(N, cg, q.rng);
                isotropic_direction(q.rng, q.u, q.v, q.w);
                q.speed = mcb_speed_of_energy(q.E); q.wgt = w * interval; q.Eold = q.E; q.told = q.t;
                if (q.wgt < P.wr) {  // weight_roulette (population_control.cpp:9-15)
                    if (mcb_urand(q.rng) < q.wgt / P.ws) q.wgt = P.ws;
                    else continue;
                }
                sink.push(q);
            }
        }
    } else if (c.n_second) {
        const DevNuclide& N = P.nuclides[c.N_fission];
        for (unsigned b = 0; b < c.n_second; b++) {
            seed = (seed * MCB_RN_JUMP40) & MCB_RN_MASK;
            Particle q = p;
            q.rng = seed;
            q.E = watt_sample(N.watt_a, N.watt_b, N.watt_g, p.E, q.rng);   // energy first, then direction (App. D-4)
            isotropic_direction(q.rng, q.u, q.v, q.w);
            q.speed = mcb_speed_of_energy(q.E); q.wgt = 1.0; q.Eold = q.E;
            sink.push(q);
        }
    }
}
// collide event, last part: k_C, implicit capture, scatter, weight_roulette (general.cpp:146-163,
// population_control.cpp:9-15).  Returns whether the particle survives.
template <bool TALLY, class DET>
__device__ __forceinline__ bool ev_collide_scatter(const DevProblem& P, Particle& p, const MacroXS& X, int uidx, const DET& D,
                                                   const CollideCtx& c, const HistoryAcc& H, HistLocal* L = nullptr)
{
    const double rXt = c.rXt;
    if (P.ksearch && c.N_fission >= 0) {  // estimate_C (Estimator.cpp:503-507)
        const double kc = mcb_div_shared(X.nf * p.wgt, X.t, rXt);
        if (L) L->kC += kc;
        else hist_add(&H.kC[p.hist], kc);
    }
    // implicit absorption (general.cpp:154-156)
    const double implicit = X.c + X.f;
    p.wgt = mcb_div_shared(p.wgt * (X.t - implicit), X.t, rXt);
    const double xi_s = mcb_urand(p.rng);
    int ln_s = 0;
    const int N_scatter = DET::present ? select_from_detail<0>(P, P.materials[c.m], D, X.s, xi_s, &ln_s)
                                       : select_nuclide(P, P.materials[c.m], uidx, p.E, 0, X.s, xi_s, &ln_s);  // Material.cpp:106-115
    if (N_scatter >= 0) {
        if (TALLY) p.Eold = p.E;  // Particle::set_speed keeps the pre-collision energy (Particle.cpp:49-56)
        scatter_sample(P.nuclides[N_scatter], p.u, p.v, p.w, p.E, p.speed, p.rng);
    }
    // weight_roulette (population_control.cpp:9-15)
    if (p.wgt < P.wr) {
        if (mcb_urand(p.rng) < mcb_div_zero_ok(p.wgt, P.ws)) p.wgt = P.ws;  // a pure absorber leaves weight 0
        else { p.wgt = 0.0; return false; }
    }
    return true;
}

// cross event, first half: surface_hit + cell_importance up to the split (general.cpp:89-115,
// population_control.cpp:21-43).  n_copy = split copies the second half will write.
// LEAN (walk kernel, instances that neither score nor keep secondaries): a particle that dies at the crossing ends its
// history, and nothing reads its stream again - the draws the reference still makes for it (importance roulette against
// a ratio of 0, weight roulette of a particle already killed) are skipped, with the same generation bit for bit
template <bool TALLY, bool LEAN = false>
__device__ __forceinline__ bool ev_cross_pre(const DevProblem& P, Particle& p, int S, const TallyAcc& T, Counters* C, unsigned& n_copy)
{
    n_copy = 0;
    if (S < 0) { p.wgt = 0.0; return false; }  // no surface ahead: cannot happen in a closed geometry
    const mcb_surface& Sf = P.surfaces[S];
    const int cell_old = p.cell;
    bool alive = true;
    if (Sf.bc == MCB_BC_TRANSMISSION) {
        p.x += p.u * MCB_EPSILON_FLOAT; p.y += p.v * MCB_EPSILON_FLOAT; p.z += p.w * MCB_EPSILON_FLOAT;
        if (TALLY) p.told = p.t;
        if (!LEAN) p.t += MCB_EPSILON_FLOAT / p.speed;   // (lean instances: below, for the particles that live on)
        int cn = -1;
#ifndef MCB_NO_CROSS_NEIGHBOR
        {   // the cell behind the surface where it is known beforehand (cross_neighbor, mcb_api.cu), else the search
            const int2 nb = __ldg(reinterpret_cast<const int2*>(P.cross_neighbor) + S);
            if ((nb.x & nb.y) != -1) {   // (both -1: nothing known about this surface)
                const double e = mcb_surf_eval(Sf, p.x, p.y, p.z);
                if (e > 0.0) cn = nb.y; else if (e < 0.0) cn = nb.x;
            }
        }
#endif
        if (cn < 0) cn = mcb_search_cell(P.cells, P.n_cells, P.surfaces, P.cell_surface, P.cell_sense, p.x, p.y, p.z);
        if (cn < 0) {  // "[WARNING] A particle is lost" (general.cpp:31-33)
            if (atomicExch(&C->lost, 1) == 0) { C->lost_pos[0] = p.x; C->lost_pos[1] = p.y; C->lost_pos[2] = p.z; }
            alive = false; p.wgt = 0.0;
        } else p.cell = cn;
    } else if (Sf.bc == MCB_BC_VACUUM) {
        alive = false; p.wgt = 0.0;
    } else {
        mcb_surf_reflect(Sf, p.u, p.v, p.w);
        p.x += p.u * MCB_EPSILON_FLOAT; p.y += p.v * MCB_EPSILON_FLOAT; p.z += p.w * MCB_EPSILON_FLOAT;
        if (TALLY) p.told = p.t;
        p.t += MCB_EPSILON_FLOAT / p.speed;
    }
    if (TALLY && T.on && has_attached(P, MCB_ATTACH_SURFACE, S)) {
        const MacroXS X0 = {0, 0, 0, 0, 0};
        MCB_SCORE_EVENT(MCB_ATTACH_SURFACE, S, p, P.cells[p.cell].material, -1, false, X0, S, 0.0);
    }
    if (LEAN && !alive) return false;
    const double Iold = P.cells[cell_old].importance, Inew = P.cells[p.cell].importance;
    if (LEAN && Inew == 0.0 && Iold > 0.0) { p.wgt = 0.0; return false; }  // ratio 0: killed whatever the draw
    if (LEAN && Sf.bc == MCB_BC_TRANSMISSION) p.t += MCB_EPSILON_FLOAT / p.speed;  // the time of a dead particle is never read
    if (Inew != Iold) {
        // a leaking particle enters importance 0: 0 / Iold is decided without dividing (a zero numerator sends the
        // IEEE division routine down its slow path); same value, same draw
        const double rat = mcb_div_zero_ok(Inew, Iold);
        if (rat < 1.0) {
            // the compiler evaluates this division whether or not the branch is taken; with rat = 0 (every leaking
            // particle) a division by zero would go down the division routine's slow path, so it gets a harmless divisor
            const double den = rat > 0.0 ? rat : 1.0;
            if (mcb_urand(p.rng) < rat) p.wgt = p.wgt / den;
            else { alive = false; p.wgt = 0.0; }
        } else {
            const int ns = (int)floor(rat + mcb_urand(p.rng));
            p.wgt = p.wgt / (ns > 0 ? (double)ns : 1.0);  // ns >= 1 here; the guard is for the speculated evaluation (see above)
            n_copy = ns > 1 ? (unsigned)(ns - 1) : 0u;
        }
    }
    return alive;
}
// cross event, second half: the split copies (population_control.cpp:44-48) and weight_roulette, which also
// draws for a particle that was just killed (w = 0 < wr), like the reference
template <class SINK>
__device__ __forceinline__ bool ev_cross_post(const DevProblem& P, Particle& p, bool alive, unsigned n_copy, SINK& sink)
{
    for (unsigned b = 0; b < n_copy; b++) {
        Particle q = p;
        q.rng = mcb_rn_child_seed(p.rng, b);
        sink.push(q);
    }
    if (p.wgt < P.wr) {
        if (mcb_urand(p.rng) < mcb_div_zero_ok(p.wgt, P.ws)) p.wgt = P.ws;  // dead particle: weight 0
        else { p.wgt = 0.0; alive = false; }
    }
    return alive;
}

// Estimator::end_history for one history (Estimator.cpp:339-346): sum += hist, squared += hist^2 per touched bin.
// The whole warp works on the table of the lane whose history ended (histories end on one or two lanes at a time).
__device__ __forceinline__ void flush_history_tallies(const DevProblem& P, const TallyAcc& T, int row, int n_touched, double* s_sum, double* s_sq,
                                                      unsigned lane)
{
    for (int i = (int)lane; i < n_touched; i += 32) {
        uint32_t t;
        double v;
        if (!tally_take(P, T, row, i, t, v)) continue;
        if (s_sum) { atomicAdd(s_sum + t, v); atomicAdd(s_sq + t, v * v); }
        else { atomicAdd(T.sum + t, v); atomicAdd(T.squared + t, v * v); }
    }
}

}  // namespace mcbe
#endif
