// mcb_kernels.cu — sm_100a kernels of the event-based particle-history transport loop.
//
// One generation (reference: one pass of the cycle body of Simulator::start(), handler.cpp:14-44) is
//   source  -> { xs_lookup -> flight -> collide | cross } until the bank is empty -> close-out kernels.
// Particles live in an SoA bank (Bank); each stage runs over an index queue of the particles whose next event it
// is; queues are rebuilt every iteration with warp-ballot stream compaction.
//
// Nothing here is GEMM-shaped: the loop is FP64 scalar work, L2-resident table gathers and HBM streams of the
// bank, so tensor cores are unused on purpose.
#include "mcb_kernels.h"

#include <cub/device/device_scan.cuh>

namespace {

constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// warp-ballot stream compaction: position of this lane's element in the queue behind *cursor (one atomic per warp).
// Must be reached by all 32 lanes of the warp.
__device__ __forceinline__ unsigned warp_append(unsigned int* cursor, bool pred)
{
    const unsigned mask = __ballot_sync(FULL, pred);
    if (mask == 0) return 0;
    const int leader = __ffs(mask) - 1;
    unsigned base = 0;
    if (lane_id() == leader) base = atomicAdd(cursor, __popc(mask));
    base = __shfl_sync(FULL, base, leader);
    return base + __popc(mask & ((1u << lane_id()) - 1));
}
// same with a per-lane element count (warp prefix sum); returns the first position of this lane's elements
template <typename T>
__device__ __forceinline__ T warp_reserve(T* cursor, unsigned n_mine)
{
    unsigned incl = n_mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned t = __shfl_up_sync(FULL, incl, d);
        if (lane_id() >= d) incl += t;
    }
    const unsigned total = __shfl_sync(FULL, incl, 31);
    T base = 0;
    if (total == 0) return 0;
    if (lane_id() == 31) base = atomicAdd(cursor, (T)total);
    base = __shfl_sync(FULL, base, 31);
    return base + (T)(incl - n_mine);
}

__device__ __forceinline__ void hist_add(double* p, double v, int shared)
{
    if (shared) atomicAdd(p, v); else *p += v;
}

// block-wide sum of a counter into a global 64-bit counter (one atomic per warp)
__device__ __forceinline__ void count_add(unsigned long long* dst, bool pred)
{
    const unsigned mask = __ballot_sync(FULL, pred);
    if (mask && lane_id() == 0) atomicAdd(dst, (unsigned long long)__popc(mask));
}

// ---------------------------------------------------------------------------------------------
// tally scoring (Estimator::score, Estimator.cpp:298-336, for filters that yield one bin: surface, cell, energy)
// ---------------------------------------------------------------------------------------------
struct ScoreState {
    double w, E, speed;
    int cell, surface_old, material, u;
    MacroXS X;  // macroscopic xs of `material` at E
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
__device__ __forceinline__ double score_value(const DevProblem& P, const mcb_score& S, const ScoreState& s, double l)
{  // Estimator.cpp:48-124
    const double kv = kernel_value(S.kernel, s, l);
    if (S.score == MCB_SCORE_FLUX) return kv;
    if (S.score == MCB_SCORE_INVERSE_VELOCITY) return kv / s.speed;
    if (s.material < 0) return 0.0;
    switch (S.score) {
    case MCB_SCORE_ABSORPTION: return macro_sigma_a(P, P.materials[s.material], s.u, s.E) * kv;
    case MCB_SCORE_SCATTER: return s.X.s * kv;
    case MCB_SCORE_CAPTURE: return s.X.c * kv;
    case MCB_SCORE_FISSION: return s.X.f * kv;
    case MCB_SCORE_NU_FISSION: return s.X.nf * kv;
    case MCB_SCORE_TOTAL: return s.X.t * kv;
    default: return 0.0;
    }
}
__device__ void estimator_score(const DevProblem& P, const TallyAcc& T, int e, const ScoreState& s, double l, int hist)
{
    const mcb_estimator E = P.estimators[e];
    int64_t idx_1D = 0;
    int64_t factor_next = 1;  // idx_factor[i+1] (Estimator.cpp:288-295), built from the last filter backwards
    for (int i = E.n_filters - 1; i >= 0; i--) {
        const mcb_filter F = P.filters[E.filter_begin + i];
        const double* g = P.filter_grid + F.grid_begin;
        int idx;
        switch (F.type) {
        case MCB_FILTER_SURFACE: idx = mcb_binary_search((double)s.surface_old, g, F.grid_n) + 1; break;  // Estimator.cpp:133-140
        case MCB_FILTER_CELL: idx = mcb_binary_search((double)s.cell, g, F.grid_n) + 1; break;            // :141-148
        default: {                                                                                         // energy :149-163
            idx = mcb_binary_search(s.E, g, F.grid_n);
            if (idx < 0 || idx >= F.grid_n - 1) return;
        }
        }
        idx_1D += (int64_t)idx * factor_next;
        factor_next *= F.size;
    }
    // factor_next is now idx_factor[0], the stride between scores
    double* acc = T.acc + (int64_t)(hist - T.first_hist);
    for (int k = 0; k < E.n_scores; k++) {
        const double v = score_value(P, P.scores[E.score_begin + k], s, l);
        const int64_t t = E.tally_begin + idx_1D + (int64_t)k * factor_next;
        if (t >= E.tally_begin && t < E.tally_begin + E.n_tallies) atomicAdd(acc + t * T.stride, v);
    }
}
__device__ __forceinline__ void score_attached(const DevProblem& P, const TallyAcc& T, int kind, int id,
                                               const ScoreState& s, double l, int hist)
{
    const int b = P.attach_begin[kind][id], e = P.attach_begin[kind][id + 1];
    for (int i = b; i < e; i++) estimator_score(P, T, P.attach_list[kind][i], s, l, hist);
}
__device__ __forceinline__ bool has_attached(const DevProblem& P, int kind, int id)
{
    return P.attach_begin[kind][id + 1] > P.attach_begin[kind][id];
}

// ---------------------------------------------------------------------------------------------
// stage kernels
// ---------------------------------------------------------------------------------------------
// source: SourceBank::get_source (Source.cpp:42-46) with j = floor(xi*N), SourcePoint / SourceDelta (Source.cpp:16-24)
// History h (shard-local) of this cycle gets the stream of nps = cycle*Nsample + (shard_begin + h)
// (RN_init_particle, Random.cpp:196-204).
__global__ void __launch_bounds__(256)
k_source(const DevProblem P, Bank B, uint32_t* active, int32_t first_hist, uint32_t count, uint64_t nps0,
         const Site* sbank, uint64_t n_sbank)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= count) return;
    const int32_t h = first_hist + (int32_t)q;
    uint64_t rng = mcb_rn_history_seed(P.seed0, nps0 + (uint64_t)h);
    const double xi = mcb_urand(rng);
    double x, y, z, u, v, w, E, t;
    int cell;
    if (sbank) {
        uint64_t j = (uint64_t)(xi * (double)n_sbank);
        if (j >= n_sbank) j = n_sbank - 1;
        const Site s = sbank[j];
        x = s.x; y = s.y; z = s.z; u = s.u; v = s.v; w = s.w; E = s.E; t = s.t; cell = s.cell;
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
    active[q] = q;
}

// xs_lookup stage: macroscopic cross sections of every queued particle at its energy in its cell's material
__global__ void __launch_bounds__(256)
k_xs_stage(const DevProblem P, Bank B, const uint32_t* __restrict__ active, uint32_t n, Counters* C)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = q < n;
    bool looked = false;
    if (valid) {
        const uint32_t i = active[q];
        const int m = P.cells[B.cell[i]].material;
        if (m >= 0) {
            const double E = B.E[i];
            const DevMaterial M = P.materials[m];
            const int u = union_index(M, E);
            MacroXS X;
            macro_xs(P, M, u, E, X);
            B.St[i] = X.t; B.Ss[i] = X.s; B.Sc[i] = X.c; B.Sf[i] = X.f; B.nSf[i] = X.nf;
            B.uidx[i] = u;
            looked = true;
        }
    }
    count_add(&C->n_lookups, looked);
}

// flight stage: surface_intersect + collision_distance + move_particle (general.cpp:40-83,177-207).
// The event queue is split in place: collisions from the front, surface hits from the back.
__global__ void __launch_bounds__(256)
k_flight(const DevProblem P, Bank B, const uint32_t* __restrict__ active, uint32_t n, uint32_t* evq, Counters* C,
         HistoryAcc H, TallyAcc T)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = q < n;
    bool to_collide = false, to_cross = false;
    uint32_t i = 0;
    if (valid) {
        i = active[q];
        const int cell = B.cell[i];
        const int m = P.cells[cell].material;
        double x = B.x[i], y = B.y[i], z = B.z[i];
        const double u = B.u[i], v = B.v[i], w = B.w[i];
        uint64_t rng = B.rng[i];
        double dsurf;
        const int S = surface_intersect(P, cell, x, y, z, u, v, w, dsurf);
        double dcol;
        if (m >= 0) dcol = -log(mcb_urand(rng)) / B.St[i];   // exponential_sample (Algorithm.cpp:123-126)
        else dcol = MCB_MAX_FLOAT_LESS;                        // vacuum (general.cpp:44-46)
        const double l = (dcol > dsurf) ? dsurf : dcol;
        to_cross = dcol > dsurf;
        to_collide = !to_cross;
        // Particle::move (Particle.cpp:66-76)
        x += u * l; y += v * l; z += w * l;
        const double speed = B.speed[i];
        const double t = B.t[i] + l / speed;
        B.x[i] = x; B.y[i] = y; B.z[i] = z; B.t[i] = t; B.rng[i] = rng;
        B.surf[i] = S;
        const int h = B.hist[i];
        const double wgt = B.wgt[i];
        if (P.ksearch && m >= 0) hist_add(&H.kTL[h], B.nSf[i] * wgt * l, P.shared_histories);  // estimate_TL (Estimator.cpp:509-512)
        if (T.on && has_attached(P, MCB_ATTACH_CELL_TL, cell)) {
            ScoreState s;
            s.w = wgt; s.E = B.E[i]; s.speed = speed; s.cell = cell; s.surface_old = -1; s.material = m;
            if (m >= 0) { s.u = B.uidx[i]; s.X.t = B.St[i]; s.X.s = B.Ss[i]; s.X.c = B.Sc[i]; s.X.f = B.Sf[i]; s.X.nf = B.nSf[i]; }
            score_attached(P, T, MCB_ATTACH_CELL_TL, cell, s, l, h);
        }
        if (to_cross && S < 0) { to_cross = false; B.wgt[i] = 0.0; }  // no surface ahead: cannot happen in a closed geometry
    }
    count_add(&C->n_tracks, valid);
    const unsigned pc = warp_append(&C->q_collide, to_collide);
    if (to_collide) evq[pc] = i;
    const unsigned px = warp_append(&C->q_cross, to_cross);
    if (to_cross) evq[n - 1 - px] = i;
}

// collide stage: Simulator::collision (general.cpp:121-163) + weight_roulette (population_control.cpp:9-15)
__global__ void __launch_bounds__(256)
k_collide(const DevProblem P, Bank B, const uint32_t* __restrict__ evq, Counters* C, uint32_t* next, HistoryAcc H,
          TallyAcc T, Site* tmp_sites, int32_t* tmp_hist, uint64_t site_cap, uint32_t n_slots, double k_eff)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = q < C->q_collide;
    uint32_t i = 0;
    int h = 0, cell = 0, m = -1, u = -1, N_fission = -1;
    double E = 0, wgt = 0, x = 0, y = 0, z = 0, t = 0;
    uint64_t rng = 0;
    MacroXS X = {0, 0, 0, 0, 0};
    unsigned bank_nu = 0, n_sites = 0, n_second = 0;
    bool alive = false, in_material = false;
    if (valid) {
        i = evq[q];
        cell = B.cell[i];
        m = P.cells[cell].material;
        in_material = m >= 0;
        h = B.hist[i];
        if (!in_material) { B.wgt[i] = 0.0; }  // vacuum: kill (general.cpp:124-128)
        else {
            E = B.E[i]; wgt = B.wgt[i]; rng = B.rng[i]; u = B.uidx[i];
            X.t = B.St[i]; X.s = B.Ss[i]; X.c = B.Sc[i]; X.f = B.Sf[i]; X.nf = B.nSf[i];
            if (T.on && has_attached(P, MCB_ATTACH_CELL_C, cell)) {
                ScoreState s;
                s.w = wgt; s.E = E; s.speed = B.speed[i]; s.cell = cell; s.surface_old = -1; s.material = m; s.u = u; s.X = X;
                score_attached(P, T, MCB_ATTACH_CELL_C, cell, s, 0.0, h);
            }
            // floor( w/k * nuSigmaF / SigmaT + xi ) (general.cpp:135-136)
            const double a = wgt / k_eff * X.nf / X.t;
            const double bn = floor(a + mcb_urand(rng));
            bank_nu = bn > 0.0 ? (unsigned)bn : 0u;
            N_fission = select_nuclide(P, P.materials[m], u, E, 1, X.nf, mcb_urand(rng), nullptr);  // Material.cpp:116-125
            if (N_fission >= 0) {
                // prompt or delayed (ksearch.cpp:24-38, fixed_source.cpp:12,41-52)
                const DevMaterial& M = P.materials[m];
                int ln = 0;
                for (int n = 0; n < M.n_nuc; n++) if (P.mat_nuclide[M.nuc_begin + n] == N_fission) { ln = n; break; }
                const double beta = micro_col(P.nuclides[N_fission], nuclide_index(M, u, ln), E, 1);
                const bool prompt = mcb_urand(rng) > beta;
                if (P.ksearch) {
                    if (!prompt) (void)mcb_urand(rng);  // precursor group pick, result unused
                    n_sites = bank_nu;
                } else if (prompt) {
                    n_second = bank_nu;
                } else {
                    // delayed, non-TDMC branch: draws are consumed, no neutron is banked (fixed_source.cpp:41-63)
                    (void)mcb_urand(rng);
                    for (unsigned b = 0; b < bank_nu; b++) { (void)mcb_urand(rng); (void)mcb_urand(rng); }
                }
            }
            x = B.x[i]; y = B.y[i]; z = B.z[i]; t = B.t[i];
        }
    }
    count_add(&C->n_collisions, valid && in_material);
    // reserve space for everything this warp banks: one atomic per warp
    const unsigned long long site0 = warp_reserve<unsigned long long>(&C->site_cursor, n_sites);
    const unsigned slot0 = warp_reserve<unsigned int>(&C->slot_cursor, n_second);
    if (valid && in_material) {
        if (n_sites) {
            // implicit_fission_ksearch (ksearch.cpp:39-46): bank_nu sites, Watt energy then isotropic direction, w = 1
            int seq0;
            if (P.shared_histories) seq0 = atomicAdd(&H.nsite[h], (int)n_sites);
            else { seq0 = H.nsite[h]; H.nsite[h] = seq0 + (int)n_sites; }
            const DevNuclide& N = P.nuclides[N_fission];
            for (unsigned b = 0; b < n_sites; b++) {
                Site s;
                s.E = watt_sample(N.watt_a, N.watt_b, N.watt_g, E, rng);
                isotropic_direction(rng, s.u, s.v, s.w);
                s.x = x; s.y = y; s.z = z; s.t = t; s.cell = cell; s.seq = seq0 + (int)b;
                if (site0 + b < site_cap) { tmp_sites[site0 + b] = s; tmp_hist[site0 + b] = h; }
                else C->overflow_sites = 1;
            }
        }
        if (n_second) {
            // implicit_fission_fixed_source, prompt branch (fixed_source.cpp:12-22): same-history secondaries
            const DevNuclide& N = P.nuclides[N_fission];
            for (unsigned b = 0; b < n_second; b++) {
                const double Es = watt_sample(N.watt_a, N.watt_b, N.watt_g, E, rng);
                double du, dv, dw;
                isotropic_direction(rng, du, dv, dw);
                const unsigned j = slot0 + b;
                if (j < n_slots) {
                    B.x[j] = x; B.y[j] = y; B.z[j] = z; B.u[j] = du; B.v[j] = dv; B.w[j] = dw;
                    B.E[j] = Es; B.speed[j] = mcb_speed_of_energy(Es); B.wgt[j] = 1.0; B.t[j] = t;
                    B.rng[j] = mcb_rn_child_seed(rng, b); B.cell[j] = cell; B.hist[j] = h;
                } else C->overflow_slots = 1;
            }
        }
        if (P.ksearch && N_fission >= 0) hist_add(&H.kC[h], X.nf * wgt / X.t, P.shared_histories);  // estimate_C (Estimator.cpp:503-507)
        // implicit absorption (general.cpp:154-156)
        const double implicit = X.c + X.f;
        wgt = wgt * (X.t - implicit) / X.t;
        alive = true;
        const int N_scatter = select_nuclide(P, P.materials[m], u, E, 0, X.s, mcb_urand(rng), nullptr);  // Material.cpp:106-115
        if (N_scatter >= 0) {
            double du = B.u[i], dv = B.v[i], dw = B.w[i], speed = B.speed[i];
            scatter_sample(P.nuclides[N_scatter].A, du, dv, dw, E, speed, rng);
            B.u[i] = du; B.v[i] = dv; B.w[i] = dw; B.E[i] = E; B.speed[i] = speed;
        }
        // weight_roulette (population_control.cpp:9-15)
        if (wgt < P.wr) {
            if (mcb_urand(rng) < wgt / P.ws) wgt = P.ws;
            else { wgt = 0.0; alive = false; }
        }
        B.wgt[i] = wgt; B.rng[i] = rng;
    }
    // survivors and their secondaries go to the next iteration's queue
    const unsigned pn = warp_append(&C->q_next, alive);
    if (alive) next[pn] = i;
    if (__any_sync(FULL, n_second > 0)) {
        unsigned ok = 0;
        for (unsigned b = 0; b < n_second; b++) if (slot0 + b < n_slots) ok++;
        const unsigned p0 = warp_reserve<unsigned int>(&C->q_next, ok);
        for (unsigned b = 0; b < ok; b++) next[p0 + b] = slot0 + b;
    }
}

// cross stage: surface_hit (general.cpp:89-115) + cell_importance + weight_roulette (population_control.cpp:9-49)
__global__ void __launch_bounds__(256)
k_cross(const DevProblem P, Bank B, const uint32_t* __restrict__ evq, uint32_t n_active, Counters* C, uint32_t* next,
        TallyAcc T, uint32_t n_slots)
{
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = q < C->q_cross;
    uint32_t i = 0;
    bool alive = false;
    unsigned n_copy = 0;
    double x = 0, y = 0, z = 0, u = 0, v = 0, w = 0, t = 0, wgt = 0, E = 0, speed = 0;
    uint64_t rng = 0;
    int cell = 0, h = 0;
    if (valid) {
        i = evq[n_active - 1 - q];
        const int S = B.surf[i];
        const mcb_surface& Sf = P.surfaces[S];
        x = B.x[i]; y = B.y[i]; z = B.z[i]; u = B.u[i]; v = B.v[i]; w = B.w[i]; t = B.t[i];
        wgt = B.wgt[i]; rng = B.rng[i]; speed = B.speed[i]; E = B.E[i];
        cell = B.cell[i]; h = B.hist[i];
        int cell_old = cell;
        alive = true;
        if (Sf.bc == MCB_BC_TRANSMISSION) {
            x += u * MCB_EPSILON_FLOAT; y += v * MCB_EPSILON_FLOAT; z += w * MCB_EPSILON_FLOAT;
            t += MCB_EPSILON_FLOAT / speed;
            const int cn = mcb_search_cell(P.cells, P.n_cells, P.surfaces, P.cell_surface, P.cell_sense, x, y, z);
            if (cn < 0) {  // "[WARNING] A particle is lost" (general.cpp:31-33)
                if (atomicExch(&C->lost, 1) == 0) { C->lost_pos[0] = x; C->lost_pos[1] = y; C->lost_pos[2] = z; }
                alive = false; wgt = 0.0;
            } else cell = cn;
        } else if (Sf.bc == MCB_BC_VACUUM) {
            alive = false; wgt = 0.0;
        } else {
            mcb_surf_reflect(Sf, u, v, w);
            x += u * MCB_EPSILON_FLOAT; y += v * MCB_EPSILON_FLOAT; z += w * MCB_EPSILON_FLOAT;
            t += MCB_EPSILON_FLOAT / speed;
        }
        if (T.on && has_attached(P, MCB_ATTACH_SURFACE, S)) {
            ScoreState s;
            s.w = wgt; s.E = E; s.speed = speed; s.cell = cell; s.surface_old = S; s.material = P.cells[cell].material;
            s.u = -1; s.X = MacroXS{0, 0, 0, 0, 0};
            if (s.material >= 0) { s.u = union_index(P.materials[s.material], E); macro_xs(P, P.materials[s.material], s.u, E, s.X); }
            score_attached(P, T, MCB_ATTACH_SURFACE, S, s, 0.0, h);
        }
        // cell_importance (population_control.cpp:21-49)
        const double Iold = P.cells[cell_old].importance, Inew = P.cells[cell].importance;
        if (Inew != Iold) {
            const double rat = Inew / Iold;
            if (rat < 1.0) {
                if (mcb_urand(rng) < rat) wgt = wgt / rat;
                else { alive = false; wgt = 0.0; }
            } else {
                const int ns = (int)floor(rat + mcb_urand(rng));
                wgt = wgt / (double)ns;
                n_copy = ns > 1 ? (unsigned)(ns - 1) : 0u;
            }
        }
    }
    count_add(&C->n_crossings, valid);
    const unsigned slot0 = warp_reserve<unsigned int>(&C->slot_cursor, n_copy);
    unsigned ok = 0;
    if (valid) {
        // split copies carry the state after the importance draw; stream j = (j+1)*2^40 draws ahead
        for (unsigned b = 0; b < n_copy; b++) {
            const unsigned j = slot0 + b;
            if (j < n_slots) {
                B.x[j] = x; B.y[j] = y; B.z[j] = z; B.u[j] = u; B.v[j] = v; B.w[j] = w;
                B.E[j] = E; B.speed[j] = speed; B.wgt[j] = wgt; B.t[j] = t;
                B.rng[j] = mcb_rn_child_seed(rng, b); B.cell[j] = cell; B.hist[j] = h;
                ok++;
            } else C->overflow_slots = 1;
        }
        // weight_roulette: also draws for a particle that was just killed (w = 0 < wr), like the reference
        if (wgt < P.wr) {
            if (mcb_urand(rng) < wgt / P.ws) wgt = P.ws;
            else { wgt = 0.0; alive = false; }
        }
        B.x[i] = x; B.y[i] = y; B.z[i] = z; B.u[i] = u; B.v[i] = v; B.w[i] = w; B.t[i] = t;
        B.wgt[i] = wgt; B.rng[i] = rng; B.cell[i] = cell;
    }
    const unsigned pn = warp_append(&C->q_next, alive);
    if (alive) next[pn] = i;
    if (__any_sync(FULL, n_copy > 0)) {
        const unsigned p0 = warp_reserve<unsigned int>(&C->q_next, ok);
        for (unsigned b = 0; b < ok; b++) next[p0 + b] = slot0 + b;
    }
}

// ---------------------------------------------------------------------------------------------
// generation close-out
// ---------------------------------------------------------------------------------------------
// fission bank in canonical order (parent history, banking order): position = offset[hist] + seq
__global__ void __launch_bounds__(256)
k_bank_order(const Site* __restrict__ tmp, const int32_t* __restrict__ tmp_hist, uint64_t n,
             const uint32_t* __restrict__ offset, Site* out)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const Site s = tmp[q];
    out[(uint64_t)offset[tmp_hist[q]] + (uint64_t)s.seq] = s;
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

void source(cudaStream_t st, const DevProblem& P, const Bank& B, uint32_t* active, int32_t first_hist, uint32_t count,
            uint64_t nps0, const Site* sbank, uint64_t n_sbank)
{
    if (count) k_source<<<blocks_for(count), 256, 0, st>>>(P, B, active, first_hist, count, nps0, sbank, n_sbank);
}
void xs_stage(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* active, uint32_t n, Counters* C)
{
    if (n) k_xs_stage<<<blocks_for(n), 256, 0, st>>>(P, B, active, n, C);
}
void flight(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* active, uint32_t n, uint32_t* evq,
            Counters* C, const HistoryAcc& H, const TallyAcc& T)
{
    if (n) k_flight<<<blocks_for(n), 256, 0, st>>>(P, B, active, n, evq, C, H, T);
}
void collide(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* evq, uint32_t n_upper, Counters* C,
             uint32_t* next, const HistoryAcc& H, const TallyAcc& T, Site* tmp_sites, int32_t* tmp_hist,
             uint64_t site_cap, uint32_t n_slots, double k_eff)
{
    if (n_upper) k_collide<<<blocks_for(n_upper), 256, 0, st>>>(P, B, evq, C, next, H, T, tmp_sites, tmp_hist, site_cap, n_slots, k_eff);
}
void cross(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* evq, uint32_t n_active,
           uint32_t n_upper, Counters* C, uint32_t* next, const TallyAcc& T, uint32_t n_slots)
{
    if (n_upper) k_cross<<<blocks_for(n_upper), 256, 0, st>>>(P, B, evq, n_active, C, next, T, n_slots);
}
void bank_order(cudaStream_t st, const Site* tmp, const int32_t* tmp_hist, uint64_t n, const uint32_t* offset, Site* out)
{
    if (n) k_bank_order<<<blocks_for(n), 256, 0, st>>>(tmp, tmp_hist, n, offset, out);
}
size_t scan_temp_bytes(uint32_t n)
{
    size_t bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return bytes;
}
void scan_sites(cudaStream_t st, void* temp, size_t temp_bytes, const int32_t* nsite, uint32_t* offset, uint32_t n)
{
    if (n) cub::DeviceScan::ExclusiveSum(temp, temp_bytes, nsite, offset, (int)n, st);
}
void reduce_k(cudaStream_t st, const double* kC, const double* kTL, uint32_t n, Counters* C)
{
    if (n) k_reduce_k<<<min(blocks_for(n), 148u * 8u), 256, 0, st>>>(kC, kTL, n, C);
}
void entropy_history(cudaStream_t st, const DevProblem& P, const Site* bank, const uint32_t* offset,
                     const int32_t* nsite, uint32_t n_hist, Counters* C)
{
    if (n_hist) k_entropy_history<<<min(blocks_for(n_hist), 148u * 8u), 256, 0, st>>>(P, bank, offset, nsite, n_hist, C);
}
void entropy_histogram(cudaStream_t st, const DevProblem& P, const Site* bank, uint64_t n, unsigned long long* bins)
{
    if (n) k_entropy_histogram<<<min(blocks_for(n), 148u * 8u), 256, 0, st>>>(P, bank, n, bins);
}
int tally_chunks(uint32_t n_hist) { return (int)((n_hist + TALLY_CHUNK - 1) / TALLY_CHUNK); }
void tally_reduce(cudaStream_t st, double* acc, int64_t stride, uint32_t n_hist, int64_t n_tallies, double* partial,
                  double* sum, double* squared)
{
    if (!n_hist || !n_tallies) return;
    const int nc = tally_chunks(n_hist);
    k_tally_partial<<<dim3(nc, (unsigned)n_tallies), 256, 0, st>>>(acc, stride, n_hist, partial, nc);
    k_tally_final<<<blocks_for(n_tallies, 128), 128, 0, st>>>(partial, nc, n_tallies, sum, squared);
}
void iota(cudaStream_t st, uint32_t* a, uint32_t n)
{
    if (n) k_iota<<<blocks_for(n), 256, 0, st>>>(a, n);
}

void xs_lookup(cudaStream_t st, const DevProblem& P, int material, const double* E, int64_t n, double* out5)
{
    if (n) k_xs_lookup<<<blocks_for(n), 256, 0, st>>>(P, material, E, n, out5);
}
void select_channel(cudaStream_t st, const DevProblem& P, int material, int kind, const double* E, const double* xi,
                    int64_t n, int32_t* out)
{
    if (n) k_select_channel<<<blocks_for(n), 256, 0, st>>>(P, material, kind, E, xi, n, out);
}
void beta(cudaStream_t st, const DevProblem& P, int material, int local_n, const double* E, int64_t n, double* out)
{
    if (n) k_beta<<<blocks_for(n), 256, 0, st>>>(P, material, local_n, E, n, out);
}
void rng(cudaStream_t st, uint64_t seed0, const uint64_t* nps, int64_t n, int ndraw, uint64_t* out)
{
    if (n) k_rng<<<blocks_for(n), 256, 0, st>>>(seed0, nps, n, ndraw, out);
}
void geometry(cudaStream_t st, const DevProblem& P, const int32_t* cell, const double* pos, const double* dir,
              int64_t n, double* out3)
{
    if (n) k_geometry<<<blocks_for(n), 256, 0, st>>>(P, cell, pos, dir, n, out3);
}
void search_cell(cudaStream_t st, const DevProblem& P, const double* pos, int64_t n, int32_t* out)
{
    if (n) k_search_cell<<<blocks_for(n), 256, 0, st>>>(P, pos, n, out);
}
void scatter(cudaStream_t st, const DevProblem& P, int nuclide, const uint64_t* nps, int64_t n, double* io5)
{
    if (n) k_scatter<<<blocks_for(n), 256, 0, st>>>(P, nuclide, nps, n, io5);
}
void watt(cudaStream_t st, const DevProblem& P, int nuclide, const uint64_t* nps, const double* E, int64_t n, double* out)
{
    if (n) k_watt<<<blocks_for(n), 256, 0, st>>>(P, nuclide, nps, E, n, out);
}

}  // namespace mcbk
