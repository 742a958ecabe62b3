// mcb_walk_roles.cu — k_walk_roles: the transport loop with SMs specialised by event type (MCB_WALK_FORM=roles).
//
// Why: both forms of k_walk (mcb_walk.cu) are bound by instruction fetch — 70-78 KB of loop body against a 32 KB per-SM
// instruction cache, GPC-level instruction-cache requests at 97 % of peak (profiles/r2c_k_walk_ncu_full_summary.txt).
// Here an SM runs ONE part of the loop: "flight" SMs do lookup + flight (+ the short surface-crossing code), "collide"
// SMs do the collision (fission banking, implicit capture, free-gas kinematics); each part is 25-30 KB of code.
// Particles live in slots of a global, L2-resident state array (the event-sorted form of k_walk with its state in
// global memory runs at the speed of the shared-memory one) and move between SMs as tiles of 32 slot numbers through
// two global multi-producer multi-consumer rings, QF (needs lookup + flight) and QC (needs a collision):
//   flight SM:   pop a tile from QF -> lookup, flight, crossings -> colliders to QC, the others back to QF
//   collide SM:  pop a tile from QC -> collision -> survivors to QF
// A lane whose history ends takes the next source particle from the bank into the same slot (-> QF).  Warps collect
// slot numbers per destination in a shared-memory stage of their block and push full tiles, so batches stay dense.
// A warp whose own ring is empty serves the other ring (the roles are a preference: load balance comes for free, the
// instruction cache is only polluted when the split is off).  k-eigenvalue problems without splitting only (one
// particle per history); every wait is bounded.
#include "mcb_events.cuh"

#include <algorithm>
#include <cstring>

using namespace mcbe;

namespace {

constexpr int R_BLOCK = 128, R_WARPS = R_BLOCK / 32;
constexpr int R_BATCHES = 2;  // slots per lane: 2 x 32 per warp
constexpr int R_WAIT_LIMIT = 1 << 19;

enum { RP_XY = 0, RP_ZU, RP_VW, RP_ES, RP_WT, RP_RNG, RP_K, RP_IDS, RP_XT, RP_XS, RP_XF, RP_FIXED };

__device__ __forceinline__ double rpack2i(int lo, int hi) { return __hiloint2double(hi, lo); }

struct GDetail {  // per-nuclide partial sums of the last lookup, in the particle's slot of the global state array
    static constexpr bool present = true;
    double2* base;
    uint32_t stride;  // slots in all
    int nn;
    __device__ __forceinline__ void set(int n, double s, double nf, double be) const
    {
        __stcg(base + (size_t)n * stride, make_double2(s, nf));
        __stcg(reinterpret_cast<double*>(base + (size_t)(nn + (n >> 1)) * stride) + (n & 1), be);
    }
    __device__ __forceinline__ double cum_s(int n) const { return __ldcg(base + (size_t)n * stride).x; }
    __device__ __forceinline__ double cum_nf(int n) const { return __ldcg(base + (size_t)n * stride).y; }
    __device__ __forceinline__ double beta(int n) const
    {
        const double2 v = __ldcg(base + (size_t)(nn + (n >> 1)) * stride);
        return (n & 1) ? v.y : v.x;
    }
};

// ---- global rings of tiles
__device__ __forceinline__ void q_push(mcbk::RoleQueue* Q, uint32_t id, unsigned n, unsigned lane, Counters* C)
{
    unsigned long long pos = 0;
    if (lane == 0) pos = atomicAdd(&Q->tail, 1ull);
    pos = __shfl_sync(FULL, pos, 0);
    const uint32_t cell = (uint32_t)pos & Q->cap_mask;
    if (lane == 0) {
        volatile unsigned long long* seq = Q->seq + cell;
        int spins = 0;
        while (*seq != pos) { if (++spins > R_WAIT_LIMIT) { C->hang = 5; break; } __nanosleep(100); }
    }
    __syncwarp();
    if (lane < n) __stcg(Q->ids + (size_t)cell * 32 + lane, id);
    if (lane == 0) __stcg(Q->cnt + cell, n);
    __threadfence();
    __syncwarp();
    if (lane == 0) {
        *(volatile unsigned long long*)(Q->seq + cell) = pos + 1ull;
        atomicAdd(&Q->avail, 1);
    }
}
__device__ __forceinline__ unsigned q_pop(mcbk::RoleQueue* Q, uint32_t& id, unsigned lane, Counters* C)
{
    int ok = 0;
    unsigned long long pos = 0;
    if (lane == 0 && *(volatile int*)&Q->avail > 0) {  // look before taking: idle warps must not hammer the counters with atomics
        if (atomicSub(&Q->avail, 1) <= 0) atomicAdd(&Q->avail, 1);
        else { ok = 1; pos = atomicAdd(&Q->head, 1ull); }
    }
    ok = __shfl_sync(FULL, ok, 0);
    if (!ok) return 0;
    pos = __shfl_sync(FULL, pos, 0);
    const uint32_t cell = (uint32_t)pos & Q->cap_mask;
    if (lane == 0) {
        volatile unsigned long long* seq = Q->seq + cell;
        int spins = 0;
        while (*seq != pos + 1ull) { if (++spins > R_WAIT_LIMIT) { C->hang = 6; break; } __nanosleep(100); }
    }
    __syncwarp();
    __threadfence();
    const unsigned n = __ldcg(Q->cnt + cell);
    if (lane < n) id = __ldcg(Q->ids + (size_t)cell * 32 + lane);
    __threadfence();
    __syncwarp();
    if (lane == 0) *(volatile unsigned long long*)(Q->seq + cell) = pos + (unsigned long long)Q->cap_mask + 1ull;
    return n;
}

// ---- per-block stage: slot numbers waiting to fill a tile
struct Stage {
    unsigned lock;
    unsigned n[2];               // 0: to QF, 1: to QC
    uint32_t id[2][R_BLOCK + 32];
};
__device__ __forceinline__ void stage_lock(Stage& S, unsigned lane)
{
    if (lane == 0) { while (atomicCAS(&S.lock, 0u, 1u) != 0u) __nanosleep(32); }
    __syncwarp();
    __threadfence_block();
}
__device__ __forceinline__ void stage_unlock(Stage& S, unsigned lane)
{
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicExch(&S.lock, 0u);
}
// every lane hands in at most one slot for destination dest (0 QF, 1 QC, -1 none); full tiles go to the rings
__device__ __forceinline__ void stage_push(Stage& S, mcbk::RoleQueue* const (&Q)[2], uint32_t slot, int dest, unsigned lane, Counters* C)
{
    const unsigned lt = (1u << lane) - 1u;
    uint32_t tile_id[2] = {0, 0};
    bool full[2] = {false, false};
    __threadfence();  // the slot's state must be visible device-wide before its number can leave through ANY warp of the block
    stage_lock(S, lane);
#pragma unroll
    for (int d = 0; d < 2; d++) {
        const unsigned m = __ballot_sync(FULL, dest == d);
        unsigned cur = S.n[d];
        if (dest == d) S.id[d][cur + __popc(m & lt)] = slot;
        cur += __popc(m);
        __syncwarp();
        if (cur >= 32u) { tile_id[d] = S.id[d][cur - 32u + lane]; cur -= 32u; full[d] = true; }
        __syncwarp();
        if (lane == 0) S.n[d] = cur;
    }
    stage_unlock(S, lane);
#pragma unroll
    for (int d = 0; d < 2; d++)
        if (full[d]) q_push(Q[d], tile_id[d], 32u, lane, C);
}
// an idle warp sends on whatever waits in its block's stage
__device__ __forceinline__ void stage_flush(Stage& S, mcbk::RoleQueue* const (&Q)[2], unsigned lane, Counters* C)
{
    uint32_t tile_id[2] = {0, 0};
    unsigned cnt[2] = {0, 0};
    stage_lock(S, lane);
#pragma unroll
    for (int d = 0; d < 2; d++) {
        cnt[d] = S.n[d] < 32u ? S.n[d] : 32u;
        if (lane < cnt[d]) tile_id[d] = S.id[d][S.n[d] - cnt[d] + lane];
        __syncwarp();
        if (lane == 0) S.n[d] -= cnt[d];
    }
    stage_unlock(S, lane);
#pragma unroll
    for (int d = 0; d < 2; d++)
        if (cnt[d]) q_push(Q[d], tile_id[d], cnt[d], lane, C);
}

template <bool TALLY>
__global__ void __launch_bounds__(R_BLOCK, 4)
k_walk_roles(const DevProblem P, const Bank B, unsigned long long n_bank, uint32_t chunk, Counters* C, HistoryAcc H, TallyAcc T, SiteReq* reqs,
             uint64_t site_cap, double k_eff, mcbk::RolesRes R)
{
    __shared__ Stage S;
    extern __shared__ double r_priv[];
    constexpr int RP_DET = TALLY ? RP_FIXED + 1 : RP_FIXED;
    double* const s_sum = (TALLY && R.priv_tallies) ? r_priv : nullptr;
    double* const s_sq = s_sum ? s_sum + R.priv_tallies : nullptr;
    __shared__ unsigned warps_done;
    if (threadIdx.x == 0) { S.lock = 0; S.n[0] = S.n[1] = 0; warps_done = 0; }
    if (s_sum) for (int i = threadIdx.x; i < 2 * R.priv_tallies; i += R_BLOCK) s_sum[i] = 0.0;
    __syncthreads();

    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const int my_role = ((int)(smid % 20u) * 5 < R.c_pct) ? 1 : 0;  // 1: collide SM, 0: flight SM
    const int n_sh = R.n_shards, my_shard = (int)((blockIdx.x * R_WARPS + (threadIdx.x >> 5)) % (unsigned)R.n_shards);
    mcbk::RoleQueue* const Q[2] = {R.qF + my_shard, R.qC + my_shard};   // this warp pushes to its own shard of each kind
    mcbk::RoleQueue* const Qbase[2] = {R.qF, R.qC};
    double2* const st = R.state;
    const uint32_t NP = R.NP;
    unsigned tracks = 0, collisions = 0, crossings = 0, lookups = 0, steals = 0, tiles = 0, tile_lanes = 0;
    bool exhausted = false;
    unsigned long long chunk_next = 0, chunk_end = 0;
    int idle_spins = 0;
    unsigned idle_pause = 500u;

    // a lane takes the next source particle from the bank into `slot` (registers -> slot state); returns false when the bank is dry
    Particle p;
    HistLocal L = {0.0, 0.0, 0};
    auto store_particle = [&](uint32_t slot, int S_hit, int uidx) {
        double2* s = st + slot;
        __stcg(s + (size_t)RP_XY * NP, make_double2(p.x, p.y));
        __stcg(s + (size_t)RP_ZU * NP, make_double2(p.z, p.u));
        __stcg(s + (size_t)RP_VW * NP, make_double2(p.v, p.w));
        __stcg(s + (size_t)RP_ES * NP, make_double2(p.E, p.speed));
        __stcg(s + (size_t)RP_WT * NP, make_double2(p.wgt, p.t));
        __stcg(s + (size_t)RP_RNG * NP, make_double2(__longlong_as_double((long long)p.rng), rpack2i(p.cell, p.hist)));
        __stcg(s + (size_t)RP_K * NP, make_double2(L.kC, L.kTL));
        __stcg(s + (size_t)RP_IDS * NP, make_double2(rpack2i(L.nsite, 0), rpack2i(S_hit, uidx)));
        if (TALLY) __stcg(s + (size_t)RP_FIXED * NP, make_double2(p.told, rpack2i(p.n_touched, 0)));
    };
    // refill: lanes flagged `want` draw bank positions (warp-level chunks of the global head counter); got = lanes served
    auto refill = [&](bool want) -> bool {
        bool got = false;
        unsigned need = __ballot_sync(FULL, want);
        while (need && !exhausted) {
            if (chunk_next == chunk_end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&C->walk_head, (unsigned long long)chunk);
                base = __shfl_sync(FULL, base, 0);
                chunk_next = base < n_bank ? base : n_bank;
                chunk_end = base + chunk < n_bank ? base + chunk : n_bank;
                if (chunk_next == chunk_end) { exhausted = true; break; }
            }
            const unsigned take = min((unsigned)__popc(need), (unsigned)(chunk_end - chunk_next));
            const unsigned rank = __popc(need & lt_mask);
            if (want && !got && rank < take) {
                const uint32_t j = (uint32_t)(chunk_next + rank);
                p.cell = B.cell[j]; p.hist = B.hist[j];
                p.x = B.x[j]; p.y = B.y[j]; p.z = B.z[j]; p.u = B.u[j]; p.v = B.v[j]; p.w = B.w[j];
                p.E = B.E[j]; p.speed = B.speed[j]; p.wgt = B.wgt[j]; p.t = B.t[j]; p.rng = B.rng[j];
                p.Eold = p.E; p.told = p.t; p.n_touched = 0; p.drow = -1;
                L.kC = 0.0; L.kTL = 0.0; L.nsite = 0;
                got = true;
            }
            chunk_next += take;
            need = __ballot_sync(FULL, want && !got);
        }
        return got;
    };

    // ---- start: every warp brings R_BATCHES x 32 histories into its own slots and queues them for their first flight
    for (int b = 0; b < R_BATCHES; b++) {
        const uint32_t slot = ((blockIdx.x * R_WARPS + (threadIdx.x >> 5)) * R_BATCHES + b) * 32u + lane;
        const bool got = refill(true);
        if (got) { p.row = (int)slot; store_particle(slot, -1, -1); if (TALLY) __stcg(st + slot + (size_t)RP_XF * NP, make_double2(0.0, p.Eold)); }
        const unsigned n_new = __popc(__ballot_sync(FULL, got));
        if (lane == 0 && n_new) atomicAdd((unsigned long long*)&C->live, (unsigned long long)n_new);
        stage_push(S, Q, slot, got ? 0 : -1, lane, C);
    }

    for (;;) {
        // ---- take a tile: the SM's own kind first, the other kind when there is none
        uint32_t slot = 0;
        int kind = my_role;
        unsigned n = 0;
        for (int pass = 0; pass < 2 && !n; pass++, kind ^= 1) {  // own kind first (own shard, then the others), then the other kind
            for (int i = 0; i < n_sh && !n; i++) n = q_pop(Qbase[kind] + (my_shard + i) % n_sh, slot, lane, C);
            if (n) { if (pass) steals++; break; }
        }
        if (!n) {
            stage_flush(S, Q, lane, C);
            long long lv = 0;
            if (lane == 0) lv = (long long)__ldcg((const unsigned long long*)&C->live);
            lv = __shfl_sync(FULL, lv, 0);
            if (lv <= 0) break;
            if (++idle_spins > R_WAIT_LIMIT) { C->hang = 7; break; }
            __nanosleep(idle_pause);
            if (idle_pause < 16000u) idle_pause *= 2u;
            continue;
        }
        idle_pause = 500u;
        tiles++; tile_lanes += n;
        const bool have = lane < n;
        int dest = -1;
        bool ended = false;
        if (have) {
            const double2* s = st + slot;
            double2 v;
            v = __ldcg(s + (size_t)RP_XY * NP); p.x = v.x; p.y = v.y;
            v = __ldcg(s + (size_t)RP_ZU * NP); p.z = v.x; p.u = v.y;
            v = __ldcg(s + (size_t)RP_VW * NP); p.v = v.x; p.w = v.y;
            v = __ldcg(s + (size_t)RP_ES * NP); p.E = v.x; p.speed = v.y;
            v = __ldcg(s + (size_t)RP_WT * NP); p.wgt = v.x; p.t = v.y;
            v = __ldcg(s + (size_t)RP_RNG * NP); p.rng = (uint64_t)__double_as_longlong(v.x); p.cell = __double2loint(v.y); p.hist = __double2hiint(v.y);
            v = __ldcg(s + (size_t)RP_K * NP); L.kC = v.x; L.kTL = v.y;
            p.row = (int)slot; p.drow = -1;
        }
        const GDetail D = {st + (size_t)RP_DET * NP + slot, NP, R.det_nn};
        if (kind == 0) {
            // ---------------- flight SM: lookup + flight, and the crossing for the lanes that reach a surface
            MacroXS X = {0, 0, 0, 0, 0};
            int uidx = -1, S_hit = -1;
            bool alive = true;
            if (have) {
                const double2 v = __ldcg(st + slot + (size_t)RP_IDS * NP);
                L.nsite = __double2loint(v.x);
                if (TALLY) { const double2 w = __ldcg(st + slot + (size_t)RP_FIXED * NP); p.told = w.x; p.n_touched = __double2loint(w.y); p.Eold = p.E; }
                else { p.told = p.t; p.n_touched = 0; p.Eold = p.E; }
                if (TALLY && P.track_old) p.Eold = __ldcg(st + slot + (size_t)RP_XF * NP).y;
                if (ev_lookup(P, p, X, uidx, D)) lookups++;
                const bool to_cross = ev_flight<TALLY>(P, p, X, uidx, H, T, C, S_hit, &L);
                tracks++;
                if (to_cross) {
                    unsigned n_copy = 0;
                    NoSink none;
                    alive = ev_cross_pre<TALLY>(P, p, S_hit, T, C, n_copy);
                    alive = ev_cross_post(P, p, alive, 0u, none);
                    crossings++;
                    dest = 0;
                } else {
                    dest = 1;
                    double2* s = st + slot;
                    __stcg(s + (size_t)RP_XT * NP, make_double2(X.t, X.nf));
                    __stcg(s + (size_t)RP_XS * NP, make_double2(X.s, X.c));
                    __stcg(s + (size_t)RP_XF * NP, make_double2(X.f, p.Eold));
                }
                ended = !alive;
            }
            __syncwarp();
            // histories that ended at a surface: close them out, take the next source particle into the slot
            if (ended && P.ksearch) { H.kC[p.hist] = L.kC; H.kTL[p.hist] = L.kTL; H.nsite[p.hist] = L.nsite; }
            if (TALLY) {
                unsigned m = __ballot_sync(FULL, ended && p.n_touched > 0);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    flush_history_tallies(T, __shfl_sync(FULL, p.row, src), __shfl_sync(FULL, p.n_touched, src), s_sum, s_sq, lane);
                    __syncwarp();
                }
            }
            const bool got = refill(ended);
            const unsigned n_gone = __popc(__ballot_sync(FULL, ended && !got));
            if (lane == 0 && n_gone) atomicAdd((unsigned long long*)&C->live, (unsigned long long)(-(long long)n_gone));
            if (ended) dest = got ? 0 : -1;
            if (have && dest >= 0) {
                p.row = (int)slot;
                store_particle(slot, S_hit, uidx);
                if (TALLY && dest == 0) __stcg(st + slot + (size_t)RP_XF * NP, make_double2(0.0, p.Eold));
            }
        } else {
            // ---------------- collide SM
            MacroXS X = {0, 0, 0, 0, 0};
            CollideCtx c = {-1, -1, 0, 0, 0.0};
            int uidx = -1;
            bool alive = false, in_material = false;
            if (have) {
                const double2* s = st + slot;
                double2 v;
                v = __ldcg(s + (size_t)RP_IDS * NP); L.nsite = __double2loint(v.x); uidx = __double2hiint(v.y);
                v = __ldcg(s + (size_t)RP_XT * NP); X.t = v.x; X.nf = v.y;
                v = __ldcg(s + (size_t)RP_XS * NP); X.s = v.x; X.c = v.y;
                v = __ldcg(s + (size_t)RP_XF * NP); X.f = v.x; p.Eold = v.y;
                if (TALLY) { v = __ldcg(s + (size_t)RP_FIXED * NP); p.told = v.x; p.n_touched = __double2loint(v.y); }
                else { p.told = p.t; p.n_touched = 0; }
                in_material = ev_collide_pre<TALLY>(P, p, X, uidx, D, T, C, k_eff, c);
                if (in_material) collisions++;
            }
            __syncwarp();
            unsigned long long site0 = 0;
            if (__any_sync(FULL, c.n_sites != 0u)) {
                unsigned v = c.n_sites;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(FULL, v, d); if (lane >= (unsigned)d) v += t; }
                const unsigned total = __shfl_sync(FULL, v, 31);
                unsigned long long base = 0;
                if (lane == 31) base = atomicAdd(&C->site_cursor, (unsigned long long)total);
                site0 = __shfl_sync(FULL, base, 31) + (v - c.n_sites);
            }
            c.n_second = 0;
            NoSink none;
            if (c.n_sites) ev_collide_bank(P, p, c, H, C, reqs, site_cap, site0, none, &L);
            __syncwarp();
            if (in_material) alive = ev_collide_scatter<TALLY>(P, p, X, uidx, D, c, H, &L);
            __syncwarp();
            ended = have && !alive;
            if (ended && P.ksearch) { H.kC[p.hist] = L.kC; H.kTL[p.hist] = L.kTL; H.nsite[p.hist] = L.nsite; }
            if (TALLY) {
                unsigned m = __ballot_sync(FULL, ended && p.n_touched > 0);
                while (m) {
                    const int src = __ffs(m) - 1;
                    m &= m - 1;
                    flush_history_tallies(T, __shfl_sync(FULL, p.row, src), __shfl_sync(FULL, p.n_touched, src), s_sum, s_sq, lane);
                    __syncwarp();
                }
            }
            const bool got = refill(ended);
            const unsigned n_gone = __popc(__ballot_sync(FULL, ended && !got));
            if (lane == 0 && n_gone) atomicAdd((unsigned long long*)&C->live, (unsigned long long)(-(long long)n_gone));
            dest = have ? ((ended && !got) ? -1 : 0) : -1;
            if (have && dest >= 0) {
                p.row = (int)slot;
                store_particle(slot, -1, uidx);
                if (TALLY) __stcg(st + slot + (size_t)RP_XF * NP, make_double2(0.0, p.Eold));
            }
        }
        stage_push(S, Q, slot, dest, lane, C);
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
        atomicAdd(&C->n_donated, (unsigned long long)tiles);          // statistics (MCB_TRACE_SHARING): tiles taken,
        atomicAdd(&C->n_shared_hist, (unsigned long long)tile_lanes); // particles in them,
        atomicAdd(&C->n_donate_refused, (unsigned long long)steals);  // tiles of the other kind
    }
    if (s_sum) {
        unsigned done = 0;
        __threadfence_block();
        __syncwarp();
        if (lane == 0) done = atomicAdd(&warps_done, 1u);
        done = __shfl_sync(FULL, done, 0);
        if (done == (unsigned)R_WARPS - 1u) {
            __threadfence_block();
            for (int i = (int)lane; i < R.priv_tallies; i += 32) {
                const double a = s_sum[i], b = s_sq[i];
                if (a != 0.0) { atomicAdd(T.sum + i, a); atomicAdd(T.squared + i, b); }
            }
        }
    }
}

}  // namespace

namespace mcbk {

extern thread_local uint64_t g_launches;

int roles_plan(int det_nn, int64_t n_tallies, int n_sm, RolesPlan* out)
{
    RolesPlan& W = *out;
    W.det_nn = std::max(det_nn, 1);
    W.priv_tallies = (n_tallies > 0 && n_tallies <= 256) ? (int)n_tallies : 0;
    W.n_sm = n_sm;
    int blocks = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_walk_roles<false>, R_BLOCK, 0);
    if (e != cudaSuccess) return (int)e;
    int blocks_t = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_t, k_walk_roles<true>, R_BLOCK, (size_t)W.priv_tallies * 16);
    if (e != cudaSuccess) return (int)e;
    W.blocks_per_sm = std::max(1, std::min(std::min(blocks, blocks_t), 4));
    W.grid = n_sm * W.blocks_per_sm;
    W.n_slots = (uint32_t)W.grid * R_WARPS * R_BATCHES * 32u;
    W.n_pairs = RP_FIXED + 1 + W.det_nn + (W.det_nn + 1) / 2;
    W.queue_cap = 1u << 12;  // tiles per shard
    return 0;
}

void walk_roles(cudaStream_t st, const DevProblem& P, const Bank& B, uint64_t n_bank, Counters* C, const HistoryAcc& H, const TallyAcc& T,
                SiteReq* reqs, uint64_t site_cap, double k_eff, const RolesPlan& W, const RolesRes& res)
{
    if (!n_bank) return;
    const uint64_t warps = (uint64_t)W.grid * R_WARPS;
    const uint32_t chunk = (uint32_t)std::max<uint64_t>(32, std::min<uint64_t>(128, n_bank / (warps * 8)));
    RolesRes R = res;
    R.det_nn = W.det_nn; R.n_pairs = W.n_pairs; R.NP = W.n_slots; R.priv_tallies = T.on ? W.priv_tallies : 0;
    if (T.on) k_walk_roles<true><<<W.grid, R_BLOCK, (size_t)R.priv_tallies * 16, st>>>(P, B, (unsigned long long)n_bank, chunk, C, H, T, reqs, site_cap, k_eff, R);
    else k_walk_roles<false><<<W.grid, R_BLOCK, 0, st>>>(P, B, (unsigned long long)n_bank, chunk, C, H, T, reqs, site_cap, k_eff, R);
    g_launches += 1;
}

}  // namespace mcbk
