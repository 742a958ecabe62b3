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
#include "mcb_events.cuh"

#include <algorithm>
#include <cstring>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

using namespace mcbe;

namespace {

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
// particle <-> bank slot (event-queue kernels keep the particle in the SoA bank between events)
__device__ __forceinline__ void load_particle(const DevProblem& P, const Bank& B, uint32_t i, Particle& p)
{
    p.cell = B.cell[i]; p.hist = B.hist[i]; p.row = p.hist; p.n_touched = 0;
    p.x = B.x[i]; p.y = B.y[i]; p.z = B.z[i]; p.u = B.u[i]; p.v = B.v[i]; p.w = B.w[i];
    p.E = B.E[i]; p.speed = B.speed[i]; p.wgt = B.wgt[i]; p.t = B.t[i]; p.rng = B.rng[i];
    p.Eold = P.track_old ? B.Eold[i] : p.E;
    p.told = P.track_time ? B.told[i] : p.t;
}
// same-history secondaries of the event-queue kernels go to bank positions [slot0, ..): the bank is a ring, positions
// grow without bound and slots behind the running pass (ring_begin) are reused
struct BankSink {
    const DevProblem& P;
    const Bank& B;
    unsigned long long slot0, ring_begin;
    uint32_t n_slots;
    Counters* C;
    unsigned n_pushed, n_ok;
    __device__ __forceinline__ BankSink(const DevProblem& P_, const Bank& B_, unsigned long long slot0_, uint32_t n_slots_, Counters* C_,
                                        unsigned long long ring_begin_ = 0)
        : P(P_), B(B_), slot0(slot0_), ring_begin(ring_begin_), n_slots(n_slots_), C(C_), n_pushed(0), n_ok(0) {}
    __device__ __forceinline__ void push(const Particle& q)
    {
        const unsigned long long pos = slot0 + n_pushed;
        n_pushed++;
        if (pos - ring_begin < n_slots) {
            const uint32_t j = (uint32_t)(pos % n_slots);
            B.x[j] = q.x; B.y[j] = q.y; B.z[j] = q.z; B.u[j] = q.u; B.v[j] = q.v; B.w[j] = q.w;
            B.E[j] = q.E; B.speed[j] = q.speed; B.wgt[j] = q.wgt; B.t[j] = q.t;
            B.rng[j] = q.rng; B.cell[j] = q.cell; B.hist[j] = q.hist;
            if (P.track_old) B.Eold[j] = q.Eold;
            if (P.track_time) B.told[j] = q.told;
            n_ok++;
        } else C->overflow_slots = 1;
    }
};

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
       unsigned long long* key, uint32_t* val, uint64_t* rng_after, int key32)
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
    const unsigned long long kq = j >= rot ? j - rot : j + n_bank - rot;
    if (key32) reinterpret_cast<uint32_t*>(key)[q] = (uint32_t)kq; else key[q] = kq;
    val[q] = q; rng_after[q] = rng;
}

__global__ void __launch_bounds__(BLOCK)
k_source(const DevProblem P, Bank B, uint32_t* active, int32_t first_hist, uint32_t count, uint64_t nps0,
         const SourceBankView V, Counters* C, const unsigned long long* __restrict__ sorted_key,
         const uint32_t* __restrict__ sorted_val, const uint64_t* __restrict__ rng_after, unsigned long long rot, uint32_t q0, int key32)
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
        if (sorted_key) { j = (key32 ? (uint64_t)reinterpret_cast<const uint32_t*>(sorted_key)[q] : sorted_key[q]) + rot; if (j >= V.n) j -= V.n; }
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
        if (S.kind == MCB_SRC_DISK_Z) {  // uniform on the disk around (x, y, z) in the plane z = const (mcb200.h)
            const double rho = S.radius * sqrt(mcb_urand(rng));
            double sa, ca;
            circle_from_draw(mcb_urand(rng), sa, ca);
            x += rho * ca; y += rho * sa;
            const int cs = mcb_search_cell(P.cells, P.n_cells, P.surfaces, P.cell_surface, P.cell_sense, x, y, z);
            if (cs >= 0) cell = cs;
            else if (atomicExch(&C->lost, 1) == 0) { C->lost_pos[0] = x; C->lost_pos[1] = y; C->lost_pos[2] = z; }  // general.cpp:31-33
        }
    }
    B.x[q] = x; B.y[q] = y; B.z[q] = z; B.u[q] = u; B.v[q] = v; B.w[q] = w;
    B.E[q] = E; B.speed[q] = mcb_speed_of_energy(E); B.wgt[q] = 1.0; B.t[q] = t;
    B.rng[q] = rng; B.cell[q] = cell; B.hist[q] = h;
    if (B.Eold) B.Eold[q] = E;  // the reference leaves energy_old uninitialised at birth; defined as E here
    if (B.told) B.told[q] = t;
    if (active) active[q] = q;  // event-queue mode: the first queue is the identity
}

// The sorted sweep over a bank spread over several GPUs, for a launch that shares the machine with the walk kernel and
// gets only a few SMs: every thread fetches ITEMS sites with all their loads in flight at once (ITEMS x 64 B per thread,
// 0.5 MB per SM), so that a handful of SMs keep NVLink as busy as the whole GPU does with one site per thread.
template <int ITEMS>
__global__ void __launch_bounds__(256)
k_source_sweep(Bank B, int32_t first_hist, uint32_t q_begin, uint32_t q_end, const SourceBankView V,
               const void* __restrict__ sorted_key, const uint32_t* __restrict__ sorted_val, const uint64_t* __restrict__ rng_after,
               unsigned long long rot, int key32)
{
    const uint32_t tile = q_begin + (blockIdx.x * 256u) * ITEMS + threadIdx.x;
    uint32_t sv[ITEMS];
    unsigned long long j[ITEMS];
    Site s[ITEMS];
    uint64_t rng[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint32_t q = tile + k * 256u;
        if (q < q_end) {
            sv[k] = sorted_val[q];
            j[k] = (key32 ? (unsigned long long)reinterpret_cast<const uint32_t*>(sorted_key)[q] : reinterpret_cast<const unsigned long long*>(sorted_key)[q]) + rot;
            if (j[k] >= V.n) j[k] -= V.n;
        }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint32_t q = tile + k * 256u;
        if (q < q_end) { s[k] = source_bank_site(V, j[k]); rng[k] = rng_after[sv[k]]; }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; k++) {
        const uint32_t q = tile + k * 256u;
        if (q < q_end) {
            double u, v, w;
            source_bank_direction(V, j[k], s[k], u, v, w);
            B.x[q] = s[k].x; B.y[q] = s[k].y; B.z[q] = s[k].z; B.u[q] = u; B.v[q] = v; B.w[q] = w;
            B.E[q] = s[k].E; B.speed[q] = mcb_speed_of_energy(s[k].E); B.wgt[q] = 1.0; B.t[q] = s[k].t;
            B.rng[q] = rng[k]; B.cell[q] = s[k].cell; B.hist[q] = first_hist + (int32_t)sv[k];
        }
    }
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
            NoDetail nd;
            if (ev_lookup(P, p, X, u, nd)) {
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
            load_particle(P, B, i, p);
            MacroXS X;
            X.t = B.St[i]; X.nf = B.nSf[i];
            int uidx = 0;
            if (T.on) { X.s = B.Ss[i]; X.c = B.Sc[i]; X.f = B.Sf[i]; uidx = B.uidx[i]; }
            int S;
            to_cross = ev_flight<true>(P, p, X, uidx, H, T, C, S);
            B.x[i] = p.x; B.y[i] = p.y; B.z[i] = p.z; B.t[i] = p.t; B.rng[i] = p.rng; B.surf[i] = S;
            if (P.track_time) B.told[i] = p.told;
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
    const NoDetail nd;
    for (unsigned long long tile = blockIdx.x; tile * BLOCK < n; tile += gridDim.x, it++) {
        const unsigned long long q = tile * BLOCK + threadIdx.x;
        const bool valid = q < n;
        uint32_t i = 0;
        Particle p;
        MacroXS X = {0, 0, 0, 0, 0};
        CollideCtx c = {-1, -1, 0, 0, 0, 0.0};
        int uidx = -1;
        bool in_material = false;
        if (valid) {
            i = evq[q];
            load_particle(P, B, i, p);
            X.t = B.St[i]; X.s = B.Ss[i]; X.c = B.Sc[i]; X.f = B.Sf[i]; X.nf = B.nSf[i]; uidx = B.uidx[i];
            in_material = ev_collide_pre<true>(P, p, X, uidx, nd, T, C, k_eff, c);
            if (in_material) collisions++;
        }
        const unsigned cntA[2] = {c.n_sites, c.n_second};
        unsigned long long* const curA[2] = {&C->site_cursor, &C->slot_cursor};
        unsigned long long posA[2];
        block_reserve<2>(scratchA[it & 1], cntA, curA, posA);
        bool alive = false;
        BankSink sink(P, B, posA[1], n_slots, C);
        if (c.n_sites | c.n_second) ev_collide_bank(P, p, c, H, C, reqs, site_cap, posA[0], sink);
        __syncwarp();  // the banking lanes rejoin the warp before the scatter kinematics
        if (in_material) alive = ev_collide_scatter<true>(P, p, X, uidx, nd, c, H);
        __syncwarp();
        if (valid) {
            B.u[i] = p.u; B.v[i] = p.v; B.w[i] = p.w; B.E[i] = p.E; B.speed[i] = p.speed; B.wgt[i] = p.wgt; B.rng[i] = p.rng;
            if (P.track_old) B.Eold[i] = p.Eold;
        }
        // survivors and their secondaries go to the next iteration's queue
        const unsigned cntB[1] = {(alive ? 1u : 0u) + sink.n_ok};
        unsigned long long* const curB[1] = {next_len};
        unsigned long long posB[1];
        block_reserve<1>(scratchB[it & 1], cntB, curB, posB);
        unsigned long long o = posB[0];
        if (alive) next[o++] = i;
        for (unsigned b = 0; b < sink.n_ok; b++) next[o++] = (uint32_t)(posA[1] + b);
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
            load_particle(P, B, i, p);
            alive = ev_cross_pre<true>(P, p, B.surf[i], T, C, n_copy);
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
        BankSink sink(P, B, slot0, n_slots, C);
        if (valid) {
            alive = ev_cross_post(P, p, alive, n_copy, sink);
            if (alive) {
                B.x[i] = p.x; B.y[i] = p.y; B.z[i] = p.z; B.u[i] = p.u; B.v[i] = p.v; B.w[i] = p.w; B.t[i] = p.t;
                B.wgt[i] = p.wgt; B.rng[i] = p.rng; B.cell[i] = p.cell;
                if (P.track_time) B.told[i] = p.told;
            }
        }
        const unsigned cntB[1] = {(alive ? 1u : 0u) + sink.n_ok};
        unsigned long long* const curB[1] = {next_len};
        unsigned long long posB[1];
        block_reserve<1>(scratchB[it & 1], cntB, curB, posB);
        unsigned long long o = posB[0];
        if (alive) next[o++] = i;
        for (unsigned b = 0; b < sink.n_ok; b++) next[o++] = (uint32_t)(slot0 + b);
    }
    for (int d = 16; d; d >>= 1) crossings += __shfl_xor_sync(FULL, crossings, d);
    if (lane_id() == 0 && crossings) atomicAdd(&C->n_crossings, (unsigned long long)crossings);
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
        load_particle(P, B, i, p);
        bool alive = true;
        while (alive) {
            MacroXS X = {0, 0, 0, 0, 0};
            XSDetail D;
            int uidx = -1, S;
            if (ev_lookup(P, p, X, uidx, D)) lookups++;
            const bool to_cross = ev_flight<true>(P, p, X, uidx, H, T, C, S);
            tracks++;
            unsigned n_new = 0;
            unsigned long long slot0 = 0;
            if (to_cross) {
                unsigned n_copy;
                alive = ev_cross_pre<true>(P, p, S, T, C, n_copy);
                crossings++;
                if (n_copy) slot0 = atomicAdd(&C->slot_cursor, (unsigned long long)n_copy);
                BankSink sink(P, B, slot0, n_slots, C);
                alive = ev_cross_post(P, p, alive, n_copy, sink);
                n_new = sink.n_ok;
            } else {
                CollideCtx c;
                alive = ev_collide_pre<true>(P, p, X, uidx, D, T, C, k_eff, c);
                if (alive) {
                    collisions++;
                    unsigned long long site0 = 0;
                    if (c.n_sites) site0 = atomicAdd(&C->site_cursor, (unsigned long long)c.n_sites);
                    if (c.n_second) slot0 = atomicAdd(&C->slot_cursor, (unsigned long long)c.n_second);
                    BankSink sink(P, B, slot0, n_slots, C);
                    if (c.n_sites | c.n_second) ev_collide_bank(P, p, c, H, C, reqs, site_cap, site0, sink);
                    n_new = sink.n_ok;
                    alive = ev_collide_scatter<true>(P, p, X, uidx, D, c, H);
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
    const int64_t tally = blockIdx.x;  // gridDim.x holds any number of tallies (gridDim.y stops at 65535)
    const int chunk = blockIdx.y;
    double* row = acc + tally * stride;
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
        partial[(tally * n_chunks + chunk) * 2 + 0] = s;
        partial[(tally * n_chunks + chunk) * 2 + 1] = q;
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
k_gather_sites(const SourceBankView V, uint64_t q0, uint64_t n, Site* out)
{
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < n) store_site(out + q, source_bank_site(V, q0 + q));
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
        const UnionPos up = union_pos(M, e);
        MacroXS X;
        macro_xs(P, M, up, e, X);
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
    const UnionPos up = union_pos(M, E[q]);
    const int u = up.u;
    MacroXS X;
    macro_xs(P, M, up, E[q], X);
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
    scatter_sample(P.nuclides[nuclide], u, v, w, E, speed, rng);
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

// a / b three ways: through one shared refined reciprocal (mcb_div_shared), as the zero-numerator shortcut where it applies
// (mcb_div_zero_ok), and as the compiler's IEEE division: the first two must equal the third bit for bit
__global__ void k_division(const double* a, const double* b, int64_t n, double* out_shared, double* out_plain)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    const double r = mcb_rcp_shared(b[q]);
    out_shared[q] = (a[q] == 0.0 && b[q] > 0.0) ? mcb_div_zero_ok(a[q], b[q]) : mcb_div_shared(a[q], b[q], r);
    out_plain[q] = a[q] / b[q];
}

inline unsigned blocks_for(uint64_t n, unsigned bs = 256) { return (unsigned)((n + bs - 1) / bs); }

}  // namespace

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
namespace mcbk {

thread_local uint64_t g_launches = 0;  // also bumped by mcb_walk.cu
uint64_t launch_count() { return g_launches; }
#define MCB_LAUNCHED(k) (g_launches += (k))

static int g_n_sm = 148;  // multiprocessors of the device in use (cudaDeviceProp::multiProcessorCount)
void set_device_sms(int n) { if (n > 0) g_n_sm = n; }
static unsigned grid_for(uint64_t n_hint)
{
    // persistent tile loops: enough blocks to fill the machine a few times over, never more than the work
    const uint64_t need = (n_hint + BLOCK - 1) / BLOCK;
    return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(need, (uint64_t)g_n_sm * 16ull * (256 / BLOCK)));
}
static void sort_draws(cudaStream_t st, const SortScratch* sort, uint32_t count, uint64_t n_bank)
{
    int bits = 1;
    while (bits < 64 && (n_bank >> bits)) bits++;
    size_t tb = sort->temp_bytes;
    if (sort->key32)
        cub::DeviceRadixSort::SortPairs(sort->temp, tb, reinterpret_cast<const uint32_t*>(sort->key_in), reinterpret_cast<uint32_t*>(sort->key_out),
                                        sort->val_in, sort->val_out, (int)count, 0, std::min(bits, 32), st);
    else
        cub::DeviceRadixSort::SortPairs(sort->temp, tb, sort->key_in, sort->key_out, sort->val_in, sort->val_out, (int)count, 0, bits, st);
    MCB_LAUNCHED((std::min(bits, sort->key32 ? 32 : 64) + 7) / 8 + 2);  // the sort's histogram / onesweep passes
}
void source(cudaStream_t st, const DevProblem& P, const Bank& B, uint32_t* active, int32_t first_hist, uint32_t count,
            uint64_t nps0, const SourceBankView& V, Counters* C, const SortScratch* sort)
{
    if (sort && V.n) {
        // draws -> sorted by site index -> the bank is read in ascending order
        k_pick<<<std::max(1u, blocks_for(count, BLOCK)), BLOCK, 0, st>>>(P.seed0, nps0, first_hist, count, V.n, sort->rot, sort->key_in, sort->val_in, sort->rng_after, sort->key32);
        sort_draws(st, sort, count, V.n);
        k_source<<<std::max(1u, blocks_for(count, BLOCK)), BLOCK, 0, st>>>(P, B, active, first_hist, count, nps0, V, C, sort->key_out,
                                                                          sort->val_out, sort->rng_after, sort->rot, 0u, sort->key32);
        MCB_LAUNCHED(2);
        return;
    }
    k_source<<<std::max(1u, blocks_for(count, BLOCK)), BLOCK, 0, st>>>(P, B, active, first_hist, count, nps0, V, C, nullptr, nullptr, nullptr, 0ull, 0u, 0);
    MCB_LAUNCHED(1);
}
// the two halves of the sorted path on their own, for a bank that arrives from the host in chunks: draws + sort
// first, then one k_source launch per chunk of the sweep
void pick_sort(cudaStream_t st, const DevProblem& P, int32_t first_hist, uint32_t count, uint64_t nps0, uint64_t n_bank,
               const SortScratch* sort)
{
    k_pick<<<std::max(1u, blocks_for(count, BLOCK)), BLOCK, 0, st>>>(P.seed0, nps0, first_hist, count, n_bank, sort->rot, sort->key_in, sort->val_in, sort->rng_after, sort->key32);
    sort_draws(st, sort, count, n_bank);
    MCB_LAUNCHED(1);
}
__global__ void k_chunk_bounds(const unsigned long long* __restrict__ sorted_key, uint32_t n, const unsigned long long* __restrict__ lo,
                               int n_lo, uint32_t* pos, int key32)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_lo) return;
    uint32_t a = 0, b = n;  // first position whose key is >= lo[c]
    while (a < b) {
        const uint32_t m = (a + b) >> 1;
        const unsigned long long km = key32 ? (unsigned long long)reinterpret_cast<const uint32_t*>(sorted_key)[m] : sorted_key[m];
        if (km < lo[c]) a = m + 1; else b = m;
    }
    pos[c] = a;
}
__global__ void k_publish(unsigned long long* dst, unsigned long long value) { __threadfence(); *dst = value; }
void publish(cudaStream_t st, unsigned long long* dst, unsigned long long value)
{
    k_publish<<<1, 1, 0, st>>>(dst, value);
    MCB_LAUNCHED(1);
}
void source_sweep(cudaStream_t st, const Bank& B, int32_t first_hist, uint32_t q0, uint32_t count, const SourceBankView& V, const SortScratch* sort)
{
    if (!count) return;
    constexpr int ITEMS = 4;
    k_source_sweep<ITEMS><<<blocks_for(count, 256 * ITEMS), 256, 0, st>>>(B, first_hist, q0, q0 + count, V, sort->key_out, sort->val_out, sort->rng_after,
                                                                        sort->rot, sort->key32);
    MCB_LAUNCHED(1);
}
void preload_side_kernels(cudaStream_t st, unsigned long long* scratch)
{
    // lazy loading: a kernel's first launch waits for every running kernel to end, so the kernels that are meant to run
    // BESIDE the persistent walk kernel are launched once (on nothing) when the context is made
    SourceBankView V;
    memset(&V, 0, sizeof(V));
    Bank B;
    memset(&B, 0, sizeof(B));
    k_source_sweep<4><<<1, 256, 0, st>>>(B, 0, 0u, 0u, V, nullptr, nullptr, nullptr, 0ull, 0);
    k_publish<<<1, 1, 0, st>>>(scratch, 0ull);
}
void chunk_bounds(cudaStream_t st, const SortScratch* sort, uint32_t n, const unsigned long long* lo, int n_lo, uint32_t* pos)
{
    k_chunk_bounds<<<1, 64, 0, st>>>(sort->key_out, n, lo, n_lo, pos, sort->key32);
    MCB_LAUNCHED(1);
}
void source_sorted_range(cudaStream_t st, const DevProblem& P, const Bank& B, uint32_t* active, int32_t first_hist, uint32_t q0,
                         uint32_t count, uint64_t nps0, const SourceBankView& V, Counters* C, const SortScratch* sort)
{
    if (!count) return;
    k_source<<<blocks_for(count, BLOCK), BLOCK, 0, st>>>(P, B, active, first_hist, count, nps0, V, C, sort->key_out, sort->val_out,
                                                         sort->rng_after, sort->rot, q0, sort->key32);
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
    if (n) { k_reduce_k<<<min(blocks_for(n), (unsigned)g_n_sm * 8u), 256, 0, st>>>(kC, kTL, n, C); MCB_LAUNCHED(1); }
}
void entropy_history(cudaStream_t st, const DevProblem& P, const Site* bank, const uint32_t* offset,
                     const int32_t* nsite, uint32_t n_hist, Counters* C)
{
    if (n_hist) { k_entropy_history<<<min(blocks_for(n_hist), (unsigned)g_n_sm * 8u), 256, 0, st>>>(P, bank, offset, nsite, n_hist, C); MCB_LAUNCHED(1); }
}
void entropy_histogram(cudaStream_t st, const DevProblem& P, const Site* bank, uint64_t n, unsigned long long* bins)
{
    if (n) { k_entropy_histogram<<<min(blocks_for(n), (unsigned)g_n_sm * 8u), 256, 0, st>>>(P, bank, n, bins); MCB_LAUNCHED(1); }
}
int tally_chunks(uint32_t n_hist) { return (int)((n_hist + TALLY_CHUNK - 1) / TALLY_CHUNK); }
void tally_reduce(cudaStream_t st, double* acc, int64_t stride, uint32_t n_hist, int64_t n_tallies, double* partial,
                  double* sum, double* squared)
{
    if (!n_hist || !n_tallies) return;
    const int nc = tally_chunks(n_hist);
    k_tally_partial<<<dim3((unsigned)n_tallies, nc), 256, 0, st>>>(acc, stride, n_hist, partial, nc);
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
void gather_sites(cudaStream_t st, const SourceBankView& V, uint64_t q0, uint64_t n, Site* out)
{
    if (n) { k_gather_sites<<<blocks_for(n), 256, 0, st>>>(V, q0, n, out); MCB_LAUNCHED(1); }
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
void division(cudaStream_t st, const double* a, const double* b, int64_t n, double* out_shared, double* out_plain)
{
    if (n) { k_division<<<blocks_for(n), 256, 0, st>>>(a, b, n, out_shared, out_plain); MCB_LAUNCHED(1); }
}
void watt(cudaStream_t st, const DevProblem& P, int nuclide, const uint64_t* nps, const double* E, int64_t n, double* out)
{
    if (n) { k_watt<<<blocks_for(n), 256, 0, st>>>(P, nuclide, nps, E, n, out); MCB_LAUNCHED(1); }
}

}  // namespace mcbk
