// mcb_walk.cu — the transport loop of one generation: k_walk, an event-based particle queue held in shared memory.
//
// Reference: the body of the sample loop of Simulator::start() (handler.cpp:19-37): every source particle and the
// secondaries of its history are followed through random_walk (general.cpp:177-211) until the history's particle
// bank (Pbank) is empty, then the history is closed out (end_history).
//
// Design (B200: 148 SMs, 227 KB of shared memory per SM, FP64 scalar code bound by issue slots and latency):
//  * Every thread block owns 160 particle SLOTS in shared memory (SoA, 16-byte pairs).  A slot is the home of one
//    history from the moment it is drawn from the source bank until it ends: the particle in flight, its
//    EstimatorK scores, the macroscopic cross sections and per-nuclide partial sums of its last lookup.
//  * Particles are SORTED BY NEXT EVENT in two block-level queues of slot numbers: "collide" and "cross".
//    A warp runs the common part (xs lookup + flight) on the 32 particles it holds, parks them in their slots,
//    appends them to the queue of their next event and takes 32 slots of ONE kind back out, so the collision and
//    crossing code runs on full warps (the history-per-lane formulation of round 1 ran collisions on 23 and
//    crossings on 8 of 32 lanes).  With 128 threads and 160 slots at least 64 slots are queued whenever a warp
//    asks, so one queue always holds a full warp.  Warps stay autonomous: the only synchronisation is a
//    shared-memory lock around the few instructions that move queue heads and tails.
//  * A lane whose history has ended draws the next source particle from the bank (chunks of a device-side head
//    counter), so lanes stay busy until the bank runs dry.
//  * Same-history secondaries (fixed-source fission neutrons, split copies) go to the history's own LIFO stack in
//    global memory — the reference's Pbank — and are followed by whichever lane holds the history, so a history is
//    always followed by one lane at a time: its k scores stay on registers, its tallies accumulate in a private
//    table and there is no pass structure and no host round trip, however long the fission chain.
//  * Fission sites leave the kernel as requests (one warp-aggregated cursor reservation per batch) and are sampled
//    and put into canonical order by k_bank_sample_order.
//
// The file is compiled twice (two "flavours"), one kernel family each:
//  * flavour 0 (this file): cycles that score no tallies: blocks of 128 threads, 4 per SM, warps free-running;
//  * flavour 1 (mcb_walk_tally.cu includes this file): scoring cycles.  The estimator code makes the loop body several
//    times larger than the instruction caches, and warps that run free each stream it from L2 on their own (ncu:
//    stall_no_inst 30 %, sm__icc_request_hit_rate 59 %).  One block of 512 threads per SM whose warps meet at a
//    barrier once per track run the same stretch of code at the same time and share the fetched lines: +47 % on the
//    TRMM tally set (profiles/README.md, round 2e).  The barrier does not pay without estimators (measured).
#ifndef MCB_WALK_FLAVOUR
#define MCB_WALK_FLAVOUR 0
#endif
#include "mcb_events.cuh"

#include <algorithm>

using namespace mcbe;

namespace {

#ifndef MCB_WALK_MINB
#define MCB_WALK_MINB 4  // measured (tools/sweep.sh): 4 blocks of 128 threads per SM (128 registers) beat 3, 5 and 6 for both forms
#endif
constexpr int WALK_RES = 32;                  // slots beyond one per thread: what makes a full batch always available
constexpr int WALK_SLOTS = BLOCK + WALK_RES;  // 160
// secondary stacks: chunks of 32 records.  Every slot owns one chunk for good (a history rarely has more than a few
// particles waiting); a history that needs more borrows chunks from its block's pool and returns them when it ends
constexpr int STACK_CHUNK = 32;
constexpr int WALK_EXTRA = 2 * BLOCK;         // spare chunks per block
constexpr int STACK_MAXCH = 32;               // chunks per history at most (1024 particles waiting)
constexpr int pow2_at_least(int n) { int p = 1; while (p < n) p <<= 1; return p; }
constexpr int WALK_QCAP = pow2_at_least(WALK_SLOTS);  // ring capacity of a queue
static_assert(WALK_QCAP >= WALK_SLOTS && (WALK_QCAP & (WALK_QCAP - 1)) == 0, "queue ring");

// slot state, pairs of doubles laid out [pair][slot] (a warp's 128-bit accesses fall on distinct banks for slots that
// differ modulo 8)
enum {
    SP_XY = 0,    // x, y
    SP_ZU,        // z, u
    SP_VW,        // v, w
    SP_ES,        // E, speed
    SP_WT,        // weight, time
    SP_RNG,       // stream state | (cell, history)
    SP_K,         // k_C, k_TL of the history so far
    SP_IDS,       // (sites banked by the history, depth of its secondary stack) | (surface ahead, union-grid index)
    SP_XT,        // SigmaT, nuSigmaF
    SP_XS,        // SigmaS, SigmaC
    SP_XF,        // SigmaF, energy_old
    SP_FIXED      // scoring instances add: time_old | (tally-table entries in use, -); then the per-nuclide detail
};

struct WalkQ {
    unsigned lock;
    unsigned headC, tailC, headX, tailX;
    unsigned warps_done;
    unsigned alloc_lock;
    int n_free;                               // spare stack chunks in free_list
    unsigned short qC[WALK_QCAP], qX[WALK_QCAP];
    unsigned short free_list[WALK_EXTRA];
};

// slot state lives in shared memory, or (MCB_WALK_GLOBAL_STATE) in a per-block global array that stays in L2, read and
// written around L1 so that L1 keeps serving the cross-section gathers
#ifdef MCB_WALK_GLOBAL_STATE
__device__ __forceinline__ void st_pair(double2* p, double2 v) { __stcg(p, v); }
__device__ __forceinline__ double2 ld_pair(const double2* p) { return __ldcg(p); }
#else
__device__ __forceinline__ void st_pair(double2* p, double2 v) { *p = v; }
__device__ __forceinline__ double2 ld_pair(const double2* p) { return *p; }
#endif
__device__ __forceinline__ double pack2i(int lo, int hi) { return __hiloint2double(hi, lo); }
__device__ __forceinline__ int unpack_lo(double d) { return __double2loint(d); }
__device__ __forceinline__ int unpack_hi(double d) { return __double2hiint(d); }

// per-nuclide partial sums of the particle's last lookup, kept in its slot: (Sigma_s, nuSigma_f) after nuclides 0..n as
// one pair per nuclide, then beta_n two to a pair
struct SlotDetail {
    static constexpr bool present = true;
    double2* base;  // pair SP_DET of this slot
    int nn;         // nuclides per material at most (problem-wide)
    __device__ __forceinline__ void set(int n, double s, double nf, double be) const
    {
        st_pair(base + n * WALK_SLOTS, make_double2(s, nf));
#ifdef MCB_WALK_GLOBAL_STATE
        __stcg(reinterpret_cast<double*>(base + (nn + (n >> 1)) * WALK_SLOTS) + (n & 1), be);
#else
        reinterpret_cast<double*>(base + (nn + (n >> 1)) * WALK_SLOTS)[n & 1] = be;
#endif
    }
    __device__ __forceinline__ double cum_s(int n) const { return ld_pair(base + n * WALK_SLOTS).x; }
    __device__ __forceinline__ double cum_nf(int n) const { return ld_pair(base + n * WALK_SLOTS).y; }
    __device__ __forceinline__ double beta(int n) const
    {
        const double2 v = ld_pair(base + (nn + (n >> 1)) * WALK_SLOTS);
        return (n & 1) ? v.y : v.x;
    }
};

// the history's LIFO stack of secondaries (the reference's Pbank, handler.cpp:20-29): record i lives in chunk i / 32 of
// the history (chunk 0 = the slot's own, the others listed in its row of the chunk table)
struct StackSink {
    StackRec* blk;              // this block's chunks
    const unsigned short* tab;  // chunk table row of the history
    int own;                    // the slot's own chunk
    int& sp;                    // records on the stack
    int nch;                    // chunks the history holds
    Counters* C;
    __device__ __forceinline__ StackRec* rec(int i) const
    {
        const int k = i / STACK_CHUNK;
        const int id = k == 0 ? own : (int)__ldcg(tab + k);
        return blk + (size_t)id * STACK_CHUNK + (i % STACK_CHUNK);
    }
    __device__ __forceinline__ void push(const Particle& q)
    {
        if (sp >= nch * STACK_CHUNK) { C->overflow_stack = 1; return; }
        double2* r = reinterpret_cast<double2*>(rec(sp));
        r[0] = make_double2(q.x, q.y); r[1] = make_double2(q.z, q.u); r[2] = make_double2(q.v, q.w);
        r[3] = make_double2(q.E, q.speed); r[4] = make_double2(q.wgt, q.t);
        r[5] = make_double2(q.Eold, __longlong_as_double((long long)q.rng));
        r[6] = make_double2(pack2i(q.cell, q.hist), pack2i(q.drow, q.tdmc));
        sp++;
    }
    __device__ __forceinline__ void pop(Particle& p)
    {
        sp--;
        const double2* r = reinterpret_cast<const double2*>(rec(sp));
        const double2 a = r[0], b = r[1], c = r[2], d = r[3], e = r[4], f = r[5], g = r[6];
        p.x = a.x; p.y = a.y; p.z = b.x; p.u = b.y; p.v = c.x; p.w = c.y; p.E = d.x; p.speed = d.y; p.wgt = e.x; p.t = e.y;
        p.Eold = f.x; p.rng = (uint64_t)__double_as_longlong(f.y); p.cell = unpack_lo(g.x); p.tdmc = unpack_hi(g.y);
        p.told = p.t;
    }
};

// ---- work sharing: secondaries handed to other lanes through a global ring (DonationQueue)
constexpr int WAIT_LIMIT = 1 << 19;  // bounded waits: a protocol error becomes an error code, never a hung GPU
__device__ __forceinline__ bool donation_push(DonationQueue* D, const Particle& q, Counters* C)
{
    if (atomicAdd(&D->count, 1) >= (int)D->cap_mask) { atomicSub(&D->count, 1); return false; }
    const unsigned long long pos = atomicAdd(&D->tail, 1ull);
    const uint32_t cell = (uint32_t)pos & D->cap_mask;
    volatile unsigned long long* seq = D->seq + cell;
    int spins = 0;
    while (*seq != pos) { if (++spins > WAIT_LIMIT) { C->hang = 1; return true; } __nanosleep(64); }
    double2* r = reinterpret_cast<double2*>(D->recs + cell);
    __stcg(r + 0, make_double2(q.x, q.y)); __stcg(r + 1, make_double2(q.z, q.u)); __stcg(r + 2, make_double2(q.v, q.w));
    __stcg(r + 3, make_double2(q.E, q.speed)); __stcg(r + 4, make_double2(q.wgt, q.t));
    __stcg(r + 5, make_double2(q.Eold, __longlong_as_double((long long)q.rng)));
    __stcg(r + 6, make_double2(pack2i(q.cell, q.hist), pack2i(q.drow, q.tdmc)));
    __threadfence();
    *seq = pos + 1ull;
    atomicAdd(&D->avail, 1);
    return true;
}
__device__ __forceinline__ bool donation_pop(DonationQueue* D, Particle& p, Counters* C)
{
    if (atomicSub(&D->avail, 1) <= 0) { atomicAdd(&D->avail, 1); return false; }
    const unsigned long long pos = atomicAdd(&D->head, 1ull);
    const uint32_t cell = (uint32_t)pos & D->cap_mask;
    volatile unsigned long long* seq = D->seq + cell;
    int spins = 0;
    while (*seq != pos + 1ull) { if (++spins > WAIT_LIMIT) { C->hang = 2; return false; } __nanosleep(64); }
    __threadfence();
    const double2* r = reinterpret_cast<const double2*>(D->recs + cell);
    const double2 a = __ldcg(r + 0), b = __ldcg(r + 1), c = __ldcg(r + 2), d = __ldcg(r + 3), e = __ldcg(r + 4), f = __ldcg(r + 5), g = __ldcg(r + 6);
    p.x = a.x; p.y = a.y; p.z = b.x; p.u = b.y; p.v = c.x; p.w = c.y; p.E = d.x; p.speed = d.y; p.wgt = e.x; p.t = e.y;
    p.Eold = f.x; p.rng = (uint64_t)__double_as_longlong(f.y); p.cell = unpack_lo(g.x); p.hist = unpack_hi(g.x); p.drow = unpack_lo(g.y); p.tdmc = unpack_hi(g.y);
    p.told = p.t;
    __threadfence();
    *seq = pos + (unsigned long long)D->cap_mask + 1ull;
    atomicSub(&D->count, 1);
    return true;
}
// chunk pool of the block: taken and returned by single lanes under a lock of its own (rare: only histories with more
// than 32 particles waiting get here); callers serialise the lanes of a warp
__device__ __forceinline__ int chunks_take(WalkQ& Q, unsigned short* tab, int nch, int want)
{
    while (atomicCAS(&Q.alloc_lock, 0u, 1u) != 0u) __nanosleep(64);
    __threadfence_block();
    int got = 0;
    while (got < want && Q.n_free > 0 && nch + got < STACK_MAXCH) { __stcg(tab + nch + got, Q.free_list[--Q.n_free]); got++; }
    __threadfence_block();
    atomicExch(&Q.alloc_lock, 0u);
    return got;
}
__device__ __forceinline__ void chunks_give(WalkQ& Q, const unsigned short* tab, int nch)
{
    while (atomicCAS(&Q.alloc_lock, 0u, 1u) != 0u) __nanosleep(64);
    __threadfence_block();
    for (int k = 1; k < nch; k++) Q.free_list[Q.n_free++] = __ldcg(tab + k);
    __threadfence_block();
    atomicExch(&Q.alloc_lock, 0u);
}

// particle_comb (population_control.cpp:55-84) on the history's stack, by the lane that follows the history (rare and
// short: banks of a few dozen particles).  The one draw comes from the stream of the particle whose walk just ended.  As
// in the reference: a particle takes at most one tooth; the new bank starts as `teeth` copies of the first particle, and
// teeth the loop does not reach keep that copy with its original weight (their streams are set (q + 1) * 2^40 draws on,
// so that the copies do not repeat each other); a tooth past the end is dropped.
__device__ __noinline__ static void comb_stack(StackSink sink, int teeth, uint64_t& rng_done, int* sp_out)
{
    const int n = sink.sp;
    double W = 0.0;
    for (int i = 0; i < n; i++) W += reinterpret_cast<const double2*>(sink.rec(i))[4].x;
    const double w_avg = W / (double)teeth;
    double tooth = mcb_urand(rng_done) * w_avg, sum = 0.0;
    auto copy = [&](int from, int to) {
        const double2* a = reinterpret_cast<const double2*>(sink.rec(from));
        double2* b = reinterpret_cast<double2*>(sink.rec(to));
        for (int k = 0; k < 7; k++) b[k] = a[k];
    };
    for (int q = 0; q < teeth; q++) copy(0, n + q);
    int j = 0;
    for (int i = 0; i < n; i++) {
        sum += reinterpret_cast<const double2*>(sink.rec(i))[4].x;
        if (sum > tooth) {
            if (j < teeth) { copy(i, n + j); reinterpret_cast<double2*>(sink.rec(n + j))[4].x = w_avg; }
            tooth += w_avg; j++;
        }
    }
    const uint64_t rng0 = (uint64_t)__double_as_longlong(reinterpret_cast<const double2*>(sink.rec(0))[5].y);
    for (int q = j; q < teeth; q++)
        reinterpret_cast<double2*>(sink.rec(n + q))[5].y = __longlong_as_double((long long)mcb_rn_skip(rng0, ((uint64_t)(q + 1)) << 40));
    for (int q = 0; q < teeth; q++) copy(n + q, q);
    *sp_out = teeth;
}

__device__ __forceinline__ void lock_acquire(WalkQ& Q, unsigned lane)
{
    if (lane == 0) {
        while (atomicCAS(&Q.lock, 0u, 1u) != 0u) __nanosleep(32);
    }
    __syncwarp();
    __threadfence_block();
}
__device__ __forceinline__ void lock_release(WalkQ& Q, unsigned lane)
{
    __threadfence_block();
    __syncwarp();
    if (lane == 0) atomicExch(&Q.lock, 0u);
}

// a unit of a shared history ends: its private table goes into the history's dense row; the last unit out turns the row
// into sum / squared.  Warp-cooperative like the flush.
__device__ __forceinline__ void merge_shared_tallies(const DevProblem& P, const TallyAcc& T, int row, int n_touched, int drow, double* s_sum, double* s_sq, unsigned lane)
{
    double* dense = T.dense + (size_t)drow * T.n_tallies;
    for (int i = (int)lane; i < n_touched; i += 32) {
        uint32_t t;
        double v;
        if (tally_take(P, T, row, i, t, v)) atomicAdd(dense + t, v);
    }
    __threadfence();
    __syncwarp();
    int left = 0;
    if (lane == 0) left = atomicSub(T.dense_pending + drow, 1) - 1;
    left = __shfl_sync(FULL, left, 0);
    if (left != 0) return;
    __threadfence();
    for (int t = (int)lane; t < T.n_tallies; t += 32) {
        const double v = __ldcg(dense + t);
        if (v != 0.0) {
            __stcg(dense + t, 0.0);
            if (s_sum) { atomicAdd(s_sum + t, v); atomicAdd(s_sq + t, v * v); }
            else { atomicAdd(T.sum + t, v); atomicAdd(T.squared + t, v * v); }
        }
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) atomicExch(T.dense_pending + drow, -1);  // the row is clean again: free for another history
}

// SourceBank::get_source (Source.cpp:42-46) in the lane that is about to follow the history, with site index floor(xi N):
// history h draws from the stream of nps = cycle * Nsample + h (RN_init_particle, Random.cpp:196-204).
// (inlined: an out-of-line callee that writes the particle through a reference pins the whole record to local memory:
// measured 5.68 -> 6.52 ms per generation.)
__device__ __forceinline__ void fused_source(const mcbk::WalkSource& SRC, uint32_t j, uint64_t chunk_seed, unsigned long long chunk_begin, Particle& p)
{
    unsigned long long site;
    if (SRC.sorted_key) {  // the draws were made and sorted by site index beforehand (k_pick)
        const uint32_t sv = __ldg(SRC.sorted_val + j);
        p.hist = SRC.first_hist + (int32_t)sv;
        p.rng = __ldg(SRC.rng_after + sv);
        site = (SRC.key32 ? (unsigned long long)__ldg(reinterpret_cast<const uint32_t*>(SRC.sorted_key) + j)
                          : __ldg(reinterpret_cast<const unsigned long long*>(SRC.sorted_key) + j)) + SRC.rot;
        if (site >= SRC.V.n) site -= SRC.V.n;
    } else {
        p.hist = SRC.first_hist + (int32_t)j;
        p.rng = mcb_rn_history_seed_from(chunk_seed, (uint32_t)(j - chunk_begin));
        const double xi = mcb_urand(p.rng);
        site = (unsigned long long)(xi * (double)SRC.V.n);
        if (site >= SRC.V.n) site = SRC.V.n - 1;
    }
    const Site sx = source_bank_site(SRC.V, site);  // local HBM, or a peer's HBM over NVLink
    p.x = sx.x; p.y = sx.y; p.z = sx.z; p.E = sx.E; p.t = sx.t; p.cell = sx.cell;
    source_bank_direction(SRC.V, site, sx, p.u, p.v, p.w);
    p.speed = mcb_speed_of_energy(p.E); p.wgt = 1.0;
}

// EXCH = true: the event-sorted form described above.  EXCH = false: every lane keeps its history from the source bank to
// its end and runs its own next event (collide and cross lanes of a warp diverge): no slot traffic and no queue, every
// warp streams through one loop body.  Which one wins is a matter of the instruction cache: the loop is ~70 KB of code
// against a 32 KB per-SM instruction cache, and warps scattered over three code regions (EXCH) fetch from the GPC-level
// cache at its limit (ncu: gcc instruction requests 98 % of peak, sm__icc hit rate 61 %), see DESIGN.md.
template <bool TALLY, bool SHARED, bool EXCH>
__global__ void __launch_bounds__(BLOCK, MCB_WALK_MINB)
k_walk(const DevProblem P, const Bank B, unsigned long long begin, unsigned long long end, uint32_t chunk, Counters* C, HistoryAcc H,
       TallyAcc T, SiteReq* reqs, uint64_t site_cap, double k_eff, mcbk::WalkRes R, const __grid_constant__ mcbk::WalkSource SRC)
{
    if (R.sm_limit > 0) {  // whole SMs are left to the kernel that runs beside this one (kernels of different shared-memory
        unsigned smid;     // configurations do not share an SM, so free block slots would not do)
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        if ((int)smid >= R.sm_limit) return;
    }
    extern __shared__ double2 w_smem[];
    constexpr int SP_DET = !EXCH ? 0 : (TALLY ? SP_FIXED + 1 : SP_FIXED);  // without the exchange a slot holds the detail only
#ifdef MCB_WALK_GLOBAL_STATE
    double2* const st = R.gstate + (size_t)blockIdx.x * R.n_pairs * WALK_SLOTS;
    WalkQ& Q = *reinterpret_cast<WalkQ*>(w_smem);
#else
    double2* const st = w_smem;
    WalkQ& Q = *reinterpret_cast<WalkQ*>(st + R.n_pairs * WALK_SLOTS);
#endif
    double* const s_sum = (TALLY && R.priv_tallies) ? reinterpret_cast<double*>(&Q + 1) : nullptr;
    double* const s_sq = s_sum ? s_sum + R.priv_tallies : nullptr;
    if (threadIdx.x == 0) { Q.lock = 0; Q.headC = Q.tailC = Q.headX = Q.tailX = 0; Q.warps_done = 0; Q.alloc_lock = 0; Q.n_free = SHARED ? WALK_EXTRA : 0; }
    if (SHARED) for (int i = threadIdx.x; i < WALK_EXTRA; i += BLOCK) Q.free_list[i] = (unsigned short)(WALK_SLOTS + i);
    if (s_sum) for (int i = threadIdx.x; i < 2 * R.priv_tallies; i += BLOCK) s_sum[i] = 0.0;
    __syncthreads();  // the only block-wide barrier: from here on the warps run on their own

    const unsigned lane = lane_id();
    const unsigned lt_mask = (1u << lane) - 1u;
    const int ctx_base = blockIdx.x * WALK_SLOTS;
    unsigned tracks = 0, collisions = 0, crossings = 0, lookups = 0;
    int my_slot = threadIdx.x;            // the slot this lane parks its particle in (it moves with every claim)
    bool have = false, exhausted = false;
    bool own = true;                      // event-sorted form: this lane holds a slot to park a particle in
    bool second_batch = EXCH && warp_id() == 0;   // warp 0 starts two batches: the block's 32 extra slots
    Particle p;
    HistLocal L = {0.0, 0.0, 0};
    int sp = 0, nch = 1;                  // records on the history's secondary stack, chunks it holds
    unsigned long long chunk_next = 0, chunk_end = 0;  // warp-uniform: this warp's private range of bank positions
    int idle_spins = 0;
    unsigned long long chunk_begin = 0;
    uint64_t chunk_seed = 0;
    for (;;) {
#if defined(MCB_WALK_SYNC) && MCB_WALK_SYNC >= 2
        if (!EXCH) __syncthreads();
#endif
        // ---- lanes without a history draw the next source particles
        unsigned idle = __ballot_sync(FULL, !have);
        while (idle && !exhausted) {
            if (chunk_next == chunk_end) {
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(&C->walk_head, (unsigned long long)chunk);
                base = __shfl_sync(FULL, base, 0);
                const unsigned long long b = begin + base;
                chunk_next = b < end ? b : end;
                chunk_end = b + chunk < end ? b + chunk : end;
                if (chunk_next == chunk_end) { exhausted = true; break; }
                if (SRC.ready) {  // the bank is being filled beside this kernel, in ascending order: wait for this chunk
                    // (one lane polls, with a growing pause: thousands of warps hammering one address would starve the
                    // very store they are waiting for)
                    int spins = 0;
                    unsigned pause = 250;
                    for (;;) {
                        unsigned long long got = 0;
                        if (lane == 0) got = *(volatile const unsigned long long*)SRC.ready;
                        got = __shfl_sync(FULL, got, 0);
                        if (got >= chunk_end) break;
                        if (++spins > WAIT_LIMIT) { C->hang = 4; break; }
                        __nanosleep(pause);
                        if (pause < 8000) pause *= 2;
                    }
                    __threadfence();
                }
                if (SRC.fused && !SRC.sorted_key) {  // stream of the chunk's first history; the others are a short skip away
                    chunk_begin = chunk_next;
                    chunk_seed = mcb_rn_history_seed(SRC.seed0, SRC.nps0 + (uint64_t)((long long)SRC.first_hist + (long long)chunk_next));
                }
            }
            const unsigned take = min((unsigned)__popc(idle), (unsigned)(chunk_end - chunk_next));
            const unsigned rank = __popc(idle & lt_mask);
            if (!have && rank < take) {
                const uint32_t j = (uint32_t)(chunk_next + rank);
                if (SRC.fused) {
                    fused_source(SRC, j, chunk_seed, chunk_begin, p);
                } else {
                    p.cell = B.cell[j]; p.hist = B.hist[j];
                    p.x = B.x[j]; p.y = B.y[j]; p.z = B.z[j]; p.u = B.u[j]; p.v = B.v[j]; p.w = B.w[j];
                    p.E = B.E[j]; p.speed = B.speed[j]; p.wgt = B.wgt[j]; p.t = B.t[j]; p.rng = B.rng[j];
                }
                p.Eold = p.E;  // the reference leaves energy_old uninitialised at birth; defined as E here
                p.told = p.t;
                p.n_touched = 0; p.drow = -1; p.tdmc = 0;
                L.kC = 0.0; L.kTL = 0.0; L.nsite = 0;
                sp = 0; nch = 1;
                have = true;
            }
            chunk_next += take;
            if (SHARED && R.donq && lane == 0) atomicAdd((unsigned long long*)&C->live, (unsigned long long)take);
            idle = __ballot_sync(FULL, !have);
        }
        if (SHARED && R.donq && exhausted && idle) {  // the bank is dry: idle lanes take secondaries other lanes handed over
            int av = 0;
            if (lane == 0) av = __ldcg(&R.donq->avail);
            av = __shfl_sync(FULL, av, 0);
            if (av > 0 && !have && own && donation_pop(R.donq, p, C)) {
                p.n_touched = 0;
                L.kC = 0.0; L.kTL = 0.0; L.nsite = 0;
                sp = 0; nch = 1;
                have = true;
            }
            __syncwarp();
        }
        // ---- common part of every track: xs lookup and flight; the particle is parked in its slot
        bool to_cross = false, census = false;  // census: the flight ended at a census time (time-dependent mode), no event follows
        MacroXS X = {0, 0, 0, 0, 0};
        int uidx = -1, S = -1;
        if (have) {
            p.row = ctx_base + my_slot;
            const SlotDetail D = {st + SP_DET * WALK_SLOTS + my_slot, R.det_nn};
            if (ev_lookup(P, p, X, uidx, D)) lookups++;
            const int event = ev_flight<TALLY, SHARED>(P, p, X, uidx, H, T, C, S, &L);
            to_cross = event == 1;
            census = event == 2;
            tracks++;
        }
        if (EXCH && have) {
            double2* s = st + my_slot;
            st_pair(s + SP_XY * WALK_SLOTS, make_double2(p.x, p.y));
            st_pair(s + SP_ZU * WALK_SLOTS, make_double2(p.z, p.u));
            st_pair(s + SP_VW * WALK_SLOTS, make_double2(p.v, p.w));
            st_pair(s + SP_ES * WALK_SLOTS, make_double2(p.E, p.speed));
            st_pair(s + SP_WT * WALK_SLOTS, make_double2(p.wgt, p.t));
            st_pair(s + SP_RNG * WALK_SLOTS, make_double2(__longlong_as_double((long long)p.rng), pack2i(p.cell, p.hist)));
            st_pair(s + SP_K * WALK_SLOTS, make_double2(L.kC, L.kTL));
            st_pair(s + SP_IDS * WALK_SLOTS, make_double2(pack2i(L.nsite, sp | (nch << 16)), pack2i(S, uidx)));
            st_pair(s + SP_XT * WALK_SLOTS, make_double2(X.t, X.nf));
            st_pair(s + SP_XS * WALK_SLOTS, make_double2(X.s, X.c));
            st_pair(s + SP_XF * WALK_SLOTS, make_double2(X.f, p.Eold));
            if (TALLY) st_pair(s + SP_FIXED * WALK_SLOTS, make_double2(p.told, pack2i(p.n_touched, p.drow)));
        }
        // ---- event queues: append what this warp holds, take one batch of a kind back out
        int kind = 0;  // 1 collide, 2 cross
        if (!EXCH) kind = have ? (census ? 3 : to_cross ? 2 : 1) : 0;
        if (EXCH) lock_acquire(Q, lane);
        if (EXCH) {
            const unsigned mC = __ballot_sync(FULL, have && !to_cross), mX = __ballot_sync(FULL, have && to_cross);
            unsigned hC = Q.headC, tC = Q.tailC, hX = Q.headX, tX = Q.tailX;
            if (have) {
                own = false;  // the slot goes into a queue with its particle
                if (!to_cross) Q.qC[(tC + __popc(mC & lt_mask)) & (WALK_QCAP - 1)] = (unsigned short)my_slot;
                else Q.qX[(tX + __popc(mX & lt_mask)) & (WALK_QCAP - 1)] = (unsigned short)my_slot;
            }
            tC += __popc(mC); tX += __popc(mX);
            __syncwarp();
            unsigned nC = 0, nX = 0;
            if (!second_batch) {
                const unsigned aC = tC - hC, aX = tX - hX;
                if (aX >= 32u) nX = 32u;
                else if (aC >= 32u) nC = 32u;
                else if (aC >= aX) { nC = aC; nX = min(aX, 32u - nC); }
                else { nX = aX; nC = min(aC, 32u - nX); }
                if (lane < nC) { my_slot = Q.qC[(hC + lane) & (WALK_QCAP - 1)]; kind = 1; own = true; }
                else if (lane < nC + nX) { my_slot = Q.qX[(hX + lane - nC) & (WALK_QCAP - 1)]; kind = 2; own = true; }
                hC += nC; hX += nX;
            }
            __syncwarp();
            if (lane == 0) { Q.headC = hC; Q.tailC = tC; Q.headX = hX; Q.tailX = tX; }
        }
        if (EXCH) lock_release(Q, lane);
        if (EXCH && second_batch) {  // warp 0, once: its first batch waits in the queues, the second goes to the extra slots
            second_batch = false;
            have = false;
            my_slot = BLOCK + (int)lane;
            own = true;
            continue;
        }
        const unsigned mK = __ballot_sync(FULL, kind != 0);
#ifdef MCB_WALK_SYNC
        // history-per-lane form with the warps of a block kept in step: they then run the same stretch of the loop body
        // at the same time and share its instruction-cache lines (the per-SM instruction cache serves one miss for all)
        if (!EXCH) {
            if (!__syncthreads_or(mK != 0u || !exhausted)) {
                // the whole block is idle and the bank is dry
                if (!(SHARED && R.donq)) break;
                // work sharing: other blocks may still hand secondaries over; leave when no unit is alive anywhere
                long long lv = 0;
                if (threadIdx.x == 0) lv = (long long)__ldcg((const unsigned long long*)&C->live);
                if (threadIdx.x == 0 && lv > 0 && ++idle_spins > WAIT_LIMIT) { C->hang = 3; lv = 0; }
                if (!__syncthreads_or(lv > 0)) break;
                __nanosleep(4000);
                continue;
            }
        } else
#endif
        if (mK == 0u) {
            if (!exhausted) continue;
            if (!(SHARED && R.donq)) break;  // nothing held, nothing queued, nothing left to draw
            // work sharing: other lanes may still hand secondaries over; leave when no unit is alive anywhere
            long long lv = 0;
            if (lane == 0) lv = (long long)__ldcg((const unsigned long long*)&C->live);
            lv = __shfl_sync(FULL, lv, 0);
            if (lv <= 0) break;
            if (++idle_spins > WAIT_LIMIT) { C->hang = 3; break; }
            __nanosleep(4000);
            continue;
        }
        have = kind != 0;
        // ---- the event itself, on a batch of one kind (mixed only when the queues run low)
        CollideCtx c = {-1, -1, 0, 0, 0, 0.0};
        unsigned n_copy = 0;
        bool alive = false, in_material = false, unit_ended = false;
        const SlotDetail D = {st + SP_DET * WALK_SLOTS + my_slot, R.det_nn};
        if (EXCH && have) {
            const double2* s = st + my_slot;
            double2 v;
            v = ld_pair(s + SP_XY * WALK_SLOTS); p.x = v.x; p.y = v.y;
            v = ld_pair(s + SP_ZU * WALK_SLOTS); p.z = v.x; p.u = v.y;
            v = ld_pair(s + SP_VW * WALK_SLOTS); p.v = v.x; p.w = v.y;
            v = ld_pair(s + SP_ES * WALK_SLOTS); p.E = v.x; p.speed = v.y;
            v = ld_pair(s + SP_WT * WALK_SLOTS); p.wgt = v.x; p.t = v.y;
            v = ld_pair(s + SP_RNG * WALK_SLOTS); p.rng = (uint64_t)__double_as_longlong(v.x); p.cell = unpack_lo(v.y); p.hist = unpack_hi(v.y);
            v = ld_pair(s + SP_K * WALK_SLOTS); L.kC = v.x; L.kTL = v.y;
            v = ld_pair(s + SP_IDS * WALK_SLOTS); L.nsite = unpack_lo(v.x); sp = unpack_hi(v.x) & 0xffff; nch = unpack_hi(v.x) >> 16; S = unpack_lo(v.y); uidx = unpack_hi(v.y);
            v = ld_pair(s + SP_XF * WALK_SLOTS); X.f = v.x; p.Eold = v.y;
            if (TALLY) { v = ld_pair(s + SP_FIXED * WALK_SLOTS); p.told = v.x; p.n_touched = unpack_lo(v.y); p.drow = unpack_hi(v.y); }
            else { p.told = p.t; p.n_touched = 0; p.drow = -1; }
            p.row = ctx_base + my_slot;
            if (kind == 1) {
                v = ld_pair(s + SP_XT * WALK_SLOTS); X.t = v.x; X.nf = v.y;
                v = ld_pair(s + SP_XS * WALK_SLOTS); X.s = v.x; X.c = v.y;
            }
        }
        if (have) {
            if (kind == 1) {
                in_material = ev_collide_pre<TALLY>(P, p, X, uidx, D, T, C, k_eff, c);
                if (in_material) collisions++;
            } else if (kind == 2) {
                alive = ev_cross_pre<TALLY, !TALLY && !SHARED>(P, p, S, T, C, n_copy);
                crossings++;
            } else {
                alive = p.wgt > 0.0;  // census: on to the next interval, or dead after the last (no roulette, general.cpp:193)
            }
        }
        __syncwarp();
        // fission-site requests: one reservation per warp
        // Lean instances (no tallies, no secondaries) bank AFTER the scatter kinematics: the reservation is one atomic on
        // a device-wide cursor, and its round trip to L2 (ncu, round 2f: the shuffle that waits for it owned 2.2 % of all
        // stall samples, at one lane) passes under the ~1000 instructions of the scatter instead of in front of them
        constexpr bool BANK_LATE = !TALLY && !SHARED && !EXCH;
        unsigned long long site0 = 0, site_base = 0;
        unsigned site_incl = 0;
        const bool any_sites = __any_sync(FULL, c.n_sites != 0u);
        if (any_sites) {
            unsigned v = c.n_sites;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const unsigned t = __shfl_up_sync(FULL, v, d); if (lane >= (unsigned)d) v += t; }
            const unsigned total = __shfl_sync(FULL, v, 31);
            if (lane == 31) site_base = atomicAdd(&C->site_cursor, (unsigned long long)total);
            site_incl = v;
            if (!BANK_LATE) site0 = __shfl_sync(FULL, site_base, 31) + (v - c.n_sites);
        }
        const double E_coll = p.E;        // energy and stream state at the collision (what the site requests record)
        const uint64_t rng_coll = p.rng;
        if (!SHARED) { c.n_second = 0; c.n_forced = 0; n_copy = 0; }  // k-eigenvalue without splitting: nothing is ever born in flight
        unsigned short* const tab = SHARED ? R.chunk_tab + (size_t)(ctx_base + my_slot) * STACK_MAXCH : nullptr;
        if (SHARED) {  // room for the particles about to be born: borrow chunks (lane by lane; rare)
            const int need = (sp + (int)(c.n_second + n_copy) + STACK_CHUNK - 1) / STACK_CHUNK;
            unsigned m = __ballot_sync(FULL, have && need > nch);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                if ((int)lane == src) nch += chunks_take(Q, tab, nch, need - nch);
                __syncwarp();
            }
        }
        StackSink sink = {SHARED ? R.stack + (size_t)blockIdx.x * (WALK_SLOTS + WALK_EXTRA) * STACK_CHUNK : nullptr, tab, my_slot, sp, nch, C};
        if (!BANK_LATE && (c.n_sites | c.n_second)) ev_collide_bank(P, p, c, H, C, reqs, site_cap, site0, sink, &L);
        __syncwarp();  // the banking lanes rejoin the warp before the scatter kinematics
        if (have && kind == 2 && (TALLY || SHARED || alive)) alive = ev_cross_post(P, p, alive, n_copy, sink);  // (lean instances: see ev_cross_pre)
        if (in_material) alive = ev_collide_scatter<TALLY>(P, p, X, uidx, D, c, H, &L);
        __syncwarp();
        if (BANK_LATE && any_sites) {
            site0 = __shfl_sync(FULL, site_base, 31) + (site_incl - c.n_sites);
            if (c.n_sites) ev_collide_bank(P, p, c, H, C, reqs, site_cap, site0, sink, &L, E_coll, rng_coll);
            __syncwarp();
        }
        if (have && !alive) {
            if (SHARED && sp > 0) {
                if (P.comb_teeth && sp >= P.comb_bank_max) {  // handler.cpp:27-28
                    const int need = (sp + P.comb_teeth + STACK_CHUNK - 1) / STACK_CHUNK;
                    if (need > nch) { nch += chunks_take(Q, tab, nch, need - nch); sink.nch = nch; }
                    if (need <= nch) comb_stack(sink, P.comb_teeth, p.rng, &sp);
                    else C->overflow_stack = 1;
                }
                sink.pop(p);  // the history goes on with its most recent secondary (handler.cpp:22)
            } else {
                // end of the history: EstimatorK::end_history inputs (Estimator.cpp:514-525), Estimator::end_history
                if (P.ksearch) { H.kC[p.hist] = L.kC; H.kTL[p.hist] = L.kTL; H.nsite[p.hist] = L.nsite; }
                have = false;
                unit_ended = true;
            }
        }
        if (TALLY) {  // Estimator::end_history of the histories (or shared units) that ended, one after the other, by the whole warp
            unsigned m = __ballot_sync(FULL, unit_ended && (p.n_touched > 0 || p.drow >= 0));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int row = __shfl_sync(FULL, p.row, src), nt = __shfl_sync(FULL, p.n_touched, src), dr = __shfl_sync(FULL, p.drow, src);
                if (dr >= 0) merge_shared_tallies(P, T, row, nt, dr, s_sum, s_sq, lane);
                else flush_history_tallies(P, T, row, nt, s_sum, s_sq, lane);
                __syncwarp();
            }
        }
        if (SHARED && R.donq) {
            const unsigned n_end = __popc(__ballot_sync(FULL, unit_ended));
            if (lane == 0 && n_end) atomicAdd((unsigned long long*)&C->live, (unsigned long long)(-(long long)n_end));
            // Work sharing.  A fission chain followed by one lane can be arbitrarily long (HEU_sphere_leakage: 170 tracks
            // per history on average, a heavy tail far beyond): once the source bank is dry a history keeps one waiting
            // secondary and hands the others to idle lanes; before that only a stack that has grown deep is relieved.
            // Hand-over is driven by demand: while the source bank still has particles every lane is busy and only a
            // stack that has grown deep is relieved; once the bank is dry a history keeps one waiting secondary and gives
            // the others away, as long as the ring is not already stocked for the idle lanes.
            bool dry = exhausted;  // this warp's own view; a warp whose lanes are all busy never asks the bank, so look
            bool stocked = false;
            if (__any_sync(FULL, have && sp > 1)) {
                unsigned long long wh = 0;
                int av = 0;
                if (lane == 0) { wh = __ldcg(&C->walk_head); av = __ldcg(&R.donq->avail); }
                dry = dry || __shfl_sync(FULL, wh, 0) >= end - begin;
                stocked = __shfl_sync(FULL, av, 0) >= 4096;
            }
            const int thr = dry ? (stocked ? 24 : 1) : 24;
            if (have && sp > thr) {
                if (TALLY && p.drow < 0) {  // the history becomes shared: its units meet in a dense tally row
                    int r = -1;
                    if (__ldcg(T.dense_cursor) < T.dense_rows) {
                        r = atomicAdd(T.dense_cursor, 1);
                        if (r >= T.dense_rows) r = -1;
                    }
                    if (r < 0) {  // all rows handed out once: take one that its history has given back (pending = -1)
                        unsigned probe = (unsigned)p.rng;
                        for (int k = 0; k < 8 && r < 0; k++) {
                            probe = probe * 1664525u + 1013904223u;
                            const int cand = (int)(probe % (unsigned)T.dense_rows);
                            if (__ldcg(T.dense_pending + cand) == -1 && atomicCAS(T.dense_pending + cand, -1, 1) == -1) r = cand;
                        }
                    } else T.dense_pending[r] = 1;
                    if (r >= 0) { __threadfence(); p.drow = r; atomicAdd(&C->n_shared_hist, 1ull); }
                    else atomicAdd(&C->n_donate_refused, 1ull);
                }
                if (!TALLY || p.drow >= 0) {
                    while (sp > thr) {
                        Particle q = p;
                        sink.pop(q);
                        if (TALLY) atomicAdd(T.dense_pending + p.drow, 1);
                        atomicAdd((unsigned long long*)&C->live, 1ull);
                        atomicAdd(&C->n_donated, 1ull);
                        if (!donation_push(R.donq, q, C)) {  // ring full: keep it
                            atomicAdd(&C->n_donate_refused, 1ull);
                            atomicAdd((unsigned long long*)&C->live, (unsigned long long)(-1ll));
                            if (TALLY) atomicSub(T.dense_pending + p.drow, 1);
                            sink.push(q);
                            break;
                        }
                    }
                }
            }
        }
        if (SHARED) {  // histories that ended holding borrowed chunks return them
            unsigned m = __ballot_sync(FULL, !have && nch > 1);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                if ((int)lane == src) { chunks_give(Q, tab, nch); nch = 1; }
                __syncwarp();
            }
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
    if (s_sum) {  // the last warp out adds the block's private bins to the cycle sums
        unsigned done = 0;
        __threadfence_block();
        __syncwarp();
        if (lane == 0) done = atomicAdd(&Q.warps_done, 1u);
        done = __shfl_sync(FULL, done, 0);
        if (done == (unsigned)WARPS - 1u) {
            __threadfence_block();
            for (int i = (int)lane; i < R.priv_tallies; i += 32) {
                const double a = s_sum[i], b = s_sq[i];
                if (a != 0.0) { atomicAdd(T.sum + i, a); atomicAdd(T.squared + i, b); }
            }
        }
    }
}

template <bool TALLY, bool SHARED, bool EXCH>
cudaError_t plan_instance(int det_nn, int priv, int n_sm, int& blocks, size_t& smem, int& n_pairs)
{
    n_pairs = (!EXCH ? 0 : (TALLY ? SP_FIXED + 1 : SP_FIXED)) + det_nn + (det_nn + 1) / 2;
#ifdef MCB_WALK_GLOBAL_STATE
    smem = sizeof(WalkQ) + (TALLY ? (size_t)priv * 2 * sizeof(double) : 0);
#else
    smem = (size_t)n_pairs * WALK_SLOTS * sizeof(double2) + sizeof(WalkQ) + (TALLY ? (size_t)priv * 2 * sizeof(double) : 0);
#endif
    // opt in to the device's full shared memory once (the attribute is per function, not per context)
    int dev = 0, optin = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (e != cudaSuccess) return e;
    if (smem > (size_t)optin) return cudaErrorInvalidConfiguration;
    e = cudaFuncSetAttribute(k_walk<TALLY, SHARED, EXCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin);
    if (e != cudaSuccess) return e;
#ifndef MCB_WALK_GLOBAL_STATE
    if (EXCH) {
        e = cudaFuncSetAttribute(k_walk<TALLY, SHARED, EXCH>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
    }
#endif
    blocks = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks, k_walk<TALLY, SHARED, EXCH>, BLOCK, smem);
    (void)n_sm;
    return e;
}

}  // namespace

namespace mcbk {

extern thread_local uint64_t g_launches;

#define MCB_CAT2(a, b) a##b
#define MCB_CAT(a, b) MCB_CAT2(a, b)
#define MCB_FLAVOURED(name) MCB_CAT(name, MCB_WALK_FLAVOUR)
constexpr bool FT = MCB_WALK_FLAVOUR != 0;  // this flavour's kernels score tallies

// this flavour's share of the plan: fields [FT] of the per-instance arrays, and its sizes folded into the common ones
int MCB_FLAVOURED(walk_plan_part)(WalkPlan& W)
{
    int blocks = 0, n_pairs = 0;
    size_t smem = 0;
    cudaError_t e;
#define MCB_PLAN(S_, E_) plan_instance<FT, S_, E_>(W.det_nn, W.priv_tallies, W.n_sm, blocks, smem, n_pairs)
    if (W.exchange) e = W.shared ? MCB_PLAN(true, true) : MCB_PLAN(false, true);
    else e = W.shared ? MCB_PLAN(true, false) : MCB_PLAN(false, false);
#undef MCB_PLAN
    if (e != cudaSuccess) return (int)e;
    if (blocks < 1) return (int)cudaErrorInvalidConfiguration;
    W.blocks_per_sm[FT] = std::min(blocks, MCB_WALK_MINB);
    W.smem_bytes[FT] = smem;
    W.n_pairs[FT] = n_pairs;
    W.block[FT] = BLOCK;
    const int grid = W.n_sm * W.blocks_per_sm[FT];
    W.max_grid = std::max(W.max_grid, grid);
    W.n_contexts = std::max<int64_t>(W.n_contexts, (int64_t)grid * WALK_SLOTS);
    if (W.shared) {
        W.stack_records = std::max(W.stack_records, (size_t)grid * (WALK_SLOTS + WALK_EXTRA) * STACK_CHUNK);
        W.chunk_tab_entries = std::max(W.chunk_tab_entries, (size_t)grid * WALK_SLOTS * STACK_MAXCH);
    }
    W.stack_max = STACK_MAXCH * STACK_CHUNK;
#ifdef MCB_WALK_GLOBAL_STATE
    W.gstate_pairs = std::max(W.gstate_pairs, (size_t)grid * n_pairs * WALK_SLOTS);
#endif
    return 0;
}

int MCB_FLAVOURED(walk_launch_info_part)(const WalkPlan& W, int out[4])
{
    cudaFuncAttributes a;
    cudaError_t e;
#define MCB_ATTR(S_, E_) cudaFuncGetAttributes(&a, k_walk<FT, S_, E_>)
    if (W.exchange) e = W.shared ? MCB_ATTR(true, true) : MCB_ATTR(false, true);
    else e = W.shared ? MCB_ATTR(true, false) : MCB_ATTR(false, false);
#undef MCB_ATTR
    if (e != cudaSuccess) return (int)e;
    out[0] = a.numRegs; out[1] = W.n_sm * W.blocks_per_sm[FT]; out[2] = BLOCK; out[3] = (int)W.smem_bytes[FT];
    return 0;
}

void MCB_FLAVOURED(walk_part)(cudaStream_t st, const DevProblem& P, const Bank& B, uint64_t begin, uint64_t end, Counters* C, const HistoryAcc& H,
                              const TallyAcc& T, SiteReq* reqs, uint64_t site_cap, double k_eff, const WalkPlan& W, StackRec* stack,
                              unsigned short* chunk_tab, DonationQueue* donq, double2* gstate, const WalkSource& src)
{
    // persistent: every resident warp draws chunks of bank positions until the generation runs dry
    const uint64_t n = end - begin;
    const unsigned resident = (unsigned)(W.n_sm * W.blocks_per_sm[FT]);
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + BLOCK - 1) / BLOCK, resident));
    const uint64_t warps = (uint64_t)grid * WARPS;
    const uint32_t chunk = (uint32_t)std::max<uint64_t>(32, std::min<uint64_t>(128, n / (warps * 8)));
    WalkRes R;
    R.stack = stack; R.chunk_tab = chunk_tab; R.donq = donq; R.gstate = gstate; R.sm_limit = W.reserve_sms > 0 ? std::max(1, W.n_sm - W.reserve_sms) : 0; R.det_nn = W.det_nn; R.n_pairs = W.n_pairs[FT]; R.priv_tallies = FT ? W.priv_tallies : 0;
    const size_t smem = W.smem_bytes[FT];
    // per flavour four instances: problems where nothing is born in flight (k-eigenvalue without splitting) carry no
    // secondary stack; the event-sorted form is a run-time option
#define MCB_WALK(S_, E_) k_walk<FT, S_, E_><<<grid, BLOCK, smem, st>>>(P, B, (unsigned long long)begin, (unsigned long long)end, chunk, C, H, T, reqs, site_cap, k_eff, R, src)
    if (W.exchange) { if (W.shared) MCB_WALK(true, true); else MCB_WALK(false, true); }
    else { if (W.shared) MCB_WALK(true, false); else MCB_WALK(false, false); }
#undef MCB_WALK
}

#if MCB_WALK_FLAVOUR == 0
int walk_plan_part1(WalkPlan& W);
int walk_launch_info_part1(const WalkPlan& W, int out[4]);
void walk_part1(cudaStream_t st, const DevProblem& P, const Bank& B, uint64_t begin, uint64_t end, Counters* C, const HistoryAcc& H,
                const TallyAcc& T, SiteReq* reqs, uint64_t site_cap, double k_eff, const WalkPlan& W, StackRec* stack,
                unsigned short* chunk_tab, DonationQueue* donq, double2* gstate, const WalkSource& src);

int walk_plan(bool shared, bool exchange, int det_nn, int64_t n_tallies, int n_sm, WalkPlan* out)
{
    WalkPlan& W = *out;
    W.det_nn = std::max(det_nn, 1);
    W.priv_tallies = (n_tallies > 0 && n_tallies <= 256) ? (int)n_tallies : 0;
    W.n_sm = n_sm;
    W.reserve_sms = 0;
    W.exchange = exchange;
    W.shared = shared;
    W.max_grid = 0; W.n_contexts = 0; W.stack_records = 0; W.chunk_tab_entries = 0; W.gstate_pairs = 0;
    int rc = walk_plan_part0(W);
    if (rc == 0) rc = walk_plan_part1(W);
    return rc;
}

int walk_launch_info(const WalkPlan& W, bool tally, int out[4])
{
    return tally ? walk_launch_info_part1(W, out) : walk_launch_info_part0(W, out);
}

void walk(cudaStream_t st, const DevProblem& P, const Bank& B, uint64_t begin, uint64_t end, Counters* C, const HistoryAcc& H,
          const TallyAcc& T, SiteReq* reqs, uint64_t site_cap, double k_eff, const WalkPlan& W, StackRec* stack, unsigned short* chunk_tab, DonationQueue* donq, double2* gstate,
          const WalkSource& src)
{
    if (end <= begin) return;
    if (T.on) walk_part1(st, P, B, begin, end, C, H, T, reqs, site_cap, k_eff, W, stack, chunk_tab, donq, gstate, src);
    else walk_part0(st, P, B, begin, end, C, H, T, reqs, site_cap, k_eff, W, stack, chunk_tab, donq, gstate, src);
    g_launches += 1;
}
#endif

}  // namespace mcbk
