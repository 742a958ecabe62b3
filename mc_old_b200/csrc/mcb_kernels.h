// mcb_kernels.h — host-callable launchers of the kernels in mcb_kernels.cu (all asynchronous on `st`).
#ifndef MCB_KERNELS_H
#define MCB_KERNELS_H

#include "mcb_device.cuh"

namespace mcbk {

void set_device_sms(int n);  // grid sizing of the persistent kernels
uint64_t launch_count();  // kernels launched by this thread through the launchers below

// stages of one generation.  `cur` = iteration % 3 selects the queue-length counter the iteration reads
// (Counters::n_active[cur]); it fills [(cur+1)%3] and clears [(cur+2)%3].  n_hint (an upper bound of the queue
// length known to the host) only sizes the grid.
// scratch of the sorted sourcing path (bank spread over several GPUs): draws, their sorted order, stream states
struct SortScratch {
    unsigned long long *key_in, *key_out;
    uint32_t *val_in, *val_out;
    uint64_t* rng_after;
    void* temp;
    size_t temp_bytes;
    unsigned long long rot;  // global index at which this rank's sweep over the bank starts
    int key32;               // keys are 32-bit (bank below 2^32 sites): half the sort
};
size_t sort_temp_bytes(uint32_t n);
// the sorted path in pieces (bank streamed in from the host): draws + sort; positions in the sorted draws at which
// the site index reaches lo[c]; k_source for positions [q0, q0 + count) of the sweep
void pick_sort(cudaStream_t st, const DevProblem& P, int32_t first_hist, uint32_t count, uint64_t nps0, uint64_t n_bank,
               const SortScratch* sort);
// the sorted sweep for positions [q0, q0 + count) with several sites in flight per thread (walk mode: no queue, no energy_old arrays)
void source_sweep(cudaStream_t st, const Bank& B, int32_t first_hist, uint32_t q0, uint32_t count, const SourceBankView& V, const SortScratch* sort);
void preload_side_kernels(cudaStream_t st, unsigned long long* scratch);
void publish(cudaStream_t st, unsigned long long* dst, unsigned long long value);  // *dst = value in stream order
void chunk_bounds(cudaStream_t st, const SortScratch* sort, uint32_t n, const unsigned long long* lo, int n_lo, uint32_t* pos);
void source_sorted_range(cudaStream_t st, const DevProblem& P, const Bank& B, uint32_t* active, int32_t first_hist, uint32_t q0,
                         uint32_t count, uint64_t nps0, const SourceBankView& V, Counters* C, const SortScratch* sort);
// sort != nullptr: the draws are sorted by site index first and the bank is read in ascending order
void source(cudaStream_t st, const DevProblem& P, const Bank& B, uint32_t* active, int32_t first_hist, uint32_t count,
            uint64_t nps0, const SourceBankView& V, Counters* C, const SortScratch* sort);
void xs_stage(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* active, int cur, uint64_t n_hint, Counters* C);
void flight(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* active, int cur, uint64_t n_hint,
            uint32_t* evq, Counters* C, const HistoryAcc& H, const TallyAcc& T);
void collide(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* evq, int cur, uint64_t n_hint, Counters* C,
             uint32_t* next, const HistoryAcc& H, const TallyAcc& T, SiteReq* reqs,
             uint64_t site_cap, uint32_t n_slots, double k_eff);
void cross(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* evq, int cur, uint64_t n_hint, Counters* C,
           uint32_t* next, const TallyAcc& T, uint32_t n_slots);
// the walk kernel (mcb_walk.cu): every source particle in bank positions [begin, end) and all secondaries of its history
// followed to the end; particles sorted by next event in shared-memory queues, 160 history contexts per thread block.
// walk_plan() sizes the launch for this device (dynamic shared memory, blocks per SM of the scoring and non-scoring
// instance) and tells how many history contexts exist at most: the caller owns the per-context secondary stacks
// (WalkPlan::stack_records StackRec, chunk table) and tally tables (TallyAcc::tab_*).
struct WalkRes {        // kernel argument
    StackRec* stack;
    unsigned short* chunk_tab;  // n_contexts x 32: borrowed stack chunks of every history
    DonationQueue* donq;        // secondaries handed over between lanes (nullptr: every history stays with one lane)
    double2* gstate;    // slot state in global memory (build option MCB_WALK_GLOBAL_STATE), else nullptr
    int32_t det_nn, n_pairs, priv_tallies;
    int32_t sm_limit;   // blocks that land on an SM with id >= sm_limit leave at once: those SMs stay free for a kernel running beside (0: none)
};
// where the lanes of the walk kernel get their source particles from: the particle bank k_source filled (first cycle,
// fixed-source decks: the deck's <sources> are sampled there), or — fused — straight from the fission bank of the last
// generation: SourceBank::get_source (Source.cpp:42-46) happens in the lane that is about to follow the history, local
// HBM or a peer's over NVLink, and the latency of those reads hides under the transport of the other warps.
struct WalkSource {
    int32_t fused, first_hist;
    uint64_t nps0, seed0;
    const void* sorted_key;      // sorted draws (bank spread over several GPUs / streamed from the host), else nullptr
    const uint32_t* sorted_val;
    const uint64_t* rng_after;
    unsigned long long rot;
    int32_t key32, pad;          // the sorted keys are 32-bit
    const unsigned long long* ready;  // bank positions filled so far by a k_source running beside the walk (nullptr: all)
    SourceBankView V;
};
struct WalkPlan {
    int n_sm, det_nn, priv_tallies, max_grid;
    int reserve_sms;     // SMs the next launch leaves to a kernel running beside it (set by the caller per launch, default 0)
    bool shared;        // secondaries can be born in flight (fixed source / splitting)
    bool exchange;      // event-sorted form (particles change lanes through shared-memory queues) or history per lane
    int blocks_per_sm[2], n_pairs[2];   // [0] cycles that score nothing, [1] scoring cycles
    int block[2];                       // threads per block
    size_t smem_bytes[2];
    int64_t n_contexts;
    size_t stack_records, chunk_tab_entries;  // sizes of the secondary-stack arrays the caller allocates (0: no secondaries)
    int stack_max;       // particles one history can have waiting
    size_t gstate_pairs; // double2 elements of the global slot-state array the caller allocates (0: state in shared memory)
};
int walk_plan(bool shared, bool exchange, int det_nn, int64_t n_tallies, int n_sm, WalkPlan* out);  // 0 or a cudaError_t
int walk_launch_info(const WalkPlan& W, bool tally, int out[4]);  // registers per thread, full grid, block size, dynamic smem
void walk(cudaStream_t st, const DevProblem& P, const Bank& B, uint64_t begin, uint64_t end, Counters* C, const HistoryAcc& H,
          const TallyAcc& T, SiteReq* reqs, uint64_t site_cap, double k_eff, const WalkPlan& W, StackRec* stack, unsigned short* chunk_tab, DonationQueue* donq, double2* gstate,
          const WalkSource& src);
void finish(cudaStream_t st, const DevProblem& P, const Bank& B, const uint32_t* active, int cur, uint64_t n_hint, Counters* C,
            uint32_t* next, const HistoryAcc& H, const TallyAcc& T, SiteReq* reqs, uint64_t site_cap,
            uint32_t n_slots, double k_eff);

// generation close-out
void bank_sample_order(cudaStream_t st, const DevProblem& P, const SiteReq* reqs, uint64_t n, const uint32_t* offset, Site* out);
size_t scan_temp_bytes(uint32_t n);
void scan_sites(cudaStream_t st, void* temp, size_t temp_bytes, const int32_t* nsite, uint32_t* offset, uint32_t n);
void reduce_k(cudaStream_t st, const double* kC, const double* kTL, uint32_t n, Counters* C);
void entropy_history(cudaStream_t st, const DevProblem& P, const Site* bank, const uint32_t* offset,
                     const int32_t* nsite, uint32_t n_hist, Counters* C);
void entropy_histogram(cudaStream_t st, const DevProblem& P, const Site* bank, uint64_t n, unsigned long long* bins);
int tally_chunks(uint32_t n_hist);
void tally_reduce(cudaStream_t st, double* acc, int64_t stride, uint32_t n_hist, int64_t n_tallies, double* partial,
                  double* sum, double* squared);
void gather_sites(cudaStream_t st, const SourceBankView& V, uint64_t q0, uint64_t n, Site* out);  // out[q] = site q0 + q of the view
void iota(cudaStream_t st, uint32_t* a, uint32_t n);
// host-facing bank layout (n x 8 doubles + n cells) <-> Site records; dir_x = n x 3 explicit directions of a bank that
// came from the host (nullptr for banks the device made: their directions are rebuilt from the stored draws)
void pack_sites(cudaStream_t st, const double* s8, const int32_t* cells, uint64_t n, Site* out, double* dir_x);
void unpack_sites(cudaStream_t st, const Site* in, const double* dir_x, uint64_t n, double* s8, int32_t* cells);

// parity / bench kernels on plain device arrays
void xs_lookup(cudaStream_t st, const DevProblem& P, int material, const double* E, int64_t n, double* out5);
void select_channel(cudaStream_t st, const DevProblem& P, int material, int kind, const double* E, const double* xi,
                    int64_t n, int32_t* out);
void beta(cudaStream_t st, const DevProblem& P, int material, int local_n, const double* E, int64_t n, double* out);
void rng(cudaStream_t st, uint64_t seed0, const uint64_t* nps, int64_t n, int ndraw, uint64_t* out);
void geometry(cudaStream_t st, const DevProblem& P, const int32_t* cell, const double* pos, const double* dir,
              int64_t n, double* out3);
void search_cell(cudaStream_t st, const DevProblem& P, const double* pos, int64_t n, int32_t* out);
void scatter(cudaStream_t st, const DevProblem& P, int nuclide, const uint64_t* nps, int64_t n, double* io5);
void division(cudaStream_t st, const double* a, const double* b, int64_t n, double* out_shared, double* out_plain);
void watt(cudaStream_t st, const DevProblem& P, int nuclide, const uint64_t* nps, const double* E, int64_t n, double* out);

}  // namespace mcbk
#endif
