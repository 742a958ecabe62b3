// mcb_api.cu — the C-ABI of include/mcb200.h: device context, generation driver, parity entry points.
//
// mcb_run_cycle() is the body of the reference's cycle loop (Simulator::start(), handler.cpp:14-44) for the
// histories this rank owns: source resampling -> event loop on the GPU -> fission bank in canonical order ->
// per-history close-outs -> (multi-GPU) NCCL all-gather of the ranks' close-out sums; the bank stays in place and is
// read by the peers over NVLink -> k update.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "mcb_kernels.h"
#include "mcb_tables.h"

#include <chrono>
namespace {

thread_local std::string g_create_error;
// MCB_TRACE_HOST=1: wall-clock marks of the host-bank cycle on stderr
inline void trace_mark(const char* what)
{
    static const bool on = getenv("MCB_TRACE_HOST") != nullptr;
    if (!on) return;
    static auto t0 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    fprintf(stderr, "[mcb trace] %10.3f ms  %s\n", ms, what);
}

// ---- NCCL is bound at run time so that a process that already holds a libnccl (e.g. torch's) shares it ----
struct Nccl {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err)
    {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD); if (lib) break; }
        if (!lib && getenv("MCB_NCCL_LIB")) lib = dlopen(getenv("MCB_NCCL_LIB"), RTLD_NOW | RTLD_GLOBAL);
        for (const char* n : names) { if (lib) break; lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); }
        if (!lib) { err = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define MCB_SYM(field, name) field = (decltype(field))dlsym(lib, name); if (!field) { err = std::string("libnccl lacks ") + name; return false; }
        MCB_SYM(GetUniqueId, "ncclGetUniqueId") MCB_SYM(CommInitRank, "ncclCommInitRank") MCB_SYM(CommDestroy, "ncclCommDestroy")
        MCB_SYM(AllReduce, "ncclAllReduce") MCB_SYM(AllGather, "ncclAllGather") MCB_SYM(Broadcast, "ncclBroadcast")
        MCB_SYM(GroupStart, "ncclGroupStart") MCB_SYM(GroupEnd, "ncclGroupEnd") MCB_SYM(GetErrorString, "ncclGetErrorString")
#undef MCB_SYM
        return true;
    }
};
Nccl g_nccl;

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count)
    {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    cudaError_t upload(const T* src, size_t count)
    {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpy(p, src, count * sizeof(T), cudaMemcpyHostToDevice);
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

struct StageTimer {  // CUDA-event pairs per kernel class; resolved at the end of a generation
    bool on = false;
    struct Rec { int stage; cudaEvent_t a, b; };
    std::vector<Rec> recs;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get()
    {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void begin(cudaStream_t st, int stage)
    {
        if (!on) return;
        Rec r{stage, get(), get()};
        cudaEventRecord(r.a, st);
        recs.push_back(r);
    }
    void end(cudaStream_t st)
    {
        if (!on) return;
        cudaEventRecord(recs.back().b, st);
    }
    void resolve(double* ms, uint64_t* count)
    {
        for (Rec& r : recs) {
            float t = 0;
            if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.stage] += t; count[r.stage]++; }
            pool.push_back(r.a);
            pool.push_back(r.b);
        }
        recs.clear();
    }
    ~StageTimer() { for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); } for (auto e : pool) cudaEventDestroy(e); }
};
enum { ST_SOURCE = 0, ST_LOOKUP, ST_FLIGHT, ST_CROSS, ST_COLLIDE, ST_CLOSEOUT, ST_BANK, ST_FINISH, ST_STEP, ST_N };
constexpr int MCB_RING = 8;  // iterations the host may run ahead of the queue lengths it has seen

}  // namespace

struct mcb_ctx {
    std::string error;
    int device = 0, rank = 0, world = 1;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    // problem (host copies of what the close-out needs)
    int ksearch = 0, entropy_on = 0;
    uint64_t n_sample = 0, n_cycle = 0, n_passive = 0, seed = 1;
    int64_t n_tallies = 0;
    int n_materials = 0, n_nuclides = 0, n_cells = 0, n_surfaces = 0;
    std::vector<int> mat_n_nuc;
    // device problem
    DevProblem P{};
    DevBuf<DevMaterial> d_materials;
    DevBuf<DevNuclide> d_nuclides;
    DevBuf<double> d_xs_rows, d_union, d_mat_density, d_filter_grid, d_entropy_grid, d_delayed, d_bank_eold, d_bank_told;
    DevBuf<int32_t> d_hrec;
    DevBuf<int32_t> d_map, d_hash, d_mat_nuclide, d_cell_surface, d_cell_sense, d_cross_neighbor, d_attach_begin[3], d_attach_list[3];
    DevBuf<mcb_surface> d_surfaces;
    DevBuf<mcb_cell> d_cells;
    DevBuf<mcb_source> d_sources;
    DevBuf<mcb_estimator> d_estimators;
    DevBuf<mcb_score> d_scores;
    DevBuf<mcb_filter> d_filters;
    DevBuf<double> d_tdmc_time, d_tdmc_interval;
    // shard
    uint64_t shard_begin = 0, shard_count = 0;
    // banks
    uint32_t n_slots = 0, batch_hist = 0;
    DevBuf<double> d_bank_f64;       // 15 double arrays of n_slots
    DevBuf<uint64_t> d_bank_rng;
    DevBuf<int32_t> d_bank_i32;      // 4 int arrays of n_slots
    Bank B{};
    DevBuf<uint32_t> d_queue;        // active, next, evq
    uint32_t *q_active = nullptr, *q_next = nullptr, *q_ev = nullptr;
    DevBuf<Counters> d_counters;
    Counters* h_counters = nullptr;  // pinned
    unsigned long long* h_ring = nullptr;  // pinned: queue lengths of the last MCB_RING iterations
    cudaEvent_t ev_ring[MCB_RING] = {};
    uint64_t finish_below = 0;       // queue length under which the tail kernel takes over
    bool split_stages = false;       // event-queue mode: one kernel per event type (cross-check / profiling) instead of the walk kernel
    mcbk::WalkPlan plan{};           // launch shape of the walk kernel on this device
    int n_sm = 148;
    DevBuf<double2> d_gstate;        // slot state of the walk kernel when it is kept in global memory (build option)
    DevBuf<StackRec> d_stack;        // per-context LIFO stacks of same-history secondaries (the reference's Pbank)
    DevBuf<unsigned short> d_chunk_tab;
    DevBuf<DonationQueue> d_donq;    // work sharing between lanes (fixed-source problems): ring of handed-over secondaries
    DevBuf<unsigned long long> d_donq_seq;
    DevBuf<StackRec> d_donq_recs;
    DevBuf<double> d_dense;          // dense tally rows of histories shared between lanes
    DevBuf<int32_t> d_dense_pending; // [0 .. rows) pending units, [rows] the row cursor
    int dense_rows = 0;
    DevBuf<uint32_t> d_tab_key;      // per-context tally tables of the walk kernel
    DevBuf<double> d_tab_val;
    DevBuf<uint16_t> d_tab_list;
    uint32_t tab_size = 0;
    bool tab_direct = false;
    bool walk_mode = true;           // history walk (one launch per pass over the bank) instead of the event-queue loop
    // per-history accumulators
    DevBuf<double> d_hist_k;         // kC, kTL
    DevBuf<int32_t> d_nsite;
    DevBuf<uint32_t> d_site_offset;
    DevBuf<unsigned char> d_scan_temp;
    HistoryAcc H{};
    // fission bank
    uint64_t site_cap = 0, global_cap = 0;
    DevBuf<SiteReq> d_site_reqs;
    DevBuf<Site> d_local_bank[2];    // this rank's canonical bank; two of them with world > 1 (peers may still read the last one)
    DevBuf<Site> d_global_bank;      // the whole bank in one array: only for banks set / read through the host API, or without P2P
    DevBuf<double> d_host_dirs;      // explicit directions of a bank set through the host API (n x 3)
    int bank_w = 0;                  // d_local_bank[bank_w] is the one the running cycle writes
    DevBuf<double> d_slice_dirs;     // explicit directions of this rank's slice of a bank handed in from the host (world > 1)
    double* peer_dirs[MCB_MAX_WORLD] = {};   // the ranks' d_slice_dirs mapped here (CUDA IPC)
    Site* peer_bank[2][MCB_MAX_WORLD] = {};  // peer_bank[b][r] = rank r's d_local_bank[b] mapped here (CUDA IPC); [rank] = own
    bool p2p = false;                // the peers' banks are mapped: source sites are read in place over NVLink
    SourceBankView view{};           // the bank the next cycle samples (n = 0: the deck's sources)
    DevBuf<unsigned long long> d_sort_key;   // sorted sourcing (segmented bank): 2 x batch keys
    DevBuf<uint32_t> d_sort_val;             // 2 x batch values
    DevBuf<uint64_t> d_sort_rng;
    DevBuf<unsigned char> d_sort_temp;
    mcbk::SortScratch sort{};
    // streamed host bank (mcb_run_cycle_host): the source bank arrives from pinned host memory in chunks while the
    // histories that draw from the chunks already there are being walked
    struct Streamed { const double* sites8; const int32_t* cells; uint64_t n; } const* streamed = nullptr;
    cudaStream_t copy_stream = nullptr;
    static constexpr int N_CHUNK = 8;
    cudaEvent_t ev_chunk[N_CHUNK] = {};
    cudaStream_t side_stream = nullptr;   // the peer-bank sweep of chunk c + 1 runs here beside the walk of chunk c
    cudaEvent_t ev_side[N_CHUNK + 1] = {};
    DevBuf<unsigned long long> d_chunk_lo;
    DevBuf<uint32_t> d_chunk_pos;
    uint32_t* h_chunk_pos = nullptr;  // pinned
    DevBuf<double> d_io_sites;       // staging for host-facing bank I/O (n x 8 doubles)
    DevBuf<int32_t> d_io_cells;
    uint64_t n_local_sites = 0;      // local bank of the last cycle
    int bank_last = 0;               // d_local_bank[bank_last] holds it
    bool source_is_bank = false;
    // tallies
    DevBuf<double> d_tally_acc, d_tally_partial, d_tally_sum, d_tally_sq;
    std::vector<double> tally_mean, tally_uncer;     // Tally::mean / uncer accumulated over active cycles
    bool tallies_final = false;
    // entropy
    DevBuf<unsigned long long> d_entropy_bins;
    // EstimatorK state (include/Estimator.h:466-502)
    uint64_t icycle = 0, Navg = 0;
    double k = 1.0, mean_accumulator = 0.0, uncer_sq_accumulator = 0.0;
    // comm
    ncclComm_t comm = nullptr;
    DevBuf<unsigned long long> d_send, d_recv;       // per-rank close-out vector and its all-gather
    unsigned long long *h_send = nullptr, *h_recv = nullptr;  // pinned
    size_t vec_len = 0;
    // timing
    StageTimer timer;
    mcb_stage_times stage{};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    // scratch for the host-buffer entry points
    int fail(int code, const char* fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        error = buf;
        return code;
    }
};

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess) return ctx->fail(MCB_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));   \
    } while (0)
#define NK(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) return ctx->fail(MCB_ERR_COMM, "%s: %s", #call, g_nccl.GetErrorString(r_)); \
    } while (0)

extern "C" {

void mcb_shard_range(uint64_t n, int32_t rank, int32_t world, uint64_t* begin, uint64_t* count)
{
    // contiguous slices in rank order: the concatenation of the ranks' canonical fission banks is the canonical
    // global bank
    if (world < 1) world = 1;
    const unsigned __int128 N = n;
    const uint64_t b = (uint64_t)(N * (unsigned)rank / (unsigned)world);
    const uint64_t e = (uint64_t)(N * (unsigned)(rank + 1) / (unsigned)world);
    if (begin) *begin = b;
    if (count) *count = e - b;
}

int mcb_warm_up(int device)
{
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) { cudaGetLastError(); return MCB_ERR_CUDA; }
    return MCB_OK;
}

int mcb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

const char* mcb_last_error(const mcb_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

static int create_impl(mcb_ctx* ctx, const mcb_problem* p, const mcb_config* cfg)
{
    if (p->abi_version != MCB_ABI_VERSION) return ctx->fail(MCB_ERR_ARG, "mcb_problem.abi_version %d != %d", p->abi_version, MCB_ABI_VERSION);
    if (p->n_sample == 0) return ctx->fail(MCB_ERR_ARG, "n_sample is zero");
    if (p->n_sources <= 0) return ctx->fail(MCB_ERR_ARG, "[ERROR] Source bank is empty...");
    bool track_old = false, track_time = false;
    for (int e = 0; e < p->n_estimators; e++)
        if (p->estimators[e].n_filters > 4) return ctx->fail(MCB_ERR_ARG, "estimator %d has more than 4 filters", e);
    for (int f = 0; f < p->n_filters; f++) {
        if (p->filters[f].type == MCB_FILTER_TIME) track_time = true;
        if (p->filters[f].type == MCB_FILTER_ENERGY_OLD) track_old = true;
    }
    for (int e = 0; e < p->n_estimators; e++) {
        const int sim = p->estimators[e].simulate;
        if (sim < MCB_SIM_NONE || sim > MCB_SIM_FISSION_DELAYED + 5) return ctx->fail(MCB_ERR_ARG, "estimator %d: unknown simulate kind %d", e, sim);
        if (sim != MCB_SIM_NONE) track_old = true;
    }
    for (int k = 0; k < p->n_scores; k++)
        if (p->scores[k].score >= MCB_SCORE_SCATTER_OLD) {
            track_old = true;
            if (p->scores[k].group < 0 || p->scores[k].group > 5) return ctx->fail(MCB_ERR_ARG, "score %d: precursor group %d", k, p->scores[k].group);
        }
    trace_mark("create: begin");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return ctx->fail(MCB_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
    ctx->device = cfg ? cfg->device : 0;
    ctx->rank = cfg ? cfg->rank : 0;
    ctx->world = cfg && cfg->world > 0 ? cfg->world : 1;
    if (ctx->rank < 0 || ctx->rank >= ctx->world) return ctx->fail(MCB_ERR_ARG, "rank %d outside world %d", ctx->rank, ctx->world);
    trace_mark("create: device count known (driver initialised)");
    CK(cudaSetDevice(ctx->device));
    if (cfg && cfg->stream) ctx->stream = (cudaStream_t)cfg->stream;
    else { CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)); ctx->own_stream = true; }
    trace_mark("create: device context and stream");
    ctx->timer.on = (cfg && (cfg->reserved & 1)) || getenv("MCB_STAGE_TIMES");

    ctx->ksearch = p->ksearch; ctx->entropy_on = p->entropy_on;
    ctx->n_sample = p->n_sample; ctx->n_cycle = p->n_cycle; ctx->n_passive = p->n_passive;
    ctx->seed = p->seed ? p->seed : 1;
    ctx->n_tallies = p->n_tallies;
    ctx->n_materials = p->n_materials; ctx->n_nuclides = p->n_nuclides; ctx->n_cells = p->n_cells; ctx->n_surfaces = p->n_surfaces;

    // ---- nuclear data: rows + per-material union grid / map / hash ----
    CK(ctx->d_xs_rows.upload(p->xs_rows, (size_t)p->n_xs_rows * MCB_XS_ROW));
    CK(ctx->d_delayed.upload(p->delayed_data, (size_t)std::max<int64_t>(p->n_delayed_data, 0)));
    std::vector<DevNuclide> nuc(p->n_nuclides);
    for (int n = 0; n < p->n_nuclides; n++) {
        const mcb_nuclide& N = p->nuclides[n];
        DevNuclide& D = nuc[n];
        D.rows = ctx->d_xs_rows.p + (size_t)N.row_begin * MCB_XS_ROW;
        D.n_rows = N.n_rows; D.has_delayed = N.has_delayed; D.A = N.A;
        D.fg_beta = std::sqrt(2.0659834e-11 * N.A);  // Reaction.cpp:33; IEEE sqrt, the same number on the host and on the device
        memcpy(D.watt_a, N.watt_a, sizeof(D.watt_a)); memcpy(D.watt_b, N.watt_b, sizeof(D.watt_b)); memcpy(D.watt_g, N.watt_g, sizeof(D.watt_g));
        memcpy(D.lambda, N.lambda, sizeof(D.lambda)); memcpy(D.fraction, N.fraction, sizeof(D.fraction));
        D.chid_E = ctx->d_delayed.p ? ctx->d_delayed.p + N.chid_E_begin : nullptr;
        for (int g = 0; g < 6; g++) { D.chid_cdf[g] = ctx->d_delayed.p ? ctx->d_delayed.p + N.chid_cdf_begin[g] : nullptr; D.chid_cdf_n[g] = N.chid_cdf_n[g]; }
    }
    CK(ctx->d_nuclides.upload(nuc.data(), nuc.size()));
    trace_mark("create: nuclear data uploaded");
    std::vector<mcb::MaterialTables> tabs(p->n_materials);
    size_t nU = 0, nmap = 0, nhash = 0;
    for (int m = 0; m < p->n_materials; m++) {
        mcb::build_material_tables(p, m, getenv("MCB_HASH_BITS") ? atoi(getenv("MCB_HASH_BITS")) : MCB_HASH_BITS_DEFAULT, tabs[m]);
        nU += tabs[m].U.size(); nmap += tabs[m].map.size(); nhash += tabs[m].hash.size();
        ctx->mat_n_nuc.push_back(tabs[m].n_nuc);
    }
    trace_mark("create: union grids / hash tables built (host)");
    std::vector<double> U; std::vector<int32_t> map, hash, hrec;
    U.reserve(nU); map.reserve(nmap); hash.reserve(nhash);
    std::vector<DevMaterial> mats(p->n_materials);
    std::vector<size_t> oU(p->n_materials), omap(p->n_materials), ohash(p->n_materials), ohrec(p->n_materials);
    for (int m = 0; m < p->n_materials; m++) {
        oU[m] = U.size(); omap[m] = map.size(); ohash[m] = hash.size(); ohrec[m] = hrec.size();
        hrec.insert(hrec.end(), tabs[m].hrec.begin(), tabs[m].hrec.end());
        hrec.resize((hrec.size() + 3) & ~(size_t)3);  // every material's records start on a 16-byte boundary
        U.insert(U.end(), tabs[m].U.begin(), tabs[m].U.end());
        map.insert(map.end(), tabs[m].map.begin(), tabs[m].map.end());
        hash.insert(hash.end(), tabs[m].hash.begin(), tabs[m].hash.end());
    }
    CK(ctx->d_union.upload(U.data(), U.size()));
    CK(ctx->d_map.upload(map.data(), map.size()));
    CK(ctx->d_hash.upload(hash.data(), hash.size()));
    CK(ctx->d_hrec.upload(hrec.data(), hrec.size()));
    for (int m = 0; m < p->n_materials; m++) {
        DevMaterial& M = mats[m];
        M.nuc_begin = p->mat_begin[m]; M.n_nuc = tabs[m].n_nuc;
        M.nU = (int32_t)tabs[m].U.size(); M.n_hash = tabs[m].n_hash; M.shift = tabs[m].shift; M.hstride = tabs[m].hrec_stride;
        M.key_min = tabs[m].key_min;
        M.U = ctx->d_union.p + oU[m]; M.map = ctx->d_map.p + omap[m]; M.hash = ctx->d_hash.p + ohash[m]; M.hrec = ctx->d_hrec.p + ohrec[m];
    }
    CK(ctx->d_materials.upload(mats.data(), mats.size()));
    const int n_mat_nuc = p->mat_begin[p->n_materials];
    CK(ctx->d_mat_nuclide.upload(p->mat_nuclide, n_mat_nuc));
    CK(ctx->d_mat_density.upload(p->mat_density, n_mat_nuc));
    // ---- geometry, sources, estimators ----
    CK(ctx->d_surfaces.upload(p->surfaces, p->n_surfaces));
    CK(ctx->d_cells.upload(p->cells, p->n_cells));
    CK(ctx->d_cell_surface.upload(p->cell_surface, p->n_cell_surface));
    CK(ctx->d_cell_sense.upload(p->cell_sense, p->n_cell_surface));
    {
        // The cell behind a surface, where search_cell's answer (general.cpp:26-34: the FIRST cell in deck order that
        // contains the point) can be told beforehand: a point on side g of surface s is outside every cell that holds
        // (s, -g); if the first cell in deck order that the side does not rule out consists of (s, g) alone, it contains
        // every such point and is the answer.  The crossing code evaluates s at the nudged point (the same arithmetic as
        // test_point, so the same decision), takes the table's cell when the value is strictly on one side, and searches
        // otherwise.  Typical hit: the one-surface "outside" / graveyard cells every leaking particle enters.
        std::vector<int32_t> nb;
        mcb::build_cross_neighbors(p, nb);   // mcb_tables.cpp
        CK(ctx->d_cross_neighbor.upload(nb.data(), nb.size()));
    }
    CK(ctx->d_sources.upload(p->sources, p->n_sources));
    CK(ctx->d_estimators.upload(p->estimators, p->n_estimators));
    CK(ctx->d_scores.upload(p->scores, p->n_scores));
    {
        // filters with the same grid share one copy on the device: the walk kernel's bin cache (filter_bin) is keyed by
        // the grid, and the estimators of a TRMM tally set all carry the same energy grid
        std::vector<mcb_filter> fl(p->filters, p->filters + p->n_filters);
        for (int f = 1; f < p->n_filters; f++)
            for (int g = 0; g < f; g++)
                if (fl[g].grid_n == fl[f].grid_n &&
                    !memcmp(p->filter_grid + fl[g].grid_begin, p->filter_grid + p->filters[f].grid_begin, sizeof(double) * (size_t)fl[f].grid_n)) {
                    fl[f].grid_begin = fl[g].grid_begin;
                    break;
                }
        CK(ctx->d_filters.upload(fl.data(), p->n_filters));
    }
    CK(ctx->d_filter_grid.upload(p->filter_grid, p->n_filter_grid));
    bool splitting = false;
    {
        double imp = -1.0;
        for (int c = 0; c < p->n_cells; c++) {
            const double I = p->cells[c].importance;
            if (I == 0.0) continue;
            if (imp < 0) imp = I; else if (I != imp) splitting = true;
        }
    }
    for (int kind = 0; kind < 3; kind++) {
        const int n_ids = kind == MCB_ATTACH_SURFACE ? p->n_surfaces : p->n_cells;
        std::vector<int32_t> begin(n_ids + 1, 0), list;
        for (int id = 0; id < n_ids; id++) {
            begin[id] = (int32_t)list.size();
            for (int e = 0; e < p->n_estimators; e++) {
                const mcb_estimator& E = p->estimators[e];
                if (E.attach != kind) continue;
                // the geometry filter: the estimator's first, or its second behind a TDMC filter (setup.cpp:703-741)
                const mcb_filter& F0 = p->filters[E.filter_begin + (p->filters[E.filter_begin].type == MCB_FILTER_TDMC ? 1 : 0)];
                for (int i = 0; i < F0.grid_n; i++)
                    if ((int)p->filter_grid[F0.grid_begin + i] == id) { list.push_back(e); break; }
            }
        }
        begin[n_ids] = (int32_t)list.size();
        CK(ctx->d_attach_begin[kind].upload(begin.data(), begin.size()));
        CK(ctx->d_attach_list[kind].upload(list.data(), list.size()));
    }
    int entropy_bins = 0;
    if (p->entropy_on) {
        const int ng = p->entropy_n[0] + p->entropy_n[1] + p->entropy_n[2];
        CK(ctx->d_entropy_grid.upload(p->entropy_grid, ng));
        entropy_bins = (p->entropy_n[0] - 1) * (p->entropy_n[1] - 1) * (p->entropy_n[2] - 1);
        CK(ctx->d_entropy_bins.alloc(entropy_bins));
    }

    DevProblem& P = ctx->P;
    P.ksearch = p->ksearch; P.n_materials = p->n_materials; P.n_nuclides = p->n_nuclides; P.n_surfaces = p->n_surfaces;
    P.n_cells = p->n_cells; P.n_sources = p->n_sources; P.n_estimators = p->n_estimators; P.entropy_on = p->entropy_on;
    P.shared_histories = (!p->ksearch || splitting) ? 1 : 0;
    P.track_old = track_old ? 1 : 0;
    P.track_time = track_time ? 1 : 0; P.pad = 0;
    P.comb_teeth = p->comb_on ? p->comb_teeth : 0; P.comb_bank_max = p->comb_bank_max;
    P.n_tdmc = 0; P.tdmc_time = nullptr; P.tdmc_interval = nullptr;
    if (p->tdmc_on) {
        if (p->ksearch) return ctx->fail(MCB_ERR_ARG, "ksearch and tdmc could not coexist");
        if (p->n_tdmc < 1 || !p->tdmc_time || !p->tdmc_interval) return ctx->fail(MCB_ERR_ARG, "tdmc_on without census times");
        CK(ctx->d_tdmc_time.upload(p->tdmc_time, p->n_tdmc));
        CK(ctx->d_tdmc_interval.upload(p->tdmc_interval, p->n_tdmc));
        P.n_tdmc = p->n_tdmc; P.tdmc_time = ctx->d_tdmc_time.p; P.tdmc_interval = ctx->d_tdmc_interval.p;
    }
    if (p->comb_on && (p->comb_teeth < 1 || p->comb_teeth > 256 || p->comb_bank_max < 1)) return ctx->fail(MCB_ERR_ARG, "particle comb: bank_max >= 1 and 1 <= teeth <= 256");
    P.wr = p->wr; P.ws = p->ws; P.seed0 = ctx->seed; P.n_sample = p->n_sample;
    P.materials = ctx->d_materials.p; P.nuclides = ctx->d_nuclides.p; P.mat_nuclide = ctx->d_mat_nuclide.p; P.mat_density = ctx->d_mat_density.p;
    P.surfaces = ctx->d_surfaces.p; P.cells = ctx->d_cells.p; P.cell_surface = ctx->d_cell_surface.p; P.cell_sense = ctx->d_cell_sense.p; P.cross_neighbor = ctx->d_cross_neighbor.p;
    P.sources = ctx->d_sources.p; P.estimators = ctx->d_estimators.p; P.scores = ctx->d_scores.p; P.filters = ctx->d_filters.p;
    P.filter_grid = ctx->d_filter_grid.p;
    for (int kind = 0; kind < 3; kind++) { P.attach_begin[kind] = ctx->d_attach_begin[kind].p; P.attach_list[kind] = ctx->d_attach_list[kind].p; }
    for (int a = 0; a < 3; a++) P.entropy_n[a] = p->entropy_n[a];
    P.entropy_bins = entropy_bins; P.entropy_grid = ctx->d_entropy_grid.p;

    trace_mark("create: problem description uploaded");
    // ---- shard and banks ----
    mcb_shard_range(p->n_sample, ctx->rank, ctx->world, &ctx->shard_begin, &ctx->shard_count);
    if (ctx->shard_count >= (1ull << 31)) return ctx->fail(MCB_ERR_ARG, "more than 2^31 histories per GPU per generation");
    ctx->split_stages = (cfg && (cfg->reserved & 2)) || (getenv("MCB_MODE") && !strcmp(getenv("MCB_MODE"), "split"));
    if (ctx->split_stages && p->tdmc_on) return ctx->fail(MCB_ERR_ARG, "the time-dependent mode runs in the walk kernel only (not with the event-queue stages)");
    ctx->walk_mode = !ctx->split_stages;
    {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, ctx->device));
        ctx->n_sm = prop.multiProcessorCount;
        mcbk::set_device_sms(ctx->n_sm);
    }
    const bool split = ctx->split_stages;
    // event-queue mode: secondaries (same-history fission neutrons, split copies) need free slots behind the primaries
    // of a batch; the walk kernel keeps them on per-history stacks and needs one slot per source particle
    const uint64_t per_hist = (split && P.shared_histories) ? 6 : 1;
    uint64_t cap = cfg && cfg->bank_capacity > 0 ? (uint64_t)cfg->bank_capacity : (1ull << 26);
    cap = std::min<uint64_t>(cap, (1ull << 31) - 1);
    uint64_t batch = std::max<uint64_t>(1, std::min<uint64_t>(ctx->shard_count, cap / per_hist));
    if (split && p->n_tallies > 0) {
        // event-queue mode: dense per-history tally accumulators of a batch, kept under ~8 GiB
        const uint64_t lim = std::max<uint64_t>(1024, (8ull << 30) / (8ull * (uint64_t)p->n_tallies));
        batch = std::min(batch, lim);
    }
    ctx->batch_hist = (uint32_t)batch;
    ctx->n_slots = (uint32_t)std::min<uint64_t>(batch * per_hist + 1024, (1ull << 31) - 1);
    const size_t ns = ctx->n_slots;
    CK(ctx->d_bank_f64.alloc((split ? 15 : 10) * ns));
    CK(ctx->d_bank_rng.alloc(ns));
    CK(ctx->d_bank_i32.alloc((split ? 4 : 2) * ns));
    {
        double* f = ctx->d_bank_f64.p;
        Bank& B = ctx->B;
        B.x = f; B.y = f + ns; B.z = f + 2 * ns; B.u = f + 3 * ns; B.v = f + 4 * ns; B.w = f + 5 * ns; B.E = f + 6 * ns;
        B.speed = f + 7 * ns; B.wgt = f + 8 * ns; B.t = f + 9 * ns;
        B.St = B.Ss = B.Sc = B.Sf = B.nSf = nullptr;
        if (split) { B.St = f + 10 * ns; B.Ss = f + 11 * ns; B.Sc = f + 12 * ns; B.Sf = f + 13 * ns; B.nSf = f + 14 * ns; }
        B.rng = ctx->d_bank_rng.p;
        int32_t* i = ctx->d_bank_i32.p;
        B.cell = i; B.hist = i + ns;
        B.uidx = B.surf = nullptr;
        if (split) { B.uidx = i + 2 * ns; B.surf = i + 3 * ns; }
        B.Eold = nullptr;
        B.told = nullptr;
        if (split && track_old) { CK(ctx->d_bank_eold.alloc(ns)); B.Eold = ctx->d_bank_eold.p; }
        if (split && track_time) { CK(ctx->d_bank_told.alloc(ns)); CK(cudaMemset(ctx->d_bank_told.p, 0, ns * sizeof(double))); B.told = ctx->d_bank_told.p; }
    }
    if (split) {
        CK(ctx->d_queue.alloc(3 * ns));
        ctx->q_active = ctx->d_queue.p; ctx->q_next = ctx->d_queue.p + ns; ctx->q_ev = ctx->d_queue.p + 2 * ns;
    }
    CK(ctx->d_counters.alloc(1));
    CK(cudaMemset(ctx->d_counters.p, 0, sizeof(Counters)));
    CK(cudaMallocHost((void**)&ctx->h_counters, sizeof(Counters)));
    CK(cudaMallocHost((void**)&ctx->h_ring, MCB_RING * sizeof(unsigned long long)));
    for (int i = 0; i < MCB_RING; i++) CK(cudaEventCreateWithFlags(&ctx->ev_ring[i], cudaEventDisableTiming));
    ctx->finish_below = getenv("MCB_FINISH_BELOW") ? strtoull(getenv("MCB_FINISH_BELOW"), nullptr, 10) : (uint64_t)ctx->n_sm * 256ull;
    const size_t nh = std::max<uint64_t>(ctx->shard_count, 1);
    CK(ctx->d_hist_k.alloc(2 * nh));
    CK(ctx->d_nsite.alloc(nh));
    CK(ctx->d_site_offset.alloc(nh + 1));
    CK(ctx->d_scan_temp.alloc(mcbk::scan_temp_bytes((uint32_t)nh) + 256));
    ctx->H.kC = ctx->d_hist_k.p; ctx->H.kTL = ctx->d_hist_k.p + nh; ctx->H.nsite = ctx->d_nsite.p;
    if (p->ksearch) {
        // the first generations of a badly converged source bank up to ~3 sites per history (k_cycle of a 14 MeV
        // point source in HEU is 2.7); leave room for 4
        ctx->site_cap = cfg && cfg->site_capacity > 0 ? (uint64_t)cfg->site_capacity : 4 * ctx->shard_count + 4096;
        CK(ctx->d_site_reqs.alloc(ctx->site_cap));
        CK(ctx->d_local_bank[0].alloc(ctx->site_cap));
        ctx->global_cap = ctx->site_cap;
        if (ctx->world > 1) {
            CK(ctx->d_local_bank[1].alloc(ctx->site_cap));
            ctx->global_cap = cfg && cfg->site_capacity > 0 ? (uint64_t)cfg->site_capacity * ctx->world : 4 * p->n_sample + 4096ull * ctx->world;
        }
    }
    if (!split) {
        // walk kernel: launch shape on this device, per-context secondary stacks and tally tables
        int det_nn = 1;
        for (int m = 0; m < p->n_materials; m++) det_nn = std::max(det_nn, tabs[m].n_nuc);
        const int rc = mcbk::walk_plan(P.shared_histories != 0, getenv("MCB_WALK_EXCHANGE") && atoi(getenv("MCB_WALK_EXCHANGE")) != 0 && !p->tdmc_on, det_nn, p->n_tallies, ctx->n_sm, &ctx->plan);
        if (rc != 0) return ctx->fail(MCB_ERR_CUDA, "walk kernel does not fit this device: %s", cudaGetErrorString((cudaError_t)rc));
        const size_t n_ctx = (size_t)ctx->plan.n_contexts;
        if (ctx->plan.gstate_pairs) CK(ctx->d_gstate.alloc(ctx->plan.gstate_pairs));
        if (ctx->plan.stack_records) { CK(ctx->d_stack.alloc(ctx->plan.stack_records)); CK(ctx->d_chunk_tab.alloc(ctx->plan.chunk_tab_entries)); }
        if (P.shared_histories && !p->ksearch && !p->comb_on && !getenv("MCB_NO_SHARING")) {  // (the comb works on a history's whole bank)
            // fixed-source problems: lanes hand waiting secondaries to idle lanes (k-eigenvalue problems only split, their
            // families are small, and their fission sites keep the order of one lane)
            const uint32_t cap = 1u << 16;
            std::vector<unsigned long long> seq(cap);
            for (uint32_t i = 0; i < cap; i++) seq[i] = i;
            CK(ctx->d_donq_seq.upload(seq.data(), cap));
            CK(ctx->d_donq_recs.alloc(cap));
            DonationQueue dq;
            memset(&dq, 0, sizeof(dq));
            dq.seq = ctx->d_donq_seq.p; dq.recs = ctx->d_donq_recs.p; dq.cap_mask = cap - 1;
            CK(ctx->d_donq.upload(&dq, 1));
            if (p->n_tallies > 0) {
                ctx->dense_rows = (int)std::max<int64_t>(64, std::min<int64_t>(65536, (int64_t)(1ull << 31) / (8 * p->n_tallies)));
                CK(ctx->d_dense.alloc((size_t)ctx->dense_rows * p->n_tallies));
                CK(cudaMemset(ctx->d_dense.p, 0, (size_t)ctx->dense_rows * p->n_tallies * sizeof(double)));
                CK(ctx->d_dense_pending.alloc((size_t)ctx->dense_rows + 1));
                CK(cudaMemset(ctx->d_dense_pending.p, 0, ((size_t)ctx->dense_rows + 1) * sizeof(int32_t)));
            }
        }
        if (p->n_tallies > 0) {
            // table of a history: direct (one position per tally, values only) while that stays under ~24 GiB in all
            // and positions fit the 16-bit touched list; else open-addressed, a power of two of at most 8192 entries
            const uint64_t dsize = ((uint64_t)p->n_tallies + 15) & ~15ull;
            ctx->tab_direct = dsize <= 65536 && n_ctx * dsize * 10 <= (24ull << 30) && !getenv("MCB_HASHED_TALLIES");
            uint32_t size = (uint32_t)dsize;
            if (!ctx->tab_direct) {
                size = 2;
                while (size < (uint32_t)std::min<int64_t>(p->n_tallies, 8192)) size <<= 1;
                while (size > 256 && n_ctx * (size_t)size * 14 > (24ull << 30)) size >>= 1;
            }
            ctx->tab_size = size;
            CK(ctx->d_tab_val.alloc(n_ctx * size)); CK(ctx->d_tab_list.alloc(n_ctx * size));
            if (ctx->tab_direct) CK(cudaMemset(ctx->d_tab_val.p, 0, n_ctx * size * sizeof(double)));
            else {
                CK(ctx->d_tab_key.alloc(n_ctx * size));
                CK(cudaMemset(ctx->d_tab_key.p, 0, n_ctx * size * sizeof(uint32_t)));
            }
        }
    }
    if (p->n_tallies > 0) {
        if (split) {
            CK(ctx->d_tally_acc.alloc((size_t)p->n_tallies * batch));
            CK(cudaMemset(ctx->d_tally_acc.p, 0, (size_t)p->n_tallies * batch * sizeof(double)));
            CK(ctx->d_tally_partial.alloc((size_t)p->n_tallies * mcbk::tally_chunks((uint32_t)batch) * 2));
        }
        CK(ctx->d_tally_sum.alloc(p->n_tallies));
        CK(ctx->d_tally_sq.alloc(p->n_tallies));
        CK(cudaMemset(ctx->d_tally_sum.p, 0, p->n_tallies * sizeof(double)));
        CK(cudaMemset(ctx->d_tally_sq.p, 0, p->n_tallies * sizeof(double)));
        ctx->tally_mean.assign(p->n_tallies, 0.0);
        ctx->tally_uncer.assign(p->n_tallies, 0.0);
    }
    trace_mark("create: banks, stacks and tally tables allocated");
    if (ctx->world > MCB_MAX_WORLD) return ctx->fail(MCB_ERR_ARG, "world %d > %d", ctx->world, MCB_MAX_WORLD);
    ctx->vec_len = 16 + (size_t)entropy_bins + 2 * (size_t)std::max<int64_t>(p->n_tallies, 0);
    CK(ctx->d_send.alloc(ctx->vec_len));
    CK(ctx->d_recv.alloc(ctx->vec_len * ctx->world));
    CK(cudaMallocHost((void**)&ctx->h_send, ctx->vec_len * sizeof(unsigned long long)));
    CK(cudaMallocHost((void**)&ctx->h_recv, ctx->vec_len * ctx->world * sizeof(unsigned long long)));
    CK(cudaEventCreate(&ctx->ev0)); CK(cudaEventCreate(&ctx->ev1)); CK(cudaEventCreate(&ctx->ev2));
    // the cross-section tables are read by every lookup and fit in L2 many times over: ask for them to persist
    {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 && ctx->d_xs_rows.n) {
            const size_t bytes = std::min<size_t>(ctx->d_xs_rows.n * sizeof(double), (size_t)prop.accessPolicyMaxWindowSize);
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min<size_t>(bytes, (size_t)prop.persistingL2CacheMaxSize));
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.base_ptr = ctx->d_xs_rows.p;
            attr.accessPolicyWindow.num_bytes = bytes;
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            if (!getenv("MCB_NO_L2_PERSIST")) cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);
            cudaGetLastError();
        }
    }
    // Kernels are loaded lazily, and loading one waits for every running kernel to end: a kernel first launched BESIDE a
    // persistent kernel that waits for it would never start.  Everything the side stream launches is launched once here.
    trace_mark("create: before kernel preload");
    mcbk::preload_side_kernels(ctx->stream, &ctx->d_counters.p->src_ready);
    CK(cudaStreamSynchronize(ctx->stream));
    trace_mark("create: done");
    return MCB_OK;
}

int mcb_create(const mcb_problem* problem, const mcb_config* config, mcb_ctx** out)
{
    if (!problem || !out) { g_create_error = "mcb_create: null argument"; return MCB_ERR_ARG; }
    mcb_ctx* ctx = new mcb_ctx;
    const int rc = create_impl(ctx, problem, config);
    if (rc != MCB_OK) {
        g_create_error = ctx->error;
        mcb_destroy(ctx);
        *out = nullptr;
        return rc;
    }
    *out = ctx;
    return MCB_OK;
}

void mcb_destroy(mcb_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->h_ring) cudaFreeHost(ctx->h_ring);
    if (ctx->h_send) cudaFreeHost(ctx->h_send);
    if (ctx->h_chunk_pos) cudaFreeHost(ctx->h_chunk_pos);
    for (int i = 0; i < mcb_ctx::N_CHUNK; i++) if (ctx->ev_chunk[i]) cudaEventDestroy(ctx->ev_chunk[i]);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    for (int i = 0; i <= mcb_ctx::N_CHUNK; i++) if (ctx->ev_side[i]) cudaEventDestroy(ctx->ev_side[i]);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->h_recv) cudaFreeHost(ctx->h_recv);
    for (int b = 0; b < 2; b++)
        for (int r = 0; r < ctx->world && r < MCB_MAX_WORLD; r++)
            if (r != ctx->rank && ctx->peer_bank[b][r]) cudaIpcCloseMemHandle(ctx->peer_bank[b][r]);
    for (int r = 0; r < ctx->world && r < MCB_MAX_WORLD; r++) if (r != ctx->rank && ctx->peer_dirs[r]) cudaIpcCloseMemHandle(ctx->peer_dirs[r]);
    for (int i = 0; i < MCB_RING; i++) if (ctx->ev_ring[i]) cudaEventDestroy(ctx->ev_ring[i]);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->ev2) cudaEventDestroy(ctx->ev2);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int mcb_comm_unique_id(char id[128])
{
    std::string err;
    if (!g_nccl.load(err)) { g_create_error = err; return MCB_ERR_COMM; }
    ncclUniqueId uid;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (g_nccl.GetUniqueId(&uid) != ncclSuccess) { g_create_error = "ncclGetUniqueId failed"; return MCB_ERR_COMM; }
    memcpy(id, &uid, 128);
    return MCB_OK;
}

int mcb_comm_init(mcb_ctx* ctx, const char id[128])
{
    if (!ctx) return MCB_ERR_ARG;
    if (ctx->world == 1) return MCB_OK;
    std::string err;
    if (!g_nccl.load(err)) return ctx->fail(MCB_ERR_COMM, "%s", err.c_str());
    CK(cudaSetDevice(ctx->device));
    ncclUniqueId uid;
    memcpy(&uid, id, 128);
    NK(g_nccl.CommInitRank(&ctx->comm, ctx->world, uid, ctx->rank));
    // Map every peer's fission banks into this process (CUDA IPC over NVLink / NVSwitch): the source kernel then reads
    // the sites it draws straight from the HBM of the rank that banked them and the bank is never gathered.
    // Random peer reads thrash the address translation of the peer mappings once the banks exceed ~1 GB (134 ms per
    // generation at 8 ranks), so the source step sorts its draws by site index and sweeps the global bank in ascending
    // order, every rank starting at its own slice (k_pick / k_source): in place for any number of ranks.  MCB_NO_P2P=1
    // (or a failed mapping) selects the NCCL gather of the slices instead.
    const bool want_p2p = getenv("MCB_P2P") ? atoi(getenv("MCB_P2P")) != 0 : true;
    if (ctx->ksearch && want_p2p && !getenv("MCB_NO_P2P")) {
        const int W = ctx->world;
        // three mappings per peer: its two (alternating) fission banks and the direction array of host-provided slices
        if (cudaMalloc((void**)&ctx->d_slice_dirs.p, ctx->site_cap * 3 * sizeof(double)) == cudaSuccess) ctx->d_slice_dirs.n = ctx->site_cap * 3;
        else { ctx->d_slice_dirs.p = nullptr; cudaGetLastError(); }
        constexpr int NH = 3;
        cudaIpcMemHandle_t mine[NH];
        memset(mine, 0, sizeof(mine));
        bool ok = cudaIpcGetMemHandle(&mine[0], ctx->d_local_bank[0].p) == cudaSuccess &&
                  cudaIpcGetMemHandle(&mine[1], ctx->d_local_bank[1].p) == cudaSuccess &&
                  ctx->d_slice_dirs.p && cudaIpcGetMemHandle(&mine[2], ctx->d_slice_dirs.p) == cudaSuccess;
        cudaGetLastError();
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
        const size_t rec = NH * sizeof(cudaIpcMemHandle_t) + 8;  // the handles + "I could export" flag
        DevBuf<unsigned char> d_mine, d_all;
        std::vector<unsigned char> h_mine(rec, 0), h_all(rec * W, 0);
        memcpy(h_mine.data(), mine, NH * sizeof(cudaIpcMemHandle_t));
        h_mine[NH * sizeof(cudaIpcMemHandle_t)] = ok ? 1 : 0;
        CK(d_mine.upload(h_mine.data(), rec));
        CK(d_all.alloc(rec * W));
        NK(g_nccl.AllGather(d_mine.p, d_all.p, rec, ncclChar, ctx->comm, ctx->stream));
        CK(cudaMemcpyAsync(h_all.data(), d_all.p, rec * W, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        for (int r = 0; r < W; r++) ok = ok && h_all[r * rec + NH * sizeof(cudaIpcMemHandle_t)] == 1;
        for (int r = 0; r < W && ok; r++) {
            for (int b = 0; b < NH; b++) {
                if (r == ctx->rank) {
                    if (b < 2) ctx->peer_bank[b][r] = ctx->d_local_bank[b].p; else ctx->peer_dirs[r] = ctx->d_slice_dirs.p;
                    continue;
                }
                cudaIpcMemHandle_t h;
                memcpy(&h, h_all.data() + r * rec + b * sizeof(cudaIpcMemHandle_t), sizeof(h));
                void* ptr = nullptr;
                if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; cudaGetLastError(); break; }
                if (b < 2) ctx->peer_bank[b][r] = (Site*)ptr; else ctx->peer_dirs[r] = (double*)ptr;
            }
        }
        // every rank must take the same path: agree on the outcome
        unsigned long long flag = ok ? 1 : 0;
        DevBuf<unsigned long long> d_flag;
        CK(d_flag.upload(&flag, 1));
        NK(g_nccl.AllReduce(d_flag.p, d_flag.p, 1, ncclUint64, ncclMin, ctx->comm, ctx->stream));
        CK(cudaMemcpyAsync(&flag, d_flag.p, sizeof(flag), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->p2p = flag == 1;
    }
    return MCB_OK;
}

double mcb_get_k(const mcb_ctx* ctx) { return ctx ? ctx->k : 0.0; }
void mcb_set_k(mcb_ctx* ctx, double k) { if (ctx) ctx->k = k; }

static double fx_to_double(unsigned long long lo, unsigned long long hi)
{
    const long double v = (long double)hi * 4294967296.0L + (long double)lo;
    return (double)(v / (long double)MCB_FX_SCALE);
}

static TallyAcc tally_acc(mcb_ctx* ctx, uint32_t h0, bool tally_on)
{
    TallyAcc T;
    memset(&T, 0, sizeof(T));
    T.acc = ctx->d_tally_acc.p;  // event-queue mode only (nullptr selects the per-history tables of the walk kernel)
    T.stride = ctx->batch_hist; T.first_hist = (int32_t)h0; T.on = tally_on && ctx->n_tallies > 0;
    T.tab_key = ctx->d_tab_key.p; T.tab_val = ctx->d_tab_val.p; T.tab_list = ctx->d_tab_list.p;
    T.tab_mask = ctx->tab_size ? ctx->tab_size - 1 : 0; T.n_tallies = (int32_t)ctx->n_tallies; T.direct = ctx->tab_direct ? 1 : 0;
    T.sum = ctx->d_tally_sum.p; T.squared = ctx->d_tally_sq.p;
    T.dense = ctx->d_dense.p; T.dense_pending = ctx->d_dense_pending.p; T.dense_rows = ctx->dense_rows;
    T.dense_cursor = ctx->d_dense_pending.p ? ctx->d_dense_pending.p + ctx->dense_rows : nullptr;
    return T;
}

// launch errors of the stage kernels (an invalid configuration would otherwise pass silently) and the device-side
// error flags, read once per batch from the counters
static int check_batch(mcb_ctx* ctx, const Counters& hc)
{
    const cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return ctx->fail(MCB_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(le));
    if (hc.lost) return ctx->fail(MCB_ERR_LOST, "[WARNING] A particle is lost:\n( x, y, z )  (%g, %g, %g )", hc.lost_pos[0], hc.lost_pos[1], hc.lost_pos[2]);
    if (hc.overflow_sites) return ctx->fail(MCB_ERR_CAPACITY, "fission bank overflow: more than %llu sites on rank %d (raise mcb_config.site_capacity)", (unsigned long long)ctx->site_cap, ctx->rank);
    if (hc.overflow_slots) return ctx->fail(MCB_ERR_CAPACITY, "particle bank overflow: more than %u slots (raise mcb_config.bank_capacity)", ctx->n_slots);
    if (hc.overflow_stack) return ctx->fail(MCB_ERR_CAPACITY, "secondary stack overflow: a history had more than %d particles waiting, or the block's spare chunks ran out", ctx->plan.stack_max);
    if (hc.hang) return ctx->fail(MCB_ERR_CUDA, "walk kernel: a bounded wait ran out (code %d live %lld: %s; src_ready %llu, walk_head %llu)", hc.hang, hc.live, hc.hang == 4 ? "source sweep beside the walk" : hc.hang == 3 ? "idle warps waiting for shared work" : "ring of handed-over secondaries", hc.src_ready, hc.walk_head);
    if (hc.overflow_tally) return ctx->fail(MCB_ERR_CAPACITY, "tally table overflow: a history touched more than %u tally bins", ctx->tab_size);
    return MCB_OK;
}

static mcbk::WalkSource walk_source(mcb_ctx* ctx, const SourceBankView* V, int32_t first_hist, uint64_t nps0, const mcbk::SortScratch* sort)
{
    mcbk::WalkSource S;
    memset(&S, 0, sizeof(S));
    S.first_hist = first_hist; S.nps0 = nps0; S.seed0 = ctx->P.seed0;
    if (V && V->n) {
        S.fused = 1; S.V = *V;
        if (sort) { S.sorted_key = sort->key_out; S.sorted_val = sort->val_out; S.rng_after = sort->rng_after; S.rot = sort->rot; S.key32 = sort->key32; }
    }
    return S;
}

static int ensure_sort_scratch(mcb_ctx* ctx)
{
    if (ctx->d_sort_key.p) return MCB_OK;
    const size_t nbh = ctx->batch_hist;
    CK(ctx->d_sort_key.alloc(2 * nbh)); CK(ctx->d_sort_val.alloc(2 * nbh)); CK(ctx->d_sort_rng.alloc(nbh));
    CK(ctx->d_sort_temp.alloc(mcbk::sort_temp_bytes((uint32_t)nbh) + 256));
    ctx->sort.key_in = ctx->d_sort_key.p; ctx->sort.key_out = ctx->d_sort_key.p + nbh;
    ctx->sort.val_in = ctx->d_sort_val.p; ctx->sort.val_out = ctx->d_sort_val.p + nbh;
    ctx->sort.rng_after = ctx->d_sort_rng.p; ctx->sort.temp = ctx->d_sort_temp.p; ctx->sort.temp_bytes = ctx->d_sort_temp.n;
    return MCB_OK;
}

// One generation whose source bank is still on the host (mcb_run_cycle_host): the draws are made and sorted by site
// index first; the bank then comes in over PCIe in N_CHUNK pieces on a copy stream, and as soon as piece c is there
// the histories that drew from it are sourced and walked — the walk of piece c runs under the copy of piece c+1.
static int transport_streamed(mcb_ctx* ctx, uint32_t nb, bool tally_on, int* n_iterations)
{
    cudaStream_t st = ctx->stream, cs = ctx->copy_stream;
    const DevProblem& P = ctx->P;
    const uint64_t n = ctx->streamed->n;
    constexpr int NC = mcb_ctx::N_CHUNK;
    const TallyAcc T = tally_acc(ctx, 0, tally_on);
    const uint64_t nps0 = ctx->icycle * ctx->n_sample + ctx->shard_begin;
    Counters* C = ctx->d_counters.p;
    Site* dst = ctx->d_local_bank[0].p;
    SourceBankView V;
    memset(&V, 0, sizeof(V));
    V.flat = dst; V.dir_x = ctx->d_host_dirs.p; V.n = n;
    { const int rc = ensure_sort_scratch(ctx); if (rc != MCB_OK) return rc; }
    ctx->sort.rot = 0;
    ctx->sort.key32 = n < (1ull << 32) ? 1 : 0;
    // the copies start right away ...
    unsigned long long lo[NC + 1];
    for (int c = 0; c <= NC; c++) lo[c] = n * (unsigned long long)c / NC;
    for (int c = 0; c < NC; c++) {
        const size_t cnt = (size_t)(lo[c + 1] - lo[c]);
        if (cnt) {
            CK(cudaMemcpyAsync(ctx->d_io_sites.p + 8 * lo[c], ctx->streamed->sites8 + 8 * lo[c], cnt * 8 * sizeof(double), cudaMemcpyHostToDevice, cs));
            CK(cudaMemcpyAsync(ctx->d_io_cells.p + lo[c], ctx->streamed->cells + lo[c], cnt * sizeof(int32_t), cudaMemcpyHostToDevice, cs));
        }
        CK(cudaEventRecord(ctx->ev_chunk[c], cs));
    }
    // ... while the draws are made, sorted, and cut where the site index crosses a chunk boundary
    ctx->timer.begin(st, ST_SOURCE);
    mcbk::pick_sort(st, P, 0, nb, nps0, n, &ctx->sort);
    CK(cudaMemcpyAsync(ctx->d_chunk_lo.p, lo, (NC + 1) * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
    mcbk::chunk_bounds(st, &ctx->sort, nb, ctx->d_chunk_lo.p, NC + 1, ctx->d_chunk_pos.p);
    ctx->timer.end(st);
    CK(cudaMemcpyAsync(ctx->h_chunk_pos, ctx->d_chunk_pos.p, (NC + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    trace_mark("draws sorted, chunk bounds known");
    const int max_nuc = ctx->mat_n_nuc.empty() ? 0 : *std::max_element(ctx->mat_n_nuc.begin(), ctx->mat_n_nuc.end());
    (void)max_nuc;
    for (int c = 0; c < NC; c++) {
        const uint32_t q0 = ctx->h_chunk_pos[c], q1 = c + 1 < NC ? ctx->h_chunk_pos[c + 1] : nb;
        const size_t cnt = (size_t)(lo[c + 1] - lo[c]);
        CK(cudaStreamWaitEvent(st, ctx->ev_chunk[c], 0));
        ctx->timer.begin(st, ST_SOURCE);
        if (cnt) mcbk::pack_sites(st, ctx->d_io_sites.p + 8 * lo[c], ctx->d_io_cells.p + lo[c], cnt, dst + lo[c], ctx->d_host_dirs.p + 3 * lo[c]);
        ctx->timer.end(st);
        if (q1 > q0) {
            CK(cudaMemsetAsync(&C->walk_head, 0, sizeof(unsigned long long), st));
            ctx->timer.begin(st, ST_STEP);
            mcbk::walk(st, P, ctx->B, q0, q1, C, ctx->H, T, ctx->d_site_reqs.p, ctx->site_cap, ctx->k, ctx->plan, ctx->d_stack.p, ctx->d_chunk_tab.p, ctx->d_donq.p, ctx->d_gstate.p,
                       walk_source(ctx, &V, 0, nps0, &ctx->sort));
            ctx->timer.end(st);
            (*n_iterations)++;
        }
    }
    CK(cudaMemcpyAsync(ctx->h_counters, C, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    trace_mark("all chunks uploaded, sourced and walked");
    return check_batch(ctx, *ctx->h_counters);
}

// the event loop over one batch of histories [h0, h0+nb) of the shard
static int transport_batch(mcb_ctx* ctx, uint32_t h0, uint32_t nb, bool tally_on, int* n_iterations)
{
    cudaStream_t st = ctx->stream;
    const DevProblem& P = ctx->P;
    const TallyAcc T = tally_acc(ctx, h0, tally_on);
    const uint64_t nps0 = ctx->icycle * ctx->n_sample + ctx->shard_begin;
    SourceBankView V = ctx->view;
    if (!ctx->source_is_bank) { memset(&V, 0, sizeof(V)); }
    Counters* C = ctx->d_counters.p;
    uint32_t* queue[2] = {ctx->q_active, ctx->q_next};
    ctx->timer.begin(st, ST_SOURCE);
    const mcbk::SortScratch* sort = nullptr;
    if (V.n && (!V.flat || getenv("MCB_FORCE_SORT"))) {  // bank spread over the ranks: read it in ascending order
        { const int rc = ensure_sort_scratch(ctx); if (rc != MCB_OK) return rc; }
        ctx->sort.rot = V.flat ? 0 : V.prefix[std::min(ctx->rank, V.n_seg - 1)];
        sort = &ctx->sort;
    }
    if (sort) ctx->sort.key32 = V.n < (1ull << 32) ? 1 : 0;
    // the walk kernel sources its lanes itself from a fission bank (only the draws are made and sorted beforehand when the
    // bank is spread over several GPUs); the deck's <sources> of the first cycle / of fixed-source decks are sampled by k_source
    // (measured on one B200: with a flat local bank k_source + k_walk take 0.42 + 5.67 ms, the fused form 6.36 ms — the site
    // reads sit on the critical path of the refilled lanes; across GPUs the fused form takes the NVLink sweep off the
    // critical path instead.  MCB_FUSED_SOURCE=0/1 overrides.)
    const bool fused_default = false;
    const bool fused = ctx->walk_mode && V.n > 0 && (getenv("MCB_FUSED_SOURCE") ? atoi(getenv("MCB_FUSED_SOURCE")) != 0 : fused_default);
    // Bank spread over several GPUs: the sorted sweep over the peers' banks runs piece by piece on a side stream BESIDE one
    // walk launch that vacates a few SMs and takes bank positions as they are published (k_publish).  Measured on 8 x B200:
    // 9.07 -> 8.52 ms per generation with 16 SMs left to the sweep (24: 8.80); on 2 GPUs the sweep is short and the walk's
    // loss of SMs eats the gain (8.33 -> 8.43), so it is on from 4 ranks.  The sweep needs that many SMs because a peer
    // read keeps a thread waiting for microseconds: bytes in flight per SM, not NVLink, bound it there.
    const int overlap_sms = getenv("MCB_SOURCE_OVERLAP_SMS") ? atoi(getenv("MCB_SOURCE_OVERLAP_SMS")) : (ctx->world >= 4 ? 16 : 0);
    const bool piecewise = ctx->walk_mode && !fused && sort && (!V.flat || getenv("MCB_FORCE_SORT")) && overlap_sms > 0 && nb >= (1u << 16);
    if (fused) { if (sort) mcbk::pick_sort(st, P, (int32_t)h0, nb, nps0, V.n, sort); }
    else if (piecewise) mcbk::pick_sort(st, P, (int32_t)h0, nb, nps0, V.n, sort);
    else mcbk::source(st, P, ctx->B, queue[0], (int32_t)h0, nb, nps0, V, C, sort);
    ctx->timer.end(st);

    if (piecewise) {
        // the sweep runs on a side stream, piece by piece, and publishes how far it has got; the walk kernel, launched
        // first and a few block slots short of the machine, takes bank positions as they become ready
        constexpr int NC = mcb_ctx::N_CHUNK;
        if (!ctx->side_stream) {
            // highest priority: its blocks go first wherever an SM has room, and priority streams get a hardware queue of
            // their own (two streams that share a queue run one after the other, whatever their dependencies say)
            int prio_lo = 0, prio_hi = 0;
            CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CK(cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, prio_hi));
            for (int i = 0; i <= NC; i++) CK(cudaEventCreateWithFlags(&ctx->ev_side[i], cudaEventDisableTiming));
        }
        if (getenv("MCB_DEBUG_OVERLAP") && atoi(getenv("MCB_DEBUG_OVERLAP")) == 1)  // debug: whole sweep first, only the publishing runs beside
            mcbk::source_sorted_range(st, P, ctx->B, nullptr, (int32_t)h0, 0, nb, nps0, V, C, sort);
        CK(cudaMemsetAsync(&C->walk_head, 0, sizeof(unsigned long long), st));
        CK(cudaMemsetAsync(&C->src_ready, 0, sizeof(unsigned long long), st));
        CK(cudaEventRecord(ctx->ev_side[NC], st));
        ctx->plan.reserve_sms = overlap_sms;
        mcbk::WalkSource ws = walk_source(ctx, nullptr, (int32_t)h0, nps0, nullptr);
        ws.ready = &C->src_ready;
        ctx->timer.begin(st, ST_STEP);
        mcbk::walk(st, P, ctx->B, 0, nb, C, ctx->H, T, ctx->d_site_reqs.p, ctx->site_cap, ctx->k, ctx->plan, ctx->d_stack.p, ctx->d_chunk_tab.p, ctx->d_donq.p, ctx->d_gstate.p, ws);
        ctx->timer.end(st);
        ctx->plan.reserve_sms = 0;
        *n_iterations += 1;
        CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_side[NC], 0));
        const int dbg = getenv("MCB_DEBUG_OVERLAP") ? atoi(getenv("MCB_DEBUG_OVERLAP")) : 0;
        for (int c = 0; c < NC; c++) {
            const uint32_t q0 = (uint32_t)((uint64_t)nb * c / NC), q1 = (uint32_t)((uint64_t)nb * (c + 1) / NC);
            if (dbg == 2) mcbk::source_sweep(ctx->side_stream, ctx->B, (int32_t)h0, q0, q1 - q0, V, sort);  // several sites per thread (measured slower)
            else if (dbg != 1) mcbk::source_sorted_range(ctx->side_stream, P, ctx->B, nullptr, (int32_t)h0, q0, q1 - q0, nps0, V, C, sort);
            mcbk::publish(ctx->side_stream, &C->src_ready, q1);
        }
        CK(cudaEventRecord(ctx->ev_side[0], ctx->side_stream));
        CK(cudaStreamWaitEvent(st, ctx->ev_side[0], 0));
        CK(cudaMemcpyAsync(ctx->h_counters, C, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return check_batch(ctx, *ctx->h_counters);
    }

    if (ctx->walk_mode) {
        // one launch follows the source particles in slots [0, nb) and every secondary of their histories to the end
        CK(cudaMemsetAsync(&C->walk_head, 0, sizeof(unsigned long long), st));
        if (ctx->d_dense_pending.p) CK(cudaMemsetAsync(ctx->d_dense_pending.p + ctx->dense_rows, 0, sizeof(int32_t), st));
        ctx->timer.begin(st, ST_STEP);
        mcbk::walk(st, P, ctx->B, 0, nb, C, ctx->H, T, ctx->d_site_reqs.p, ctx->site_cap, ctx->k, ctx->plan, ctx->d_stack.p, ctx->d_chunk_tab.p, ctx->d_donq.p, ctx->d_gstate.p,
                   walk_source(ctx, fused ? &V : nullptr, (int32_t)h0, nps0, sort));
        ctx->timer.end(st);
        *n_iterations += 1;
        CK(cudaMemcpyAsync(ctx->h_counters, C, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (getenv("MCB_TRACE_SHARING")) fprintf(stderr, "[mcb sharing] donated %llu  shared histories %llu  refused %llu  idle waits %llu  tracks %llu\n", ctx->h_counters->n_donated, ctx->h_counters->n_shared_hist, ctx->h_counters->n_donate_refused, ctx->h_counters->n_idle_waits, ctx->h_counters->n_tracks);
        return check_batch(ctx, *ctx->h_counters);
    }

    // Event loop.  Queue lengths live on the device; the host launches iterations ahead of what it knows and
    // learns the lengths from asynchronous copies into a pinned ring (an iteration on an empty queue is a no-op).
    const int R = MCB_RING;
    uint64_t known_n = nb;  // last length seen (sizes the grids)
    int it = 0, seen = 0;
    bool tail = nb <= ctx->finish_below;
    const uint64_t finish_below = ctx->finish_below;
    while (!tail) {
        const int cur = it % 3, nxt = (it + 1) % 3;
        uint32_t* q_in = queue[it & 1];
        uint32_t* q_out = queue[(it + 1) & 1];
        {
            ctx->timer.begin(st, ST_LOOKUP);
            mcbk::xs_stage(st, P, ctx->B, q_in, cur, known_n, C);
            ctx->timer.end(st);
            ctx->timer.begin(st, ST_FLIGHT);
            mcbk::flight(st, P, ctx->B, q_in, cur, known_n, ctx->q_ev, C, ctx->H, T);
            ctx->timer.end(st);
            ctx->timer.begin(st, ST_COLLIDE);
            mcbk::collide(st, P, ctx->B, ctx->q_ev, cur, known_n, C, q_out, ctx->H, T, ctx->d_site_reqs.p,
                          ctx->site_cap, ctx->n_slots, ctx->k);
            ctx->timer.end(st);
            ctx->timer.begin(st, ST_CROSS);
            mcbk::cross(st, P, ctx->B, ctx->q_ev, cur, known_n, C, q_out, T, ctx->n_slots);
            ctx->timer.end(st);
        }
        CK(cudaMemcpyAsync(&ctx->h_ring[it % R], &C->n_active[nxt], sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(ctx->ev_ring[it % R], st));
        it++;
        while (seen < it) {
            if (it - seen >= R - 1) CK(cudaEventSynchronize(ctx->ev_ring[seen % R]));
            else {
                const cudaError_t q = cudaEventQuery(ctx->ev_ring[seen % R]);
                if (q == cudaErrorNotReady) break;
                if (q != cudaSuccess) return ctx->fail(MCB_ERR_CUDA, "event loop: %s", cudaGetErrorString(q));
            }
            known_n = ctx->h_ring[seen % R];
            seen++;
            if (known_n <= finish_below) { tail = true; break; }
            // in problems with secondaries the queue can grow again: keep the grid hint generous
            if (P.shared_histories) known_n = std::max<uint64_t>(2 * known_n, 4096);
        }
    }
    // drain: everything launched so far has to finish before the tail kernel picks up the current queue
    CK(cudaStreamSynchronize(st));
    uint64_t n_left = it ? ctx->h_ring[(it - 1) % R] : nb;
    while (n_left > 0) {
        const int cur = it % 3, nxt = (it + 1) % 3;
        CK(cudaMemsetAsync(&C->n_active[nxt], 0, sizeof(unsigned long long), st));
        ctx->timer.begin(st, ST_FINISH);
        mcbk::finish(st, P, ctx->B, queue[it & 1], cur, n_left, C, queue[(it + 1) & 1], ctx->H, T, ctx->d_site_reqs.p,
                     ctx->site_cap, ctx->n_slots, ctx->k);
        ctx->timer.end(st);
        CK(cudaMemcpyAsync(&ctx->h_ring[0], &C->n_active[nxt], sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        n_left = ctx->h_ring[0];
        it++;
    }
    *n_iterations += it;
    CK(cudaMemcpyAsync(ctx->h_counters, C, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    { const int rc = check_batch(ctx, *ctx->h_counters); if (rc != MCB_OK) return rc; }
    if (T.on) {
        ctx->timer.begin(st, ST_CLOSEOUT);
        mcbk::tally_reduce(st, ctx->d_tally_acc.p, T.stride, nb, ctx->n_tallies, ctx->d_tally_partial.p, ctx->d_tally_sum.p, ctx->d_tally_sq.p);
        ctx->timer.end(st);
        CK(cudaGetLastError());
    }
    return MCB_OK;
}

int mcb_run_cycle(mcb_ctx* ctx, mcb_cycle_result* out)
{
    if (!ctx) return MCB_ERR_ARG;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const bool tally_on = ctx->icycle >= ctx->n_passive;  // handler.cpp:15
    const uint64_t launches0 = mcbk::launch_count();
    if (ctx->source_is_bank && ctx->view.n == 0 && !ctx->streamed) return ctx->fail(MCB_ERR_ARG, "[ERROR] Source bank is empty...");
    if (ctx->world > 1 && !ctx->comm) return ctx->fail(MCB_ERR_COMM, "world > 1 but mcb_comm_init was not called");
    Counters* C = ctx->d_counters.p;
    CK(cudaEventRecord(ctx->ev0, st));
    CK(cudaMemsetAsync(C, 0, sizeof(Counters), st));
    const size_t nh = std::max<uint64_t>(ctx->shard_count, 1);
    CK(cudaMemsetAsync(ctx->d_hist_k.p, 0, 2 * nh * sizeof(double), st));
    CK(cudaMemsetAsync(ctx->d_nsite.p, 0, nh * sizeof(int32_t), st));
    int iterations = 0;
    if (ctx->streamed) {
        const int rc = transport_streamed(ctx, (uint32_t)ctx->shard_count, tally_on, &iterations);
        if (rc != MCB_OK) return rc;
    } else
    for (uint64_t h0 = 0; h0 < ctx->shard_count; h0 += ctx->batch_hist) {
        const uint32_t nb = (uint32_t)std::min<uint64_t>(ctx->batch_hist, ctx->shard_count - h0);
        const int rc = transport_batch(ctx, (uint32_t)h0, nb, tally_on, &iterations);
        if (rc != MCB_OK) return rc;
    }
    // ---- close-out on this rank ----
    ctx->timer.begin(st, ST_CLOSEOUT);
    if (ctx->ksearch) mcbk::reduce_k(st, ctx->H.kC, ctx->H.kTL, (uint32_t)ctx->shard_count, C);
    ctx->timer.end(st);
    CK(cudaMemcpyAsync(ctx->h_counters, C, sizeof(Counters), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const uint64_t n_local = ctx->h_counters->site_cursor;
    // with world > 1 the banks alternate: peers may still be drawing their sources from the one written last cycle
    const int bw = ctx->world > 1 ? ctx->bank_w : 0;
    Site* const wbank = ctx->d_local_bank[bw].p;
    if (ctx->ksearch) {
        ctx->timer.begin(st, ST_BANK);
        mcbk::scan_sites(st, ctx->d_scan_temp.p, ctx->d_scan_temp.n, ctx->H.nsite, ctx->d_site_offset.p, (uint32_t)ctx->shard_count);
        mcbk::bank_sample_order(st, ctx->P, ctx->d_site_reqs.p, n_local, ctx->d_site_offset.p, wbank);
        ctx->timer.end(st);
        if (ctx->entropy_on) {
            ctx->timer.begin(st, ST_CLOSEOUT);
            mcbk::entropy_history(st, ctx->P, wbank, ctx->d_site_offset.p, ctx->H.nsite, (uint32_t)ctx->shard_count, C);
            CK(cudaMemsetAsync(ctx->d_entropy_bins.p, 0, ctx->d_entropy_bins.n * sizeof(unsigned long long), st));
            mcbk::entropy_histogram(st, ctx->P, wbank, n_local, ctx->d_entropy_bins.p);
            ctx->timer.end(st);
        }
    }
    ctx->n_local_sites = n_local;
    ctx->bank_last = bw;
    CK(cudaEventRecord(ctx->ev1, st));

    // ---- gather the sums: [0..9] fixed-point limbs, [10..15] counters, then tally sum / squared (as doubles) ----
    const size_t nt = (size_t)std::max<int64_t>(ctx->n_tallies, 0);
    std::vector<unsigned long long> red(16 + ctx->d_entropy_bins.n, 0);
    std::vector<double> tsum(nt, 0.0), tsq(nt, 0.0);
    uint64_t n_global_sites = n_local;
    {
        CK(cudaMemcpyAsync(ctx->h_counters, C, sizeof(Counters), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        const Counters& hc = *ctx->h_counters;
        if (hc.overflow_fixed) return ctx->fail(MCB_ERR_ARG, "a per-history k score left the fixed-point range");
        for (int j = 0; j < 5; j++) { red[2 * j] = hc.fx_lo[j]; red[2 * j + 1] = hc.fx_hi[j]; }
        red[10] = n_local; red[11] = hc.n_tracks; red[12] = hc.n_collisions; red[13] = hc.n_lookups; red[14] = hc.n_crossings;
        red[15] = ctx->shard_count;
        if (ctx->d_entropy_bins.n) CK(cudaMemcpyAsync(red.data() + 16, ctx->d_entropy_bins.p, ctx->d_entropy_bins.n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        if (nt && tally_on) {
            CK(cudaMemcpyAsync(tsum.data(), ctx->d_tally_sum.p, nt * sizeof(double), cudaMemcpyDeviceToHost, st));
            CK(cudaMemcpyAsync(tsq.data(), ctx->d_tally_sq.p, nt * sizeof(double), cudaMemcpyDeviceToHost, st));
            CK(cudaMemsetAsync(ctx->d_tally_sum.p, 0, nt * sizeof(double), st));
            CK(cudaMemsetAsync(ctx->d_tally_sq.p, 0, nt * sizeof(double), st));
        }
        CK(cudaStreamSynchronize(st));
    }
    SourceBankView next_view;
    memset(&next_view, 0, sizeof(next_view));
    next_view.flat = wbank; next_view.n = n_local;
    if (ctx->world > 1) {
        // One all-gather of every rank's close-out vector: [0..9] fixed-point limbs, [10] sites, [11..15] counters,
        // entropy histogram, tally sum / squared (double bit patterns).  Every rank adds the vectors in rank order:
        // integer sums are exact, double sums have a fixed order, so all ranks hold identical results, and the site
        // counts give the global index range of every rank's bank slice.
        const int W = ctx->world;
        const size_t VL = ctx->vec_len, nb = ctx->d_entropy_bins.n;
        unsigned long long* hs = ctx->h_send;
        memcpy(hs, red.data(), (16 + nb) * sizeof(unsigned long long));
        if (nt) { memcpy(hs + 16 + nb, tsum.data(), nt * sizeof(double)); memcpy(hs + 16 + nb + nt, tsq.data(), nt * sizeof(double)); }
        CK(cudaMemcpyAsync(ctx->d_send.p, hs, VL * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
        NK(g_nccl.AllGather(ctx->d_send.p, ctx->d_recv.p, VL, ncclUint64, ctx->comm, st));
        CK(cudaMemcpyAsync(ctx->h_recv, ctx->d_recv.p, VL * W * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        std::vector<unsigned long long> counts(W);
        std::fill(red.begin(), red.end(), 0ull);
        std::fill(tsum.begin(), tsum.end(), 0.0); std::fill(tsq.begin(), tsq.end(), 0.0);
        for (int r = 0; r < W; r++) {
            const unsigned long long* v = ctx->h_recv + (size_t)r * VL;
            counts[r] = v[10];
            for (size_t i = 0; i < 16 + nb; i++) red[i] += v[i];
            if (nt && tally_on) {
                const double* ts = reinterpret_cast<const double*>(v + 16 + nb);
                for (size_t t = 0; t < nt; t++) { tsum[t] += ts[t]; tsq[t] += ts[nt + t]; }
            }
        }
        if (ctx->ksearch) {
            uint64_t total = 0;
            for (auto c : counts) total += c;
            n_global_sites = total;
            if (ctx->p2p) {
                // nothing moves: next cycle's source kernel reads each drawn site from the HBM of the rank that banked it
                next_view.flat = nullptr; next_view.n_seg = W; next_view.n = total;
                uint64_t off = 0;
                for (int r = 0; r < W; r++) { next_view.seg[r] = ctx->peer_bank[bw][r]; next_view.prefix[r] = off; off += counts[r]; }
                next_view.prefix[W] = off;
                ctx->bank_w ^= 1;
            } else {
                // no peer access: each rank's slice is broadcast into place (all-gather-v)
                if (total > ctx->global_cap) return ctx->fail(MCB_ERR_CAPACITY, "global fission bank overflow: %llu sites", (unsigned long long)total);
                if (ctx->d_global_bank.n < ctx->global_cap) CK(ctx->d_global_bank.alloc(ctx->global_cap));
                NK(g_nccl.GroupStart());
                uint64_t off = 0;
                for (int r = 0; r < W; r++) {
                    if (counts[r]) NK(g_nccl.Broadcast(wbank, ctx->d_global_bank.p + off, counts[r] * sizeof(Site), ncclChar, r, ctx->comm, st));
                    off += counts[r];
                }
                NK(g_nccl.GroupEnd());
                next_view.flat = ctx->d_global_bank.p; next_view.n = total;
            }
        }
    }
    CK(cudaEventRecord(ctx->ev2, st));
    CK(cudaEventSynchronize(ctx->ev2));
    {
        double ms[ST_N] = {0}; uint64_t cnt[ST_N] = {0};
        ctx->timer.resolve(ms, cnt);
        mcb_stage_times& S = ctx->stage;
        S.ms_source += ms[ST_SOURCE]; S.ms_lookup += ms[ST_LOOKUP]; S.ms_flight += ms[ST_FLIGHT]; S.ms_cross += ms[ST_CROSS];
        S.ms_collide += ms[ST_COLLIDE]; S.ms_closeout += ms[ST_CLOSEOUT]; S.ms_bank += ms[ST_BANK]; S.ms_finish += ms[ST_FINISH]; S.ms_step += ms[ST_STEP];
        S.n_source += cnt[ST_SOURCE]; S.n_lookup += cnt[ST_LOOKUP]; S.n_flight += cnt[ST_FLIGHT]; S.n_cross += cnt[ST_CROSS];
        S.n_collide += cnt[ST_COLLIDE]; S.n_closeout += cnt[ST_CLOSEOUT]; S.n_bank += cnt[ST_BANK]; S.n_finish += cnt[ST_FINISH]; S.n_step += cnt[ST_STEP];
        S.units_lookup += ctx->h_counters->n_lookups;
    }

    // ---- cycle close-out on the host (identical on every rank) ----
    mcb_cycle_result r;
    memset(&r, 0, sizeof(r));
    const double Ns = (double)ctx->n_sample;
    if (tally_on && nt) {  // Estimator::end_cycle (Estimator.cpp:347-360)
        for (size_t t = 0; t < nt; t++) {
            const double mean = tsum[t] / Ns;
            const double uncer_squared = (tsq[t] / Ns - mean * mean) / (Ns - 1.0);
            ctx->tally_mean[t] += mean;
            ctx->tally_uncer[t] += uncer_squared;
        }
    }
    r.k_sum_C = fx_to_double(red[0], red[1]); r.k_sum_TL = fx_to_double(red[2], red[3]);
    r.k_sq_C = fx_to_double(red[4], red[5]); r.k_sq_TL = fx_to_double(red[6], red[7]);
    const double H_sum = fx_to_double(red[8], red[9]);
    if (ctx->ksearch) {  // EstimatorK::report_cycle (Estimator.cpp:526-561)
        const double mean_C = r.k_sum_C / Ns, mean_TL = r.k_sum_TL / Ns;
        const double mean = (mean_C + mean_TL) / 2;
        r.H = H_sum / Ns;
        r.k_cycle = mean;
        ctx->k = mean;
        if (tally_on) {
            const double us_C = (r.k_sq_C / Ns - mean_C * mean_C) / (Ns - 1.0);
            const double us_TL = (r.k_sq_TL / Ns - mean_TL * mean_TL) / (Ns - 1.0);
            ctx->Navg++;
            ctx->mean_accumulator += mean;
            ctx->uncer_sq_accumulator += us_C + us_TL;
            r.k_avg = ctx->mean_accumulator / ctx->Navg;
            r.k_uncer = std::sqrt(ctx->uncer_sq_accumulator) / ctx->Navg / 2;
        }
        if (ctx->d_entropy_bins.n) {
            unsigned long long tot = 0;
            for (size_t b = 0; b < ctx->d_entropy_bins.n; b++) tot += red[16 + b];
            double Hc = 0.0;
            for (size_t b = 0; b < ctx->d_entropy_bins.n; b++) if (red[16 + b]) { const double pb = (double)red[16 + b] / (double)tot; Hc -= pb * std::log2(pb); }
            r.H_cycle_conventional = Hc;
        }
        ctx->view = next_view;
        ctx->source_is_bank = true;
    }
    r.n_sites = ctx->ksearch ? n_global_sites : 0;
    r.n_tracks = red[11]; r.n_collisions = red[12]; r.n_lookups = red[13]; r.n_crossings = red[14]; r.n_histories = red[15];
    float ms_t = 0, ms_x = 0;
    cudaEventElapsedTime(&ms_t, ctx->ev0, ctx->ev1);
    cudaEventElapsedTime(&ms_x, ctx->ev1, ctx->ev2);
    r.ms_transport = ms_t; r.ms_exchange = ms_x;
    r.n_iterations = iterations;
    r.n_kernel_launches = mcbk::launch_count() - launches0;
    r.lost = 0;
    ctx->icycle++;
    if (out) *out = r;
    return MCB_OK;
}

int mcb_get_tallies(mcb_ctx* ctx, double* mean, double* uncer, int64_t n)
{
    if (!ctx) return MCB_ERR_ARG;
    if (n > ctx->n_tallies) n = ctx->n_tallies;
    // Estimator::end_simulation (Estimator.cpp:361-367); the stored accumulators are left untouched
    uint64_t n_active = ctx->icycle > ctx->n_passive ? ctx->icycle - ctx->n_passive : 0;
    if (n_active == 0) n_active = 1;
    const double Nactive = (double)n_active;
    for (int64_t t = 0; t < n; t++) {
        if (mean) mean[t] = ctx->tally_mean[t] / Nactive;
        if (uncer) uncer[t] = std::sqrt(ctx->tally_uncer[t]) / Nactive;
    }
    return MCB_OK;
}

int mcb_get_stage_times(mcb_ctx* ctx, mcb_stage_times* out)
{
    if (!ctx || !out) return MCB_ERR_ARG;
    *out = ctx->stage;
    return MCB_OK;
}
void mcb_reset_stage_times(mcb_ctx* ctx)
{
    if (ctx) memset(&ctx->stage, 0, sizeof(ctx->stage));
}
void mcb_set_stage_timing(mcb_ctx* ctx, int on)
{
    if (ctx) ctx->timer.on = on != 0;
}

static int64_t read_bank(mcb_ctx* ctx, const Site* bank, const double* dir_x, uint64_t have, double* out, int32_t* cells, int64_t max_n)
{
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MCB_ERR_CUDA;
    const int64_t n = std::min<int64_t>((int64_t)have, max_n);
    if (n <= 0) return 0;
    // Site records -> the host-facing layout on the device, then straight into the caller's buffers
    if (ctx->d_io_sites.n < (size_t)n * 8) {  // grow with headroom: the bank size wanders from generation to generation
        const size_t cap = (size_t)n + (size_t)n / 4 + 1024;
        if (ctx->d_io_sites.alloc(cap * 8) != cudaSuccess || ctx->d_io_cells.alloc(cap) != cudaSuccess)
            return ctx->fail(MCB_ERR_CUDA, "out of device memory for the bank staging buffers");
    }
    mcbk::unpack_sites(ctx->stream, bank, dir_x, (uint64_t)n, ctx->d_io_sites.p, ctx->d_io_cells.p);
    cudaError_t e = cudaSuccess;
    if (out) e = cudaMemcpyAsync(out, ctx->d_io_sites.p, (size_t)n * 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && cells) e = cudaMemcpyAsync(cells, ctx->d_io_cells.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return ctx->fail(MCB_ERR_CUDA, "fission bank read-back failed: %s", cudaGetErrorString(e));
    return n;
}

int64_t mcb_get_fission_bank(mcb_ctx* ctx, double* out, int32_t* cells, int64_t max_n)
{
    if (!ctx) return MCB_ERR_ARG;
    return read_bank(ctx, ctx->d_local_bank[ctx->bank_last].p, nullptr, ctx->n_local_sites, out, cells, max_n);
}

int64_t mcb_get_source_bank(mcb_ctx* ctx, double* out, int32_t* cells, int64_t max_n)
{
    if (!ctx) return MCB_ERR_ARG;
    if (!ctx->source_is_bank) return 0;
    if (ctx->view.flat) return read_bank(ctx, ctx->view.flat, ctx->view.dir_x, ctx->view.n, out, cells, max_n);
    // segmented bank (peer slices): materialise the part asked for
    const int64_t n = std::min<int64_t>((int64_t)ctx->view.n, max_n);
    if (n <= 0) return 0;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MCB_ERR_CUDA;
    if (ctx->d_global_bank.n < (size_t)n && ctx->d_global_bank.alloc((size_t)n) != cudaSuccess)
        return ctx->fail(MCB_ERR_CUDA, "out of device memory for the gathered bank");
    mcbk::gather_sites(ctx->stream, ctx->view, 0, (uint64_t)n, ctx->d_global_bank.p);
    return read_bank(ctx, ctx->d_global_bank.p, nullptr, (uint64_t)n, out, cells, max_n);
}

int mcb_set_source_bank(mcb_ctx* ctx, const double* sites8, const int32_t* cells, int64_t n)
{
    if (!ctx || n < 0) return MCB_ERR_ARG;
    if (!ctx->ksearch) return ctx->fail(MCB_ERR_ARG, "mcb_set_source_bank: not a k-eigenvalue problem");
    CK(cudaSetDevice(ctx->device));
    const uint64_t cap = ctx->world > 1 ? ctx->global_cap : ctx->site_cap;
    if ((uint64_t)n > cap) return ctx->fail(MCB_ERR_CAPACITY, "source bank of %lld sites exceeds the capacity %llu", (long long)n, (unsigned long long)cap);
    Site* dst = ctx->d_local_bank[0].p;
    if (ctx->world > 1) {
        if (ctx->d_global_bank.n < (size_t)std::max<int64_t>(n, 1)) CK(ctx->d_global_bank.alloc((size_t)std::max<int64_t>(n, 1)));
        dst = ctx->d_global_bank.p;
    }
    if (ctx->d_io_sites.n < (size_t)n * 8) {
        const size_t cap = (size_t)n + (size_t)n / 4 + 1024;
        CK(ctx->d_io_sites.alloc(cap * 8)); CK(ctx->d_io_cells.alloc(cap));
    }
    if (n) {
        CK(cudaMemcpyAsync(ctx->d_io_sites.p, sites8, (size_t)n * 8 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpyAsync(ctx->d_io_cells.p, cells, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        if (ctx->d_host_dirs.n < (size_t)n * 3) CK(ctx->d_host_dirs.alloc(((size_t)n + (size_t)n / 4 + 1024) * 3));
        mcbk::pack_sites(ctx->stream, ctx->d_io_sites.p, ctx->d_io_cells.p, (uint64_t)n, dst, ctx->d_host_dirs.p);
    }
    CK(cudaStreamSynchronize(ctx->stream));
    memset(&ctx->view, 0, sizeof(ctx->view));
    ctx->view.flat = dst; ctx->view.dir_x = n ? ctx->d_host_dirs.p : nullptr; ctx->view.n = (uint64_t)n;
    ctx->source_is_bank = true;
    return MCB_OK;
}

int mcb_run_cycle_host(mcb_ctx* ctx, const double* in_sites8, const int32_t* in_cells, int64_t n_in, double* out_sites8,
                       int32_t* out_cells, int64_t max_out, int64_t* n_out, mcb_cycle_result* out)
{
    if (!ctx || n_in < 0 || !n_out) return MCB_ERR_ARG;
    if (!ctx->ksearch) return ctx->fail(MCB_ERR_ARG, "mcb_run_cycle_host: not a k-eigenvalue problem");
    if (n_in == 0) return ctx->fail(MCB_ERR_ARG, "[ERROR] Source bank is empty...");
    CK(cudaSetDevice(ctx->device));
    // the pipelined form needs one rank, one particle per history, one batch, the walk kernel; otherwise the three
    // plain calls do the same thing one after the other
    const bool pipelined = ctx->world == 1 && !ctx->P.shared_histories && ctx->walk_mode && ctx->shard_count <= ctx->batch_hist &&
                           (uint64_t)n_in <= ctx->site_cap && !getenv("MCB_NO_STREAM");
    if (ctx->world > 1 && ctx->p2p && ctx->walk_mode && !getenv("MCB_NO_SLICES")) {
        // Several GPUs: the host bank is one array of which every rank owns a slice.  Rank r uploads sites
        // [n r / W, n (r + 1) / W) into its own (peer-mapped) bank buffer, the generation reads the slices in place over
        // NVLink like it reads the banks it made itself, and rank r writes its canonical slice of the new bank back at
        // its global offset: per rank 1/W of the bytes each way, whatever W.
        const int W = ctx->world;
        const int src = ctx->bank_w ^ 1;  // the buffer the running cycle does not write
        uint64_t a = 0, cnt = 0;
        mcb_shard_range((uint64_t)n_in, ctx->rank, W, &a, &cnt);
        if (cnt > ctx->site_cap) return ctx->fail(MCB_ERR_CAPACITY, "source bank slice of %llu sites exceeds the capacity %llu", (unsigned long long)cnt, (unsigned long long)ctx->site_cap);
        if (ctx->d_io_sites.n < (size_t)std::max<uint64_t>(cnt, 1) * 8) { const size_t cap = cnt + cnt / 4 + 1024; CK(ctx->d_io_sites.alloc(cap * 8)); CK(ctx->d_io_cells.alloc(cap)); }
        if (cnt) {
            CK(cudaMemcpyAsync(ctx->d_io_sites.p, in_sites8 + 8 * a, cnt * 8 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->d_io_cells.p, in_cells + a, cnt * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
            mcbk::pack_sites(ctx->stream, ctx->d_io_sites.p, ctx->d_io_cells.p, cnt, ctx->d_local_bank[src].p, ctx->d_slice_dirs.p);
        }
        // nobody reads a slice before its owner has written it
        NK(g_nccl.AllReduce(ctx->d_send.p, ctx->d_send.p, 1, ncclUint64, ncclSum, ctx->comm, ctx->stream));
        memset(&ctx->view, 0, sizeof(ctx->view));
        ctx->view.n_seg = W; ctx->view.n = (uint64_t)n_in;
        for (int r = 0; r < W; r++) {
            uint64_t b0 = 0, c0 = 0;
            mcb_shard_range((uint64_t)n_in, r, W, &b0, &c0);
            ctx->view.seg[r] = ctx->peer_bank[src][r]; ctx->view.seg_dir[r] = ctx->peer_dirs[r]; ctx->view.prefix[r] = b0;
        }
        ctx->view.prefix[W] = (uint64_t)n_in;
        ctx->source_is_bank = true;
        const int rc = mcb_run_cycle(ctx, out);
        if (rc != MCB_OK) return rc;
        // this rank's slice [n' r / W, n' (r + 1) / W) of the new bank (its own sites but for a few at the ends, which it
        // reads from its neighbours in place), written at its global offset: the slices of the next call
        const uint64_t total = ctx->view.n;
        if ((int64_t)total > max_out) return ctx->fail(MCB_ERR_CAPACITY, "the new bank has %llu sites, the output buffers hold %lld", (unsigned long long)total, (long long)max_out);
        uint64_t off = 0, mine = 0;
        mcb_shard_range(total, ctx->rank, W, &off, &mine);
        if (mine) {
            if (ctx->d_io_sites.n < (size_t)mine * 8) { const size_t cap = mine + mine / 4 + 1024; CK(ctx->d_io_sites.alloc(cap * 8)); CK(ctx->d_io_cells.alloc(cap)); }
            if (ctx->d_global_bank.n < (size_t)mine) CK(ctx->d_global_bank.alloc((size_t)mine + (size_t)mine / 4 + 1024));
            mcbk::gather_sites(ctx->stream, ctx->view, off, mine, ctx->d_global_bank.p);
            mcbk::unpack_sites(ctx->stream, ctx->d_global_bank.p, nullptr, mine, ctx->d_io_sites.p, ctx->d_io_cells.p);
            if (out_sites8) CK(cudaMemcpyAsync(out_sites8 + 8 * off, ctx->d_io_sites.p, mine * 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
            if (out_cells) CK(cudaMemcpyAsync(out_cells + off, ctx->d_io_cells.p, mine * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
        *n_out = (int64_t)total;
        return MCB_OK;
    }
    if (!pipelined) {
        int rc = mcb_set_source_bank(ctx, in_sites8, in_cells, n_in);
        if (rc == MCB_OK) rc = mcb_run_cycle(ctx, out);
        if (rc != MCB_OK) return rc;
        const int64_t got = mcb_get_source_bank(ctx, out_sites8, out_cells, max_out);
        if (got < 0) return (int)got;
        *n_out = got;
        return MCB_OK;
    }
    constexpr int NC = mcb_ctx::N_CHUNK;
    if (!ctx->copy_stream) {
        CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int c = 0; c < NC; c++) CK(cudaEventCreateWithFlags(&ctx->ev_chunk[c], cudaEventDisableTiming));
        CK(ctx->d_chunk_lo.alloc(NC + 1)); CK(ctx->d_chunk_pos.alloc(NC + 1));
        CK(cudaMallocHost((void**)&ctx->h_chunk_pos, (NC + 1) * sizeof(uint32_t)));
    }
    const size_t need = (size_t)std::max<int64_t>(n_in, 1);
    if (ctx->d_io_sites.n < need * 8) { const size_t cap = need + need / 4 + 1024; CK(ctx->d_io_sites.alloc(cap * 8)); CK(ctx->d_io_cells.alloc(cap)); }
    if (ctx->d_host_dirs.n < need * 3) CK(ctx->d_host_dirs.alloc((need + need / 4 + 1024) * 3));
    // everything queued on the launch stream so far must be done before the copy stream overwrites the staging buffers
    CK(cudaStreamSynchronize(ctx->stream));
    mcb_ctx::Streamed sd{in_sites8, in_cells, (uint64_t)n_in};
    ctx->streamed = &sd;
    trace_mark("host-bank cycle begins");
    const int rc = mcb_run_cycle(ctx, out);
    ctx->streamed = nullptr;
    if (rc != MCB_OK) return rc;
    trace_mark("generation closed out");
    // the new bank goes back in pieces as well: piece c is unpacked while piece c-1 is on the wire
    const int64_t n = std::min<int64_t>((int64_t)ctx->view.n, max_out);
    if (ctx->d_io_sites.n < (size_t)std::max<int64_t>(n, 1) * 8) {
        const size_t cap = (size_t)n + (size_t)n / 4 + 1024;
        CK(ctx->d_io_sites.alloc(cap * 8)); CK(ctx->d_io_cells.alloc(cap));
    }
    for (int c = 0; c < NC; c++) {
        const uint64_t a = (uint64_t)n * c / NC, b = (uint64_t)n * (c + 1) / NC;
        if (b == a) continue;
        mcbk::unpack_sites(ctx->stream, ctx->view.flat + a, nullptr, b - a, ctx->d_io_sites.p + 8 * a, ctx->d_io_cells.p + a);
        CK(cudaEventRecord(ctx->ev_chunk[c], ctx->stream));
        CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_chunk[c], 0));
        if (out_sites8) CK(cudaMemcpyAsync(out_sites8 + 8 * a, ctx->d_io_sites.p + 8 * a, (b - a) * 8 * sizeof(double), cudaMemcpyDeviceToHost, ctx->copy_stream));
        if (out_cells) CK(cudaMemcpyAsync(out_cells + a, ctx->d_io_cells.p + a, (b - a) * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->copy_stream));
    }
    CK(cudaStreamSynchronize(ctx->copy_stream));
    trace_mark("new bank in host memory");
    *n_out = n;
    return MCB_OK;
}

}  // extern "C"

// ---- parity / bench entry points ----
template <typename F>
static int with_buffers(mcb_ctx* ctx, F f)
{
    CK(cudaSetDevice(ctx->device));
    const int rc = f();
    if (rc != MCB_OK) return rc;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(ctx->stream));
    return MCB_OK;
}

extern "C" {
#define H2D(buf, src, count) CK((buf).alloc(count)); if (count) CK(cudaMemcpyAsync((buf).p, src, (count) * sizeof(*(buf).p), cudaMemcpyHostToDevice, ctx->stream))
#define D2H(dst, buf, count) if (count) CK(cudaMemcpyAsync(dst, (buf).p, (count) * sizeof(*(buf).p), cudaMemcpyDeviceToHost, ctx->stream))

int mcb_xs_lookup_batch(mcb_ctx* ctx, int32_t material, const double* E, int64_t n, double* out5)
{
    if (!ctx || n < 0) return MCB_ERR_ARG;
    if (material < 0 || material >= ctx->n_materials) return ctx->fail(MCB_ERR_ARG, "material %d out of range", material);
    return with_buffers(ctx, [&]() -> int {
        DevBuf<double> dE, dO;
        H2D(dE, E, (size_t)n);
        CK(dO.alloc((size_t)n * 5));
        mcbk::xs_lookup(ctx->stream, ctx->P, material, dE.p, n, dO.p);
        D2H(out5, dO, (size_t)n * 5);
        CK(cudaStreamSynchronize(ctx->stream));
        return MCB_OK;
    });
}

int mcb_xs_lookup_device(mcb_ctx* ctx, int32_t material, const double* dE, int64_t n, double* dout5, float* ms)
{
    if (!ctx || n < 0) return MCB_ERR_ARG;
    if (material < 0 || material >= ctx->n_materials) return ctx->fail(MCB_ERR_ARG, "material %d out of range", material);
    CK(cudaSetDevice(ctx->device));
    if (ms) CK(cudaEventRecord(ctx->ev0, ctx->stream));
    mcbk::xs_lookup(ctx->stream, ctx->P, material, dE, n, dout5);
    CK(cudaGetLastError());
    if (ms) {
        CK(cudaEventRecord(ctx->ev1, ctx->stream));
        CK(cudaEventSynchronize(ctx->ev1));
        CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    }
    return MCB_OK;
}

int mcb_select_channel_batch(mcb_ctx* ctx, int32_t material, int32_t kind, const double* E, const double* xi, int64_t n,
                             int32_t* nuclide)
{
    if (!ctx || n < 0) return MCB_ERR_ARG;
    if (material < 0 || material >= ctx->n_materials) return ctx->fail(MCB_ERR_ARG, "material %d out of range", material);
    return with_buffers(ctx, [&]() -> int {
        DevBuf<double> dE, dX; DevBuf<int32_t> dO;
        H2D(dE, E, (size_t)n); H2D(dX, xi, (size_t)n);
        CK(dO.alloc((size_t)n));
        mcbk::select_channel(ctx->stream, ctx->P, material, kind, dE.p, dX.p, n, dO.p);
        D2H(nuclide, dO, (size_t)n);
        CK(cudaStreamSynchronize(ctx->stream));
        return MCB_OK;
    });
}

int mcb_beta_batch(mcb_ctx* ctx, int32_t material, int32_t local_nuclide, const double* E, int64_t n, double* out)
{
    if (!ctx || n < 0) return MCB_ERR_ARG;
    if (material < 0 || material >= ctx->n_materials) return ctx->fail(MCB_ERR_ARG, "material %d out of range", material);
    if (local_nuclide < 0 || local_nuclide >= ctx->mat_n_nuc[material]) return ctx->fail(MCB_ERR_ARG, "nuclide %d out of range", local_nuclide);
    return with_buffers(ctx, [&]() -> int {
        DevBuf<double> dE, dO;
        H2D(dE, E, (size_t)n);
        CK(dO.alloc((size_t)n));
        mcbk::beta(ctx->stream, ctx->P, material, local_nuclide, dE.p, n, dO.p);
        D2H(out, dO, (size_t)n);
        CK(cudaStreamSynchronize(ctx->stream));
        return MCB_OK;
    });
}

int mcb_rng_batch(mcb_ctx* ctx, const uint64_t* nps, int64_t n, int32_t ndraw, uint64_t* seeds_out)
{
    if (!ctx || n < 0 || ndraw < 0) return MCB_ERR_ARG;
    return with_buffers(ctx, [&]() -> int {
        DevBuf<uint64_t> dN, dO;
        H2D(dN, nps, (size_t)n);
        CK(dO.alloc((size_t)n * ndraw));
        mcbk::rng(ctx->stream, ctx->seed, dN.p, n, ndraw, dO.p);
        D2H(seeds_out, dO, (size_t)n * ndraw);
        CK(cudaStreamSynchronize(ctx->stream));
        return MCB_OK;
    });
}

int mcb_geometry_batch(mcb_ctx* ctx, const int32_t* cell, const double* pos3, const double* dir3, int64_t n, double* out3)
{
    if (!ctx || n < 0) return MCB_ERR_ARG;
    for (int64_t i = 0; i < n; i++) if (cell[i] < 0 || cell[i] >= ctx->n_cells) return ctx->fail(MCB_ERR_ARG, "cell %d out of range", cell[i]);
    return with_buffers(ctx, [&]() -> int {
        DevBuf<int32_t> dC; DevBuf<double> dP, dD, dO;
        H2D(dC, cell, (size_t)n); H2D(dP, pos3, (size_t)n * 3); H2D(dD, dir3, (size_t)n * 3);
        CK(dO.alloc((size_t)n * 3));
        mcbk::geometry(ctx->stream, ctx->P, dC.p, dP.p, dD.p, n, dO.p);
        D2H(out3, dO, (size_t)n * 3);
        CK(cudaStreamSynchronize(ctx->stream));
        return MCB_OK;
    });
}

int mcb_search_cell_batch(mcb_ctx* ctx, const double* pos3, int64_t n, int32_t* cell)
{
    if (!ctx || n < 0) return MCB_ERR_ARG;
    return with_buffers(ctx, [&]() -> int {
        DevBuf<double> dP; DevBuf<int32_t> dO;
        H2D(dP, pos3, (size_t)n * 3);
        CK(dO.alloc((size_t)n));
        mcbk::search_cell(ctx->stream, ctx->P, dP.p, n, dO.p);
        D2H(cell, dO, (size_t)n);
        CK(cudaStreamSynchronize(ctx->stream));
        return MCB_OK;
    });
}

int mcb_scatter_batch(mcb_ctx* ctx, int32_t nuclide, const uint64_t* nps, int64_t n, double* io5)
{
    if (!ctx || n < 0) return MCB_ERR_ARG;
    if (nuclide < 0 || nuclide >= ctx->n_nuclides) return ctx->fail(MCB_ERR_ARG, "nuclide %d out of range", nuclide);
    return with_buffers(ctx, [&]() -> int {
        DevBuf<uint64_t> dN; DevBuf<double> dIO;
        H2D(dN, nps, (size_t)n); H2D(dIO, io5, (size_t)n * 5);
        mcbk::scatter(ctx->stream, ctx->P, nuclide, dN.p, n, dIO.p);
        D2H(io5, dIO, (size_t)n * 5);
        CK(cudaStreamSynchronize(ctx->stream));
        return MCB_OK;
    });
}

int mcb_watt_batch(mcb_ctx* ctx, int32_t nuclide, const uint64_t* nps, const double* E, int64_t n, double* Eout)
{
    if (!ctx || n < 0) return MCB_ERR_ARG;
    if (nuclide < 0 || nuclide >= ctx->n_nuclides) return ctx->fail(MCB_ERR_ARG, "nuclide %d out of range", nuclide);
    return with_buffers(ctx, [&]() -> int {
        DevBuf<uint64_t> dN; DevBuf<double> dE, dO;
        H2D(dN, nps, (size_t)n); H2D(dE, E, (size_t)n);
        CK(dO.alloc((size_t)n));
        mcbk::watt(ctx->stream, ctx->P, nuclide, dN.p, dE.p, n, dO.p);
        D2H(Eout, dO, (size_t)n);
        CK(cudaStreamSynchronize(ctx->stream));
        return MCB_OK;
    });
}

int mcb_walk_launch_info(mcb_ctx* ctx, int32_t scoring, int32_t out4[4])
{
    if (!ctx || !out4) return MCB_ERR_ARG;
    if (!ctx->walk_mode) return ctx->fail(MCB_ERR_ARG, "event-queue mode has no walk kernel");
    CK(cudaSetDevice(ctx->device));
    int o[4];
    const int rc = mcbk::walk_launch_info(ctx->plan, scoring != 0, o);
    if (rc != 0) return ctx->fail(MCB_ERR_CUDA, "cudaFuncGetAttributes: %s", cudaGetErrorString((cudaError_t)rc));
    for (int i = 0; i < 4; i++) out4[i] = o[i];
    return MCB_OK;
}

int mcb_division_batch(mcb_ctx* ctx, const double* a, const double* b, int64_t n, double* out_shared, double* out_plain)
{
    if (!ctx || n < 0) return MCB_ERR_ARG;
    return with_buffers(ctx, [&]() -> int {
        DevBuf<double> dA, dB, dS, dP;
        H2D(dA, a, (size_t)n); H2D(dB, b, (size_t)n);
        CK(dS.alloc((size_t)n)); CK(dP.alloc((size_t)n));
        mcbk::division(ctx->stream, dA.p, dB.p, n, dS.p, dP.p);
        D2H(out_shared, dS, (size_t)n); D2H(out_plain, dP, (size_t)n);
        CK(cudaStreamSynchronize(ctx->stream));
        return MCB_OK;
    });
}

// per-history k scores of the last cycle on this rank (k_C then k_TL, shard-local history order): parity tests
int64_t mcb_get_history_k(mcb_ctx* ctx, double* kC, double* kTL, int64_t max_n)
{
    if (!ctx) return MCB_ERR_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return MCB_ERR_CUDA;
    const int64_t n = std::min<int64_t>((int64_t)ctx->shard_count, max_n);
    if (n <= 0) return 0;
    if (cudaMemcpy(kC, ctx->H.kC, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess ||
        cudaMemcpy(kTL, ctx->H.kTL, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess)
        return ctx->fail(MCB_ERR_CUDA, "history k read-back failed");
    return n;
}

}  // extern "C"
