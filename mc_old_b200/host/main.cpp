// MCB.exe <dir> — the reference's command line (Main.cpp:8-31) on top of the C ABI: reads <dir>/input.xml and
// ./xs_library/<ZAID>.txt (relative to the CWD, setup.cpp:326; MCB_XS_LIBRARY overrides), runs the transport loop on
// the GPU (mcb_run_cycle replaces Simulator::start(), handler.cpp:11-48), prints the reference's banners and
// per-cycle lines (Estimator.cpp:536-554) and writes <dir>/output.h5 (report.cpp:9-52).
//
// Several GPUs: start one process per GPU with RANK / WORLD_SIZE / LOCAL_RANK in the environment (what torchrun, srun
// or a shell loop provide); the histories of every generation are sharded over the ranks, rank 0 prints and writes
// the output.  The 128-byte NCCL id travels through a file (MCB_ID_FILE, default
// $XDG_RUNTIME_DIR|$TMPDIR|/tmp/mcb_nccl_id.<run id>.<MASTER_PORT>.<uid>, tagged with the job so that a stale one is refused).
#include <chrono>
#include <thread>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "mcb200.h"
#include <sys/stat.h>
#include <ctime>

#include "mcb200_host.h"

// MCB_TIMING=1: wall seconds of the program's phases on stderr (deck, context, first cycle, the other cycles, output)
static double wall_s()
{
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

int main(int argc, char* argv[])
{
    const bool timing = getenv("MCB_TIMING") != nullptr;
    double t_mark[6] = {wall_s(), 0, 0, 0, 0, 0};
    if (argc == 1) {
        std::cout << "[ERROR] Please provide input.xml directory...\n";
        std::exit(EXIT_FAILURE);
    }
    const std::string io_dir = std::string(argv[1]) + "/";
    const char* xs = getenv("MCB_XS_LIBRARY") ? getenv("MCB_XS_LIBRARY") : "./xs_library";
    int flags = 0;
    for (int i = 2; i < argc; i++) if (!strcmp(argv[i], "--ignore-trmm")) flags |= MCBH_IGNORE_TRMM;
    mcb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    const int world = getenv("WORLD_SIZE") ? atoi(getenv("WORLD_SIZE")) : 1;
    const int rank = getenv("RANK") ? atoi(getenv("RANK")) : 0;
    const bool root = rank == 0;
    cfg.device = getenv("MCB_DEVICE") ? atoi(getenv("MCB_DEVICE")) : (getenv("LOCAL_RANK") ? atoi(getenv("LOCAL_RANK")) : 0);
    if (world <= 1 && !getenv("CUDA_VISIBLE_DEVICES")) {
        // One process, one GPU: the driver initialises every GPU it can see (measured on an 8-GPU box: 0.6-1.6 s of the
        // ~1 s a shipped-size deck takes from start to output.h5), so only the one in use is left visible
        setenv("CUDA_VISIBLE_DEVICES", std::to_string(cfg.device).c_str(), 1);
        cfg.device = 0;
    }
    cfg.rank = rank; cfg.world = world > 1 ? world : 1;
    // the CUDA driver and context start on a second thread while the deck and the xs_library are read (0.2 s of the ~1 s
    // they take); a failure shows up in mcb_create with its message
    std::thread warm([dev = cfg.device] { mcb_warm_up(dev); });
    mcbh_deck* deck = mcbh_load_deck(io_dir.c_str(), xs, flags);
    if (!deck) { warm.join(); std::cout << mcbh_last_error() << "\n"; std::exit(EXIT_FAILURE); }
    const mcb_problem* p = mcbh_problem(deck);
    t_mark[1] = wall_s();
    warm.join();
    mcb_ctx* ctx = nullptr;
    if (mcb_create(p, &cfg, &ctx) != MCB_OK) { std::cout << mcb_last_error(nullptr) << "\n"; std::exit(EXIT_FAILURE); }
    if (world > 1) {
        // The NCCL unique id goes from rank 0 to the others through a file.  Its name carries the launcher's job
        // identity (torchrun's TORCHELASTIC_RUN_ID / MASTER_PORT, or MCB_ID_FILE) and lives in a directory of the
        // user's; the payload is tagged with the same job token, so an id left behind by a crashed run under the same
        // name is not mistaken for this run's (rank 0 removes any such file before it writes, the others accept only
        // a file written after they started waiting).
        const char* tmpdir = getenv("XDG_RUNTIME_DIR") ? getenv("XDG_RUNTIME_DIR") : (getenv("TMPDIR") ? getenv("TMPDIR") : "/tmp");
        const std::string job = std::string(getenv("TORCHELASTIC_RUN_ID") ? getenv("TORCHELASTIC_RUN_ID") : "none") + "." +
                                (getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0") + "." + std::to_string((long)getuid());
        const std::string id_file = getenv("MCB_ID_FILE") ? getenv("MCB_ID_FILE") : std::string(tmpdir) + "/mcb_nccl_id." + job;
        const time_t t_start = time(nullptr);
        char id[128], tag[64];
        memset(tag, 0, sizeof(tag));
        snprintf(tag, sizeof(tag), "%s", job.c_str());
        if (root) {
            std::remove(id_file.c_str());  // stale id of an earlier run
            if (mcb_comm_unique_id(id) != MCB_OK) { std::cout << mcb_last_error(nullptr) << "\n"; std::exit(EXIT_FAILURE); }
            std::ofstream f(id_file + ".tmp", std::ios::binary);
            f.write(tag, sizeof(tag));
            f.write(id, 128);
            f.close();
            std::rename((id_file + ".tmp").c_str(), id_file.c_str());
        } else {
            for (int tries = 0;; tries++) {
                struct stat sb;
                char got[64];
                std::ifstream f(id_file, std::ios::binary);
                // only a file with this job's tag that is not older than this process (minus clock slack)
                if (f && stat(id_file.c_str(), &sb) == 0 && sb.st_mtime + 120 >= t_start && f.read(got, sizeof(got)) &&
                    !memcmp(got, tag, sizeof(tag)) && f.read(id, 128)) break;
                if (tries > 6000) { std::cout << "[ERROR] no NCCL id in " << id_file << "\n"; std::exit(EXIT_FAILURE); }
                usleep(10000);
            }
        }
        if (mcb_comm_init(ctx, id) != MCB_OK) { std::cout << mcb_last_error(ctx) << "\n"; std::exit(EXIT_FAILURE); }
        if (root) std::remove(id_file.c_str());
    }
    t_mark[2] = wall_s();
    if (root) std::cout << "\nSimulation setup done,\nNow running the simulation...\n\n";

    std::vector<double> k_cycle, H_cycle, k_avg, k_uncer;
    uint64_t n_track = 0;
    for (uint64_t icycle = 0; icycle < p->n_cycle; icycle++) {  // handler.cpp:14
        mcb_cycle_result r;
        if (mcb_run_cycle(ctx, &r) != MCB_OK) { std::cout << mcb_last_error(ctx) << "\n"; std::exit(EXIT_FAILURE); }
        if (icycle == 0) t_mark[3] = wall_s();
        n_track += r.n_tracks;  // general.cpp:76
        if (p->ksearch) {       // EstimatorK::report_cycle (Estimator.cpp:526-561)
            k_cycle.push_back(r.k_cycle); H_cycle.push_back(r.H);
            if (root) std::cout << icycle + 1 << "   " << r.k_cycle;
            if (icycle >= p->n_passive) {
                k_avg.push_back(r.k_avg); k_uncer.push_back(r.k_uncer);
                if (root) std::cout << "   " << r.k_avg << "   +/-   " << r.k_uncer;
            }
            if (root) std::cout << "   (" << r.H << ")" << std::endl;
        }
    }
    t_mark[4] = wall_s();
    if (!root) {  // every rank holds the same global results; rank 0 reports them
        mcb_destroy(ctx);
        mcbh_free_deck(deck);
        return 0;
    }
    std::cout << "Simulation done!\n\nReporting simulation output...\n";
    std::vector<double> mean((size_t)p->n_tallies), uncer((size_t)p->n_tallies);
    mcb_get_tallies(ctx, mean.data(), uncer.data(), p->n_tallies);
    const std::string out = io_dir + "output.h5";
    if (mcbh_write_output(deck, out.c_str(), n_track, k_cycle.data(), H_cycle.data(), (int32_t)k_cycle.size(), k_avg.data(),
                          k_uncer.data(), (int32_t)k_avg.size(), mean.data(), uncer.data(), p->n_tallies) != 0) {
        std::cout << mcbh_last_error() << "\n";
        std::exit(EXIT_FAILURE);
    }
    std::cout << "Simulation output done!\n";
    t_mark[5] = wall_s();
    if (timing)
        fprintf(stderr, "[mcb timing] deck %.3f s  context %.3f s  first cycle %.3f s  other %llu cycles %.3f s  output %.3f s\n",
                t_mark[1] - t_mark[0], t_mark[2] - t_mark[1], t_mark[3] - t_mark[2], (unsigned long long)(p->n_cycle - 1),
                t_mark[4] - t_mark[3], t_mark[5] - t_mark[4]);
    mcb_destroy(ctx);
    mcbh_free_deck(deck);
    if (timing) {
        const double now = std::chrono::duration<double>(std::chrono::system_clock::now().time_since_epoch()).count();
        fprintf(stderr, "[mcb timing] epoch main %.3f  epoch end %.3f  (teardown %.3f s)\n", now - wall_s(), now, wall_s() - t_mark[5]);
    }
    return 0;
}
