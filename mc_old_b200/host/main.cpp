// MCB.exe <dir> — the reference's command line (Main.cpp:8-31) on top of the C ABI: reads <dir>/input.xml and
// ./xs_library/<ZAID>.txt (relative to the CWD, setup.cpp:326; MCB_XS_LIBRARY overrides), runs the transport loop on
// the GPU (mcb_run_cycle replaces Simulator::start(), handler.cpp:11-48), prints the reference's banners and
// per-cycle lines (Estimator.cpp:536-554) and writes <dir>/output.h5 (report.cpp:9-52).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "mcb200.h"
#include "mcb200_host.h"

int main(int argc, char* argv[])
{
    if (argc == 1) {
        std::cout << "[ERROR] Please provide input.xml directory...\n";
        std::exit(EXIT_FAILURE);
    }
    const std::string io_dir = std::string(argv[1]) + "/";
    const char* xs = getenv("MCB_XS_LIBRARY") ? getenv("MCB_XS_LIBRARY") : "./xs_library";
    int flags = 0;
    for (int i = 2; i < argc; i++) if (!strcmp(argv[i], "--ignore-trmm")) flags |= MCBH_IGNORE_TRMM;
    mcbh_deck* deck = mcbh_load_deck(io_dir.c_str(), xs, flags);
    if (!deck) { std::cout << mcbh_last_error() << "\n"; std::exit(EXIT_FAILURE); }
    const mcb_problem* p = mcbh_problem(deck);
    mcb_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = getenv("MCB_DEVICE") ? atoi(getenv("MCB_DEVICE")) : 0;
    cfg.rank = 0; cfg.world = 1;
    mcb_ctx* ctx = nullptr;
    if (mcb_create(p, &cfg, &ctx) != MCB_OK) { std::cout << mcb_last_error(nullptr) << "\n"; std::exit(EXIT_FAILURE); }
    std::cout << "\nSimulation setup done,\nNow running the simulation...\n\n";

    std::vector<double> k_cycle, H_cycle, k_avg, k_uncer;
    uint64_t n_track = 0;
    for (uint64_t icycle = 0; icycle < p->n_cycle; icycle++) {  // handler.cpp:14
        mcb_cycle_result r;
        if (mcb_run_cycle(ctx, &r) != MCB_OK) { std::cout << mcb_last_error(ctx) << "\n"; std::exit(EXIT_FAILURE); }
        n_track += r.n_tracks;  // general.cpp:76
        if (p->ksearch) {       // EstimatorK::report_cycle (Estimator.cpp:526-561)
            k_cycle.push_back(r.k_cycle); H_cycle.push_back(r.H);
            std::cout << icycle + 1 << "   " << r.k_cycle;
            if (icycle >= p->n_passive) {
                k_avg.push_back(r.k_avg); k_uncer.push_back(r.k_uncer);
                std::cout << "   " << r.k_avg << "   +/-   " << r.k_uncer;
            }
            std::cout << "   (" << r.H << ")\n";
        }
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
    mcb_destroy(ctx);
    mcbh_free_deck(deck);
    return 0;
}
