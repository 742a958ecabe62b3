// deck — host-side problem setup: input.xml + xs_library text files -> flattened mcb_problem.
//
// Mirrors the grammar and semantics of the reference's Simulator constructor
// (src/simulator/setup.cpp:30-1069): same element and attribute
// names, same defaults, same ID assignment (insertion index), same derived data
// (sigma_t, beta, Watt g, filter grids, entropy mesh by repeated addition) and the
// same error messages.  Instead of an object graph of shared_ptrs it fills plain
// vectors that `view()` exposes as the POD `mcb_problem` of include/mcb200.h.
#ifndef MCB_DECK_H
#define MCB_DECK_H

#include <string>
#include <vector>

#include "mcb200.h"

namespace mcb {

enum DeckFlags {
    DECK_IGNORE_TRMM = 1,  // accept a <trmm> block but do not create its tally set (transport + k only)
};

struct Deck {
    // names, for reporting (Estimator::report, Estimator.cpp:368-422)
    std::string simulation_name;
    std::string mode = "fixed source";
    std::vector<std::string> nuclide_names, material_names, surface_names, cell_names;
    bool trmm_present = false;   // the deck has a <trmm> block
    bool trmm_built = false;     // ... and its nine estimators were created (not DECK_IGNORE_TRMM)

    // flattened storage
    std::vector<mcb_nuclide> nuclides;
    std::vector<double> xs_rows;
    std::vector<double> delayed_data;
    std::vector<int32_t> mat_begin{0}, mat_nuclide;
    std::vector<double> mat_density;
    std::vector<mcb_surface> surfaces;
    std::vector<mcb_cell> cells;
    std::vector<int32_t> cell_surface, cell_sense;
    std::vector<mcb_source> sources;
    std::vector<mcb_estimator> estimators;
    std::vector<mcb_score> scores;
    std::vector<mcb_filter> filters;
    std::vector<double> filter_grid;
    std::vector<double> entropy_grid;
    std::vector<double> tdmc_time, tdmc_interval;

    mcb_problem p{};  // scalar fields are filled by load(); pointers by view()

    // io_dir must end with '/' like the reference's (Main.cpp:16); xs_dir is the directory that
    // holds <ZAID>.txt (the reference hard-codes "./xs_library/", setup.cpp:326).
    // Returns false with `error` set to the reference's message where it has one.
    bool load(const std::string& io_dir, const std::string& xs_dir, int flags, std::string& error);
    bool load_string(const std::string& xml_text, const std::string& xs_dir, int flags, std::string& error);
    const mcb_problem* view();

    // search_cell (general.cpp:26-34) on the host, -1 when lost
    int search_cell(double x, double y, double z) const;
};

}  // namespace mcb
#endif
