// C entry points of libmcbhost.so (include/mcb200_host.h).
#include <algorithm>
#include <string>

#include "deck.h"
#include "mcb200_host.h"
#include "mcb_tables.h"
#include "h5lite.h"
#include "trmm_eigen.h"

struct mcbh_deck { mcb::Deck deck; };

static thread_local std::string g_error;

extern "C" {

mcbh_deck* mcbh_load_deck(const char* io_dir, const char* xs_dir, int flags)
{
    std::string dir = io_dir ? io_dir : "";
    if (dir.empty() || dir.back() != '/') dir += "/";
    mcbh_deck* d = new mcbh_deck;
    if (!d->deck.load(dir, xs_dir ? xs_dir : "./xs_library", flags, g_error)) { delete d; return nullptr; }
    d->deck.view();
    return d;
}
mcbh_deck* mcbh_load_deck_string(const char* xml_text, const char* xs_dir, int flags)
{
    mcbh_deck* d = new mcbh_deck;
    if (!d->deck.load_string(xml_text ? xml_text : "", xs_dir ? xs_dir : "./xs_library", flags, g_error)) { delete d; return nullptr; }
    d->deck.view();
    return d;
}
void mcbh_free_deck(mcbh_deck* d) { delete d; }
const char* mcbh_last_error(void) { return g_error.c_str(); }
const mcb_problem* mcbh_problem(mcbh_deck* d) { return d ? d->deck.view() : nullptr; }
void mcbh_set_run(mcbh_deck* d, uint64_t n_sample, uint64_t n_cycle, uint64_t n_passive, uint64_t seed)
{
    if (!d) return;
    if (n_sample) d->deck.p.n_sample = n_sample;
    if (n_cycle) { d->deck.p.n_cycle = n_cycle; d->deck.p.n_passive = n_passive; }
    if (seed) d->deck.p.seed = seed;
}
void mcbh_info(mcbh_deck* d, int64_t out[16])
{
    const mcb_problem* p = d->deck.view();
    const int64_t v[16] = {(int64_t)p->n_sample, (int64_t)p->n_cycle, (int64_t)p->n_passive, p->ksearch, p->n_nuclides,
                           p->n_materials, p->n_surfaces, p->n_cells, p->n_estimators, p->n_tallies, p->n_sources,
                           p->entropy_on, p->n_xs_rows, p->n_scores, p->n_filters, d->deck.trmm_present ? 1 : 0};
    for (int i = 0; i < 16; i++) out[i] = v[i];
}
const char* mcbh_name(const mcbh_deck* d, int kind, int index)
{
    if (!d || index < 0) return nullptr;
    const std::vector<std::string>* v = nullptr;
    switch (kind) {
    case 0: v = &d->deck.nuclide_names; break;
    case 1: v = &d->deck.material_names; break;
    case 2: v = &d->deck.surface_names; break;
    case 3: v = &d->deck.cell_names; break;
    case 4: return index < (int)d->deck.estimators.size() ? d->deck.estimators[index].name : nullptr;
    case 5: return index < (int)d->deck.scores.size() ? d->deck.scores[index].name : nullptr;
    default: return nullptr;
    }
    return index < (int)v->size() ? (*v)[index].c_str() : nullptr;
}
int mcbh_estimator_info(const mcbh_deck* d, int e, int64_t out[8])
{
    if (!d || e < 0 || e >= (int)d->deck.estimators.size()) return -1;
    const mcb_estimator& E = d->deck.estimators[e];
    const int64_t v[8] = {E.attach, E.score_begin, E.n_scores, E.filter_begin, E.n_filters, E.tally_begin, E.n_tallies, E.simulate};
    for (int i = 0; i < 8; i++) out[i] = v[i];
    return 0;
}
int mcbh_filter_info(const mcbh_deck* d, int f, int64_t out[4])
{
    if (!d || f < 0 || f >= (int)d->deck.filters.size()) return -1;
    const mcb_filter& F = d->deck.filters[f];
    out[0] = F.type; out[1] = F.grid_begin; out[2] = F.grid_n; out[3] = F.size;
    return 0;
}
const double* mcbh_filter_grid(const mcbh_deck* d) { return d ? d->deck.filter_grid.data() : nullptr; }
const char* mcbh_mode(const mcbh_deck* d) { return d ? d->deck.mode.c_str() : nullptr; }
const char* mcbh_simulation_name(const mcbh_deck* d) { return d ? d->deck.simulation_name.c_str() : nullptr; }
int mcbh_search_cell(const mcbh_deck* d, double x, double y, double z) { return d ? d->deck.search_cell(x, y, z) : -1; }

// TRM assembly (report.cpp:53-157) from the means of the TRMM tally set (the deck's last nine estimators):
// TRM (G+6)^2 row-major, inverse_speed G, C_initial 6, psi_initial G.  Returns G, or -1 without a TRMM set.
int mcbh_trm_assemble(mcbh_deck* d, const double* tally_mean, double* TRM, double* inverse_speed, double* C_initial,
                      double* psi_initial)
{
    if (!d || !d->deck.trmm_built || d->deck.estimators.size() < 9) return -1;
    const size_t e0 = d->deck.estimators.size() - 9;
    const mcb_estimator& Es = d->deck.estimators[e0];
    const double* simple = tally_mean + Es.tally_begin;
    const double* scatter = tally_mean + d->deck.estimators[e0 + 1].tally_begin;
    const double* prompt = tally_mean + d->deck.estimators[e0 + 2].tally_begin;
    const double* delayed[6];
    for (int j = 0; j < 6; j++) delayed[j] = tally_mean + d->deck.estimators[e0 + 3 + j].tally_begin;
    const int score_N = Es.n_scores;
    const int G = Es.n_tallies / score_N, J = 6, N = G + J;
    for (int i = 0; i < N * N; i++) TRM[i] = 0.0;
    int idx = 0;
    for (int f = 0; f < G; f++) {  // M (report.cpp:62-78)
        for (int i = 0; i < G; i++) {
            if (i == f) {
                TRM[idx] = -simple[i];
                TRM[idx] += scatter[i + i * G];
                TRM[idx] += prompt[i + i * G];
                TRM[idx] /= simple[i + G];
            } else {
                TRM[idx] = scatter[f + i * G];
                TRM[idx] += prompt[f + i * G];
                TRM[idx] /= simple[i + G];
            }
            idx++;
        }
        idx += J;
    }
    for (int j = 0; j < J; j++) {  // D (:80-88)
        for (int g = 0; g < G; g++) {
            TRM[idx] = simple[2 * G + j * G + g];
            TRM[idx] /= simple[G + g];
            idx++;
        }
        idx += J;
    }
    double lambda[6];
    for (int j = 0; j < J; j++) {  // L (:90-102)
        double num = 0.0, denom = 0.0;
        for (int g = 0; g < G; g++) { num += simple[2 * G + j * G + g]; denom += simple[2 * G + J * G + j * G + g]; }
        lambda[j] = num / denom;
        TRM[N * (G + j) + G + j] = -lambda[j];
    }
    for (int g = 0; g < G; g++) {  // P (:104-116)
        for (int j = 0; j < J; j++) {
            double num = 0.0, denom = 0.0;
            for (int gp = 0; gp < G; gp++) { num += delayed[j][g + gp * G]; denom += simple[2 * G + j * G + gp]; }
            TRM[N * g + G + j] = num / denom * lambda[j];
        }
    }
    for (int i = 0; i < G; i++) inverse_speed[i] = simple[i + (score_N - 1) * G] / simple[i + G];  // :125-135
    for (int j = 0; j < J; j++) {  // :137-147
        C_initial[j] = 0.0;
        for (int g = 0; g < G; g++) C_initial[j] += simple[2 * G + J * G + j * G + g];
    }
    for (int g = 0; g < G; g++) psi_initial[g] = simple[g + G];  // :149-157
    return G;
}

// Simulator::report (report.cpp:9-52) + Estimator::report (Estimator.cpp:368-422) + EstimatorK::report (:562-594)
int mcbh_write_output(mcbh_deck* d, const char* path, uint64_t n_track, const double* k_cycle, const double* H_cycle,
                      int32_t n_cycle, const double* k_avg, const double* k_uncer, int32_t n_active,
                      const double* tally_mean, const double* tally_uncer, int64_t n_tallies)
{
    if (!d || !path) return -1;
    const mcb_problem* p = d->deck.view();
    if (n_tallies != p->n_tallies) { g_error = "mcbh_write_output: tally count does not match the deck"; return -1; }
    h5lite::File f;
    h5lite::Group& summary = f.root.group("summary");
    summary.dataset_u64("Ncycle", p->n_cycle);
    summary.dataset_u64("Nsample", p->n_sample);
    summary.dataset_u64("Npassive", p->n_passive);
    summary.dataset_u64("Ntrack", n_track);
    summary.dataset_string("mode", d->deck.mode);
    h5lite::Group& roulette = summary.group("survival_roulette");
    roulette.dataset_f64("wr", p->wr);
    roulette.dataset_f64("ws", p->ws);
    if (p->tdmc_on) summary.group("tdmc").dataset_f64("time", {(uint64_t)p->n_tdmc}, p->tdmc_time);  // report.cpp:40-46
    static const char* const f_name[] = {"surface", "cell", "energy", "energy_initial", "time", "time"};  // Estimator.h:307-369
    static const char* const f_unit[] = {"id#", "id#", "eV", "eV", "s", "s"};
    for (size_t e = 0; e < d->deck.estimators.size(); e++) {
        const mcb_estimator& E = d->deck.estimators[e];
        h5lite::Group& g = f.root.group(E.name);
        std::string indexing;
        std::vector<uint64_t> dims;
        uint64_t per_score = 1;
        for (int i = 0; i < E.n_filters; i++) {
            const mcb_filter& F = d->deck.filters[E.filter_begin + i];
            indexing += std::string("[") + f_name[F.type] + "]";
            h5lite::Dataset& ds = g.dataset_f64(f_name[F.type], {(uint64_t)F.grid_n}, d->deck.filter_grid.data() + F.grid_begin);
            ds.attr_string("unit", f_unit[F.type]);
            dims.push_back((uint64_t)F.size);
            per_score *= (uint64_t)F.size;
        }
        g.attr_string("indexing", indexing);
        for (int k = 0; k < E.n_scores; k++) {
            h5lite::Group& sg = g.group(d->deck.scores[E.score_begin + k].name);
            const int64_t t0 = E.tally_begin + (int64_t)k * (int64_t)per_score;
            sg.dataset_f64("mean", dims, tally_mean + t0);
            sg.dataset_f64("uncertainty", dims, tally_uncer + t0);
        }
    }
    if (p->ksearch) {
        h5lite::Group& ks = f.root.group("ksearch");
        ks.dataset_f64("k_cycle", {(uint64_t)n_cycle}, k_cycle);
        ks.dataset_f64("H_cycle", {(uint64_t)n_cycle}, H_cycle);
        ks.dataset_f64("mean", n_active > 0 ? k_avg[n_active - 1] : 0.0);
        ks.dataset_f64("uncertainty", n_active > 0 ? k_uncer[n_active - 1] : 0.0);
        h5lite::Group& ka = ks.group("k_active");
        ka.dataset_f64("mean", {(uint64_t)n_active}, k_avg);
        ka.dataset_f64("uncertainty", {(uint64_t)n_active}, k_uncer);
    }
    if (d->deck.trmm_built) {  // report.cpp:53-157
        const mcb_estimator& Es = d->deck.estimators[d->deck.estimators.size() - 9];
        const int G = Es.n_tallies / Es.n_scores, N = G + 6;
        std::vector<double> TRM((size_t)N * N), inv(G), Ci(6), psi(G);
        mcbh_trm_assemble(d, tally_mean, TRM.data(), inv.data(), Ci.data(), psi.data());
        f.root.dataset_f64("TRM", {(uint64_t)N, (uint64_t)N}, TRM.data());
        f.root.dataset_f64("inverse_speed", {(uint64_t)G}, inv.data());
        f.root.dataset_f64("C_initial", {6}, Ci.data());
        f.root.dataset_f64("psi_initial", {(uint64_t)G}, psi.data());
    }
    return f.write(path, g_error) ? 0 : -1;
}

int mcbh_tdmc(const mcbh_deck* d, double* time_out, double* interval_out)
{
    if (!d) return 0;
    const size_t n = d->deck.tdmc_time.size();
    for (size_t i = 0; i < n; i++) {
        if (time_out) time_out[i] = d->deck.tdmc_time[i];
        if (interval_out) interval_out[i] = d->deck.tdmc_interval[i];
    }
    return (int)n;
}

int mcbh_trmm_postprocess(const char* output_h5)
{
    return mcbhost::trmm_postprocess(output_h5 ? output_h5 : "", g_error) ? 0 : -1;
}

int mcbh_eigen_general(int32_t n, const double* A, double* w_pairs, double* v_pairs)
{
    std::vector<std::complex<double>> w, V;
    if (!A || !w_pairs || !v_pairs) { g_error = "mcbh_eigen_general: null argument"; return -1; }
    if (!mcbhost::eigen_general(n, A, w, V, g_error)) return -1;
    for (size_t i = 0; i < w.size(); i++) { w_pairs[2 * i] = w[i].real(); w_pairs[2 * i + 1] = w[i].imag(); }
    for (size_t i = 0; i < V.size(); i++) { v_pairs[2 * i] = V[i].real(); v_pairs[2 * i + 1] = V[i].imag(); }
    return 0;
}

int mcbh_union_indices(mcbh_deck* d, int material, const double* E, int64_t n, int32_t* idx_out, int64_t stats[4])
{
    if (!d) return -1;
    const mcb_problem* p = d->deck.view();
    if (material < 0 || material >= p->n_materials) return -1;
    mcb::MaterialTables T;
    mcb::build_material_tables(p, material, MCB_HASH_BITS_DEFAULT, T);
    for (int64_t i = 0; i < n; i++) {
        const int u = mcb_union_count_less(T.U.data(), T.hash.data(), T.key_min, T.n_hash, T.shift, (int32_t)T.U.size(), E[i]) - 1;
        // the device reads the indices through the bin records (mcb_union_lookup): both routes must agree
        const int32_t* rec = nullptr;
        bool from_rec = false;
        const int lo = mcb_union_lookup(T.U.data(), T.hrec.data(), T.hrec_stride, T.key_min, T.n_hash, T.shift, E[i], &rec, &from_rec);
        if (lo - 1 != u) return -2;
        for (int k = 0; k < T.n_nuc; k++) {
            int idx = u < 0 ? -1 : T.map[(size_t)u * T.n_nuc + k];
            if ((from_rec ? rec[2 + k] : T.map[(size_t)u * T.n_nuc + k]) != idx) return -2;
            if (idx == MCB_MAP_BISECT) {
                const mcb_nuclide& N = p->nuclides[p->mat_nuclide[p->mat_begin[material] + k]];
                idx = mcb_row_bisect(p->xs_rows + (size_t)N.row_begin * MCB_XS_ROW, N.n_rows, E[i]);
            }
            idx_out[i * T.n_nuc + k] = idx;
        }
    }
    if (stats) { stats[0] = (int64_t)T.U.size(); stats[1] = T.n_hash; stats[2] = T.shift; stats[3] = T.max_bin; }
    return T.n_nuc;
}

int mcbh_cross_neighbors(mcbh_deck* d, int32_t* out, int32_t max_n)
{
    if (!d || !out) return -1;
    std::vector<int32_t> nb;
    mcb::build_cross_neighbors(d->deck.view(), nb);
    if ((int64_t)nb.size() > max_n) return -1;
    std::copy(nb.begin(), nb.end(), out);
    return (int)nb.size();
}

}  // extern "C"
