#include "deck.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>

#include "../csrc/mcb_physics.h"
#include "xml_lite.h"

namespace mcb {

namespace {

int find_name(const std::vector<std::string>& v, const std::string& n)
{
    for (size_t i = 0; i < v.size(); i++) { if (v[i] == n) return (int)i; }
    return -1;
}

void copy_name(char* dst, size_t cap, const std::string& s)
{
    std::memset(dst, 0, cap);
    std::strncpy(dst, s.c_str(), cap - 1);
}

// DistributionWatt constructor (Distribution.cpp:14-23)
void watt_g(const double* a, const double* b, double* g)
{
    for (int i = 0; i < 3; i++) {
        const double C = (1.0 + a[i] * b[i] / 8.0);
        g[i] = std::sqrt(C * C - 1.0) + C;
    }
}

// named scalar ("double") distributions (setup.cpp:193-241)
struct NamedDist1 { std::string name; mcb_dist1 d; };
// named point distributions (setup.cpp:244-289)
struct NamedPoint {
    std::string name;
    int kind;  // MCB_DIR_*
    double xyz[3];
    mcb_dist1 comp[3];
};

// filter grid attributes shared by <estimator><filter> and <trmm><filter> (setup.cpp:744-777)
bool parse_filter_grid(const XmlNode& f, std::vector<double>& g, const std::string& who, std::string& error)
{
    g.clear();
    if (f.attribute("grid")) {
        std::istringstream iss(f.attribute("grid").value());
        for (double s; iss >> s;) g.push_back(s);
    } else if (f.attribute("grid_linear")) {
        double a = 0, b = 0, step = 0;
        std::istringstream iss(f.attribute("grid_linear").value());
        iss >> a >> step >> b;
        g.push_back(a);
        while (g.back() < b) g.push_back(g.back() + step);
        g.pop_back();
        g.push_back(b);
    } else if (f.attribute("grid_lethargy")) {
        double a = 0, b = 0, N = 0, step;
        std::istringstream iss(f.attribute("grid_lethargy").value());
        iss >> a >> b >> N;
        step = std::log(b / a) / N;
        g.push_back(0.0);
        while (g.size() != N + 1) g.push_back(g.back() + step);
        std::reverse(g.begin(), g.end());
        for (size_t i = 0; i < g.size(); i++) g[i] = b * std::exp(-g[i]);
    } else {
        error = "[ERROR] Need filter grid for " + who;
        return false;
    }
    return true;
}

// One xs_library row per text line: E sigma_s sigma_c sigma_f nu [nu_delayed]; a missing 6th column reads as 0.
// For the 6-column files this equals the reference's token-stream read (setup.cpp:366); for the 5-column files
// (001001, 002003, 005011) the token stream misaligns — fixed here and in the oracle (patch C).
bool load_zaid(const std::string& xs_dir, const std::string& zaid, Deck& D, mcb_nuclide& N, std::string& error)
{
    const std::string path = xs_dir + "/" + zaid + ".txt";
    std::ifstream f(path);
    if (!f) { error = "Failed to read A in library file " + path; return false; }
    std::string line;
    auto tokens = [](const std::string& l, double* c, int maxn) {
        std::istringstream ls(l);
        int n = 0;
        while (n < maxn && (ls >> c[n])) n++;
        return n;
    };
    double c[7];
    if (!std::getline(f, line) || tokens(line, c, 1) < 1) { error = "Failed to read A in library file " + path; return false; }
    N.A = c[0];
    for (int i = 0; i < 3; i++) {
        if (!std::getline(f, line) || tokens(line, c, 2) < 2) { error = "Faled to read ab in library file " + path; return false; }
        N.watt_a[i] = c[0];
        N.watt_b[i] = c[1];
    }
    watt_g(N.watt_a, N.watt_b, N.watt_g);
    N.row_begin = (int64_t)(D.xs_rows.size() / MCB_XS_ROW);
    N.n_rows = 0;
    while (std::getline(f, line)) {
        const int n = tokens(line, c, 6);
        if (n < 5) continue;
        if (n < 6) c[5] = 0.0;
        // row = E, sigma_s, sigma_c, sigma_f, nu, beta ; beta = nu_d != 0 ? nu_d/nu : nu_d  (setup.cpp:367-375)
        const double beta = (c[5] != 0) ? c[5] / c[4] : c[5];
        const double row[MCB_XS_ROW] = {c[0], c[1], c[2], c[3], c[4], beta};
        D.xs_rows.insert(D.xs_rows.end(), row, row + MCB_XS_ROW);
        N.n_rows++;
    }
    if (N.n_rows == 0) { error = "no cross-section rows in " + path; return false; }
    for (int i = 0; i < 6; i++) { N.lambda[i] = 1.0; N.fraction[i] = 0.0; N.chid_cdf_begin[i] = 0; N.chid_cdf_n[i] = 0; }
    N.chid_E_begin = 0; N.chid_E_n = 0;
    N.has_delayed = 0;
    // delayed data iff the first sigma_f is non-zero (setup.cpp:377)
    if (D.xs_rows[(size_t)N.row_begin * MCB_XS_ROW + 3] != 0) {
        const std::string dpath = xs_dir + "/" + zaid + "D.txt";
        std::ifstream d(dpath);
        if (!d) { error = "cannot open delayed-neutron file " + dpath; return false; }
        for (int i = 0; i < 6; i++) d >> N.lambda[i];
        for (int i = 0; i < 6; i++) d >> N.fraction[i];
        std::vector<double> dE;
        std::vector<std::vector<double>> cdf(6, std::vector<double>(1, 0.0));
        while (d >> c[0] >> c[1] >> c[2] >> c[3] >> c[4] >> c[5] >> c[6]) {
            dE.push_back(c[0]);
            for (int i = 0; i < 6; i++) cdf[i].push_back(c[i + 1]);
        }
        // left-Riemann accumulation with truncation at the first zero and the reference's index shift
        // (setup.cpp:398-409).  The reference reads d_E[d_E.size()] when no zero is met; we stop there instead.
        for (int j = 0; j < 6; j++) {
            for (size_t i = 1; i < dE.size() + 1; i++) {
                if (cdf[j][i] == 0 || i >= dE.size()) { cdf[j].resize(i); break; }
                cdf[j][i] = cdf[j][i - 1] + cdf[j][i] * (dE[i] - dE[i - 1]);
            }
        }
        N.has_delayed = 1;
        N.chid_E_begin = (int32_t)D.delayed_data.size();
        N.chid_E_n = (int32_t)dE.size();
        D.delayed_data.insert(D.delayed_data.end(), dE.begin(), dE.end());
        for (int j = 0; j < 6; j++) {
            N.chid_cdf_begin[j] = (int32_t)D.delayed_data.size();
            N.chid_cdf_n[j] = (int32_t)cdf[j].size();
            D.delayed_data.insert(D.delayed_data.end(), cdf[j].begin(), cdf[j].end());
        }
    }
    return true;
}

}  // namespace

int Deck::search_cell(double x, double y, double z) const
{
    return mcb_search_cell(cells.data(), (int)cells.size(), surfaces.data(), cell_surface.data(), cell_sense.data(), x, y, z);
}

bool Deck::load(const std::string& io_dir, const std::string& xs_dir, int flags, std::string& error)
{
    std::ifstream f(io_dir + "input.xml", std::ios::binary);
    if (!f) { error = "cannot open " + io_dir + "input.xml"; return false; }
    std::ostringstream ss;
    ss << f.rdbuf();
    return load_string(ss.str(), xs_dir, flags, error);
}

bool Deck::load_string(const std::string& xml_text, const std::string& xs_dir, int flags, std::string& error)
{
    XmlNode doc;
    if (!xml_parse_string(xml_text, doc, error)) { error = "input.xml: " + error; return false; }
    *this = Deck();
    p.abi_version = MCB_ABI_VERSION;
    p.wr = 0.001;  // simulator.h:90-91
    p.ws = 1.0;
    p.seed = 1;    // Random.cpp:97,105
    p.n_cycle = 1;
    p.n_passive = 0;

    static const XmlNode empty;
    auto top = [&](const char* n) -> const XmlNode& { const XmlNode* c = doc.child(n); return c ? *c : empty; };

    // ---- basic parameters (setup.cpp:42-50) ----
    const XmlNode& sim = top("simulation");
    const XmlNode* description = sim.child("description");
    const XmlNode* ksearch = sim.child("ksearch");
    const XmlNode* entropy = sim.child("entropy");
    const XmlNode* tdmc = sim.child("tdmc");
    if (description) {
        simulation_name = description->attribute("name").value();
        p.n_sample = (uint64_t)description->attribute("samples").as_double();
    }
    if (p.n_sample == 0) { error = "[INPUT ERROR] <simulation><description samples=.../> missing or zero"; return false; }

    // ---- population control (setup.cpp:57-67) ----
    if (const XmlNode* ctrl = doc.child("population_control")) {
        if (const XmlNode* comb = ctrl->child("particle_comb")) {
            p.comb_on = 1;
            p.comb_bank_max = comb->attribute("bank_max").as_int();
            p.comb_teeth = comb->attribute("teeth").as_int();
            if (p.comb_teeth < 1 || p.comb_teeth > 256 || p.comb_bank_max < 1) {
                error = "[INPUT ERROR] <particle_comb> needs bank_max >= 1 and 1 <= teeth <= 256";
                return false;
            }
        }
    }

    // ---- ksearch + entropy (setup.cpp:73-128) ----
    if (ksearch) {
        p.ksearch = 1;
        if (entropy) {
            const char* ax[3] = {"x", "y", "z"};
            p.entropy_on = 1;
            for (int a = 0; a < 3; a++) {
                const XmlNode* e = entropy->child(ax[a]);
                if (!e) { error = std::string("[INPUT ERROR] <entropy> needs <") + ax[a] + ">"; return false; }
                const double mn = e->attribute("min").as_double();
                const double mx = e->attribute("max").as_double();
                const int step = e->attribute("step").as_int();
                if (step < 1) { error = "[INPUT ERROR] <entropy> step must be >= 1"; return false; }
                const double d = (mx - mn) / step;
                // grid by repeated addition (setup.cpp:88-92)
                double v = mn;
                entropy_grid.push_back(v);
                for (int i = 0; i < step; i++) { v = v + d; entropy_grid.push_back(v); }
                p.entropy_n[a] = step + 1;
            }
        }
        const uint64_t active = (uint64_t)ksearch->attribute("active_cycles").as_double();
        p.n_passive = (uint64_t)ksearch->attribute("passive_cycles").as_double();
        p.n_cycle = active + p.n_passive;
        mode = "k-eigenvalue";
    }
    // ---- TDMC (setup.cpp:133-169) ----
    if (tdmc) {
        p.tdmc_on = 1;
        tdmc_interval.push_back(0.0);
        if (tdmc->attribute("time")) {
            std::istringstream iss(tdmc->attribute("time").value());
            for (double s; iss >> s;) {
                tdmc_time.push_back(s);
                tdmc_interval.push_back(s - tdmc_interval.back());  // previous interval, not previous time (setup.cpp:146)
            }
            tdmc_interval.erase(tdmc_interval.begin());
        } else if (tdmc->attribute("time_linear")) {
            double a = 0.0, b = 0.0, step = 0.0;
            std::istringstream iss(tdmc->attribute("time_linear").value());
            iss >> a >> step >> b;
            step = (b - a) / step;
            if (!(step > 0.0)) { error = "[INPUT ERROR] <tdmc time_linear=\"a n b\"> needs b > a and n > 0"; return false; }
            tdmc_time.push_back(a);
            tdmc_interval.push_back(a - tdmc_interval.back());
            while (tdmc_time.back() < b) {
                tdmc_time.push_back(tdmc_time.back() + step);
                tdmc_interval.push_back(tdmc_time.back() - tdmc_interval.back());
            }
            tdmc_interval.erase(tdmc_interval.begin());
            tdmc_time.pop_back();
            tdmc_time.push_back(b);
        }
        if (ksearch) { error = "ksearch and tdmc could not coexist"; return false; }
        if (tdmc_time.empty()) { error = "[INPUT ERROR] <tdmc> needs a time grid (the reference reads past the end of an empty one)"; return false; }
        mode = "time-dependent";
    }

    // ---- user distributions, resolved iteratively (setup.cpp:174-305) ----
    std::vector<NamedDist1> dist1;
    std::vector<NamedPoint> distp;
    {
        const XmlNode& dn = top("distributions");
        const size_t total = dn.kids.size();
        size_t set = 0;
        auto find1 = [&](const std::string& n) -> const NamedDist1* { for (auto& d : dist1) if (d.name == n) return &d; return nullptr; };
        auto findp = [&](const std::string& n) -> const NamedPoint* { for (auto& d : distp) if (d.name == n) return &d; return nullptr; };
        while (set < total) {
            const size_t before = set;
            for (const XmlNode& d : dn.kids) {
                const std::string type = d.name;
                const std::string name = d.attribute("name").value();
                const std::string data = d.attribute("datatype").value();
                if (data == "double") {
                    if (find1(name)) continue;
                    NamedDist1 nd; nd.name = name; std::memset(&nd.d, 0, sizeof(nd.d));
                    if (type == "delta") {
                        nd.d.kind = MCB_DIST_DELTA; nd.d.a = d.attribute("val").as_double();
                    } else if (type == "uniform") {
                        nd.d.kind = MCB_DIST_UNIFORM; nd.d.a = d.attribute("a").as_double(); nd.d.b = d.attribute("b").as_double();
                    } else if (type == "watt") {
                        // U-235 parameters, hard-coded in the reference (setup.cpp:216-225)
                        const double a[3] = {0.988, 0.988, 1.028}, b[3] = {2.249, 2.249, 2.084};
                        nd.d.kind = MCB_DIST_WATT;
                        std::memcpy(nd.d.watt_a, a, sizeof(a)); std::memcpy(nd.d.watt_b, b, sizeof(b));
                        watt_g(nd.d.watt_a, nd.d.watt_b, nd.d.watt_g);
                    } else {
                        error = "unsupported distribution with data type " + data; return false;
                    }
                    dist1.push_back(nd);
                } else if (data == "point") {
                    if (findp(name)) continue;
                    NamedPoint np; np.name = name; std::memset(np.xyz, 0, sizeof(np.xyz)); std::memset(np.comp, 0, sizeof(np.comp));
                    if (type == "delta") {
                        np.kind = MCB_DIR_DELTA;
                        np.xyz[0] = d.attribute("x").as_double(); np.xyz[1] = d.attribute("y").as_double(); np.xyz[2] = d.attribute("z").as_double();
                    } else if (type == "isotropic") {
                        np.kind = MCB_DIR_ISOTROPIC;
                    } else if (type == "independentXYZ") {
                        const NamedDist1* dx = find1(d.attribute("x").value());
                        const NamedDist1* dy = find1(d.attribute("y").value());
                        const NamedDist1* dz = find1(d.attribute("z").value());
                        if (!dx || !dy || !dz) continue;  // not resolved yet
                        np.kind = MCB_DIR_XYZ; np.comp[0] = dx->d; np.comp[1] = dy->d; np.comp[2] = dz->d;
                    } else {
                        error = "unsupported " + data + " distribution of type " + type; return false;
                    }
                    distp.push_back(np);
                } else {
                    error = "unsupported distribution with data type " + data; return false;
                }
                set++;
            }
            if (before == set) { error = "distributions could not be resolved. "; return false; }
        }
    }

    // ---- nuclides (setup.cpp:310-467) ----
    for (const XmlNode* n : top("nuclides").children("nuclide")) {
        mcb_nuclide N;
        std::memset(&N, 0, sizeof(N));
        N.A = MCB_MAX_FLOAT;
        if (n->attribute("ZAID")) {
            if (!load_zaid(xs_dir, n->attribute("ZAID").value(), *this, N, error)) return false;
        } else {
            if (n->attribute("A")) N.A = n->attribute("A").as_double();
            double capture = 0.0;
            for (const XmlNode& r : n->kids) {
                if (!r.attribute("xs")) { error = "[ERROR-INPUT] Unknown XS type..."; return false; }
                if (r.name == "capture") capture = r.attribute("xs").as_double();
                else { error = "User defined nuclide only support capture now"; return false; }
            }
            // constant capture-only nuclide = one-row table; total = absorb = capture, scatter = fission = 0
            // (the reference leaves these reactions null and crashes; oracle patch A does the same as here)
            N.row_begin = (int64_t)(xs_rows.size() / MCB_XS_ROW);
            N.n_rows = 1;
            const double row[MCB_XS_ROW] = {0.0, 0.0, capture, 0.0, 0.0, 0.0};
            xs_rows.insert(xs_rows.end(), row, row + MCB_XS_ROW);
            for (int i = 0; i < 6; i++) { N.lambda[i] = 1.0; N.fraction[i] = 0.0; }
            watt_g(N.watt_a, N.watt_b, N.watt_g);
        }
        nuclides.push_back(N);
        nuclide_names.push_back(n->attribute("name").value());
    }

    // ---- materials (setup.cpp:474-492) ----
    for (const XmlNode* m : top("materials").children("material")) {
        for (const XmlNode* n : m->children("nuclide")) {
            const int id = find_name(nuclide_names, n->attribute("name").value());
            if (id < 0) { error = "[INPUT_ERROR] Unknown nuclide found..."; return false; }
            mat_nuclide.push_back(id);
            mat_density.push_back(n->attribute("density").as_double());
        }
        if ((int)mat_nuclide.size() - mat_begin.back() > MCB_MAX_MAT_NUCLIDES) {
            error = "material has more than MCB_MAX_MAT_NUCLIDES nuclides"; return false;
        }
        mat_begin.push_back((int32_t)mat_nuclide.size());
        material_names.push_back(m->attribute("name").value());
    }

    // ---- surfaces (setup.cpp:499-584) ----
    for (const XmlNode& s : top("surfaces").kids) {
        mcb_surface S;
        std::memset(&S, 0, sizeof(S));
        std::string bc = "transmission";
        if (s.attribute("bc")) bc = s.attribute("bc").value();
        if (bc == "transmission") S.bc = MCB_BC_TRANSMISSION;
        else if (bc == "reflective") S.bc = MCB_BC_REFLECTIVE;
        else if (bc == "vacuum") S.bc = MCB_BC_VACUUM;
        else { error = "[INPUT ERROR] unknown boundary condition " + bc; return false; }  // reference: uninitialised (quirk 18)
        const std::string type = s.name;
        if (type == "plane_x") { S.type = MCB_SURF_PLANE_X; S.p[0] = s.attribute("x").as_double(); }
        else if (type == "plane_y") { S.type = MCB_SURF_PLANE_Y; S.p[0] = s.attribute("y").as_double(); }
        else if (type == "plane_z") { S.type = MCB_SURF_PLANE_Z; S.p[0] = s.attribute("z").as_double(); }
        else if (type == "plane") {
            S.type = MCB_SURF_PLANE;
            const double a = s.attribute("a").as_double(), b = s.attribute("b").as_double(), c = s.attribute("c").as_double();
            S.p[0] = a; S.p[1] = b; S.p[2] = c; S.p[3] = s.attribute("d").as_double();
            const double L = 2.0 / (a * a + b * b + c * c);  // Geometry.cpp:19-22
            S.p[4] = L * a; S.p[5] = L * b; S.p[6] = L * c;
        } else if (type == "sphere") {
            S.type = MCB_SURF_SPHERE;
            S.p[0] = s.attribute("x").as_double(); S.p[1] = s.attribute("y").as_double(); S.p[2] = s.attribute("z").as_double();
            S.p[3] = s.attribute("r").as_double(); S.p[4] = S.p[3] * S.p[3];
        } else if (type == "cylinder_x") {
            S.type = MCB_SURF_CYL_X;
            S.p[0] = s.attribute("y").as_double(); S.p[1] = s.attribute("z").as_double();
            S.p[2] = s.attribute("r").as_double(); S.p[3] = S.p[2] * S.p[2];
        } else if (type == "cylinder_z") {
            S.type = MCB_SURF_CYL_Z;
            S.p[0] = s.attribute("x").as_double(); S.p[1] = s.attribute("y").as_double();
            S.p[2] = s.attribute("r").as_double(); S.p[3] = S.p[2] * S.p[2];
        } else {
            error = " unkown surface type " + type; return false;
        }
        surfaces.push_back(S);
        surface_names.push_back(s.attribute("name").value());
    }

    // ---- cells (setup.cpp:591-629) ----
    for (const XmlNode* c : top("cells").children("cell")) {
        mcb_cell C;
        std::memset(&C, 0, sizeof(C));
        std::string name = "Cell " + std::to_string(cells.size() + 1);
        if (c->attribute("name")) name = c->attribute("name").value();
        C.importance = 1.0;
        if (c->attribute("importance")) C.importance = c->attribute("importance").as_double();
        C.material = -1;
        if (c->attribute("material")) {
            C.material = find_name(material_names, c->attribute("material").value());
            if (C.material < 0) { error = "[INPUT_ERROR] Unknown material in cell"; return false; }
        }
        C.surf_begin = (int32_t)cell_surface.size();
        for (const XmlNode* s : c->children("surface")) {
            const int id = find_name(surface_names, s->attribute("name").value());
            if (id < 0) { error = "[INPUT_ERROR] Unknown surface"; return false; }
            cell_surface.push_back(id);
            cell_sense.push_back(s->attribute("sense").as_int());
        }
        C.surf_end = (int32_t)cell_surface.size();
        cells.push_back(C);
        cell_names.push_back(name);
    }

    // ---- estimators: only the first <estimators> block is read (setup.cpp:637-805, quirk 19) ----
    int64_t n_tallies = 0;
    for (const XmlNode* e : top("estimators").children("estimator")) {
        mcb_estimator E;
        std::memset(&E, 0, sizeof(E));
        std::string e_name = "Estimator #" + std::to_string(estimators.size() + 1);
        if (e->attribute("name")) e_name = e->attribute("name").value();
        copy_name(E.name, sizeof(E.name), e_name);
        std::string e_type = "TL";
        if (e->attribute("type")) e_type = e->attribute("type").value();
        int kernel;
        if (e_type == "TL") kernel = MCB_KERNEL_TRACK;
        else if (e_type == "C") kernel = MCB_KERNEL_COLLISION;
        else { error = "[ERROR] Unsupported score type in estimator " + e_name; return false; }
        if (p.tdmc_on) kernel = MCB_KERNEL_VELOCITY;  // every estimator of a time-dependent run scores w * v (setup.cpp:659-661)
        if (!e->attribute("scores")) { error = "[ERROR] There is no score in estimator " + e_name; return false; }
        E.score_begin = (int32_t)scores.size();
        std::istringstream iss(e->attribute("scores").value());
        for (std::string s; iss >> s;) {
            mcb_score S;
            std::memset(&S, 0, sizeof(S));
            copy_name(S.name, sizeof(S.name), s);
            if (s == "flux") S.score = MCB_SCORE_FLUX;
            else if (s == "absorption") S.score = MCB_SCORE_ABSORPTION;
            else if (s == "scatter") S.score = MCB_SCORE_SCATTER;
            else if (s == "capture") S.score = MCB_SCORE_CAPTURE;
            else if (s == "fission") S.score = MCB_SCORE_FISSION;
            else if (s == "nu-fission") S.score = MCB_SCORE_NU_FISSION;
            else if (s == "total") S.score = MCB_SCORE_TOTAL;
            else if (s == "cross") { kernel = MCB_KERNEL_NEUTRON; S.score = MCB_SCORE_FLUX; }  // kernel swap persists (setup.cpp:688-690)
            else { error = "[ERROR] Unsuported score type " + s + " in estimator " + e_name; return false; }
            S.kernel = kernel;
            scores.push_back(S);
        }
        E.n_scores = (int32_t)scores.size() - E.score_begin;

        E.filter_begin = (int32_t)filters.size();
        std::vector<double> grid;
        // TDMC filter: first in the index order (setup.cpp:703-708)
        if (e->child("tdmc")) {
            mcb_filter F;
            std::memset(&F, 0, sizeof(F));
            F.type = MCB_FILTER_TDMC;
            F.grid_begin = (int32_t)filter_grid.size(); F.grid_n = (int32_t)tdmc_time.size(); F.size = F.grid_n;
            filter_grid.insert(filter_grid.end(), tdmc_time.begin(), tdmc_time.end());
            filters.push_back(F);
        }
        // attach to geometries; their IDs form the next filter's grid (setup.cpp:709-741)
        for (const XmlNode* s : e->children("surface")) {
            const int id = find_name(surface_names, s->attribute("name").value());
            if (id < 0) { error = "[ERROR] Unknown surface label " + s->attribute("name").value() + " in estimator " + e_name; return false; }
            grid.push_back(id);
        }
        const bool has_surface = !grid.empty();
        for (const XmlNode* c : e->children("cell")) {
            const int id = find_name(cell_names, c->attribute("name").value());
            if (id < 0) { error = "[ERROR] Unknown cell label " + c->attribute("name").value() + " in estimator " + e_name; return false; }
            grid.push_back(id);
        }
        if (has_surface && e->child("cell")) {
            error = "unsupported: estimator " + e_name + " attached to both surfaces and cells"; return false;
        }
        mcb_filter F;
        std::memset(&F, 0, sizeof(F));
        if (e->child("surface")) { F.type = MCB_FILTER_SURFACE; E.attach = MCB_ATTACH_SURFACE; }
        else if (e->child("cell")) { F.type = MCB_FILTER_CELL; E.attach = (e_type == "TL") ? MCB_ATTACH_CELL_TL : MCB_ATTACH_CELL_C; }
        else { error = "[ERROR] Estimator " + e_name + " needs to be attached somewhere"; return false; }
        F.grid_begin = (int32_t)filter_grid.size(); F.grid_n = (int32_t)grid.size(); F.size = F.grid_n;
        filter_grid.insert(filter_grid.end(), grid.begin(), grid.end());
        filters.push_back(F);

        for (const XmlNode* f : e->children("filter")) {
            const std::string f_name = f->attribute("type").value();
            if (!parse_filter_grid(*f, grid, "estimator " + e_name, error)) return false;
            if (f_name == "energy") F.type = MCB_FILTER_ENERGY;
            else if (f_name == "time") F.type = MCB_FILTER_TIME;
            else { error = "[ERROR] Unknown filter type for estimator " + e_name; return false; }
            if (grid.size() < 2) { error = "[ERROR] filter grid of estimator " + e_name + " needs two points"; return false; }
            F.grid_begin = (int32_t)filter_grid.size(); F.grid_n = (int32_t)grid.size(); F.size = F.grid_n - 1;
            filter_grid.insert(filter_grid.end(), grid.begin(), grid.end());
            filters.push_back(F);
        }
        E.n_filters = (int32_t)filters.size() - E.filter_begin;
        int64_t nt = E.n_scores;
        for (int i = 0; i < E.n_filters; i++) nt *= filters[E.filter_begin + i].size;
        E.tally_begin = (int32_t)n_tallies;
        E.n_tallies = (int32_t)nt;
        n_tallies += nt;
        estimators.push_back(E);
    }
    p.n_tallies = n_tallies;

    // ---- TRMM tally set (setup.cpp:811-1013) ----
    if (doc.child("trmm")) {
        trmm_present = true;
        if (!p.ksearch) { error = "[ERROR] TRMM should be run in ksearch mode"; return false; }
        if (!(flags & DECK_IGNORE_TRMM)) {
            const XmlNode& T = *doc.child("trmm");
            // nine estimators in this order (setup.cpp:828-842,1004-1012): simple, scatter, fission_prompt, delayed 1..6
            struct Spec { std::string name; int simulate; std::vector<mcb_score> sc; };
            auto mk = [](const std::string& n, int score, int kernel, int group) {
                mcb_score S;
                std::memset(&S, 0, sizeof(S));
                copy_name(S.name, sizeof(S.name), n);
                S.score = score; S.kernel = kernel; S.group = group;
                return S;
            };
            std::vector<Spec> specs;
            Spec simple{"TRM_simple", MCB_SIM_NONE, {}};
            simple.sc.push_back(mk("collision", MCB_SCORE_TOTAL, MCB_KERNEL_TRACK_VELOCITY, 0));                  // setup.cpp:848-851
            simple.sc.push_back(mk("flux", MCB_SCORE_FLUX, MCB_KERNEL_TRACK, 0));                                 // :853-856
            for (int i = 0; i < 6; i++)                                                                            // :858-864
                simple.sc.push_back(mk("NuFissionDelayed_" + std::to_string(i + 1), MCB_SCORE_NU_FISSION_DELAYED_OLD, MCB_KERNEL_TRACK, i));
            for (int i = 0; i < 6; i++)                                                                            // :866-873
                simple.sc.push_back(mk("NuFissionDelayedLambda_" + std::to_string(i + 1), MCB_SCORE_NU_FISSION_DELAYED_DECAY_OLD, MCB_KERNEL_TRACK, i));
            simple.sc.push_back(mk("inverse_speed", MCB_SCORE_INVERSE_VELOCITY, MCB_KERNEL_TRACK, 0));            // :875-878
            specs.push_back(simple);
            specs.push_back({"TRM_matrix_scatter", MCB_SIM_SCATTER, {mk("InScatter", MCB_SCORE_SCATTER_OLD, MCB_KERNEL_TRACK_VELOCITY, 0)}});  // :890-892
            specs.push_back({"TRM_matrix_fission_prompt", MCB_SIM_FISSION_PROMPT,
                             {mk("NuFissionPrompt", MCB_SCORE_NU_FISSION_PROMPT_OLD, MCB_KERNEL_TRACK_VELOCITY, 0)}});                         // :894-897
            for (int i = 0; i < 6; i++)                                                                                                            // :880-888
                specs.push_back({"TRMM_matrix_fission_delayed_" + std::to_string(i + 1), MCB_SIM_FISSION_DELAYED + i,
                                 {mk("NuFissionDelayedEmission_" + std::to_string(i + 1), MCB_SCORE_NU_FISSION_DELAYED_OLD, MCB_KERNEL_TRACK_VELOCITY, i)}});
            // filters: the cell filter, then per <filter>: energy -> [energy_initial][energy] for the matrix estimators
            std::vector<double> cell_grid;
            for (const XmlNode* c : T.children("cell")) {
                const int id = find_name(cell_names, c->attribute("name").value());
                if (id < 0) { error = "[ERROR] Unknown cell label " + c->attribute("name").value() + " in trmm"; return false; }
                cell_grid.push_back(id);
            }
            struct FSpec { int type; std::vector<double> grid; };
            std::vector<FSpec> extra;
            for (const XmlNode* f : T.children("filter")) {
                std::vector<double> grid;
                if (!parse_filter_grid(*f, grid, "trmm", error)) return false;
                if (!f->attribute("type")) { error = "[ERROR] Need filter type for trmm"; return false; }
                const std::string f_name = f->attribute("type").value();
                if (f_name == "energy") extra.push_back({MCB_FILTER_ENERGY, grid});
                else if (f_name == "time") extra.push_back({MCB_FILTER_TIME, grid});
                else { error = "[ERROR] Unknown filter type for trmm"; return false; }
                if (grid.size() < 2) { error = "[ERROR] filter grid of trmm needs two points"; return false; }
            }
            for (const Spec& sp : specs) {
                mcb_estimator E;
                std::memset(&E, 0, sizeof(E));
                copy_name(E.name, sizeof(E.name), sp.name);
                E.attach = MCB_ATTACH_CELL_TL;
                E.simulate = sp.simulate;
                E.score_begin = (int32_t)scores.size();
                scores.insert(scores.end(), sp.sc.begin(), sp.sc.end());
                E.n_scores = (int32_t)sp.sc.size();
                E.filter_begin = (int32_t)filters.size();
                auto add_filter = [&](int type, const std::vector<double>& grid) {
                    mcb_filter F;
                    std::memset(&F, 0, sizeof(F));
                    F.type = type;
                    F.grid_begin = (int32_t)filter_grid.size(); F.grid_n = (int32_t)grid.size();
                    F.size = (type == MCB_FILTER_CELL || type == MCB_FILTER_SURFACE) ? F.grid_n : F.grid_n - 1;
                    filter_grid.insert(filter_grid.end(), grid.begin(), grid.end());
                    filters.push_back(F);
                };
                add_filter(MCB_FILTER_CELL, cell_grid);
                for (const FSpec& fs : extra) {
                    if (fs.type == MCB_FILTER_ENERGY && sp.simulate != MCB_SIM_NONE) add_filter(MCB_FILTER_ENERGY_OLD, fs.grid);
                    add_filter(fs.type, fs.grid);
                }
                E.n_filters = (int32_t)filters.size() - E.filter_begin;
                int64_t nt = E.n_scores;
                for (int i = 0; i < E.n_filters; i++) nt *= filters[E.filter_begin + i].size;
                E.tally_begin = (int32_t)n_tallies;
                E.n_tallies = (int32_t)nt;
                n_tallies += nt;
                estimators.push_back(E);
            }
            p.n_tallies = n_tallies;
            trmm_built = true;
        }
    }

    // ---- sources (setup.cpp:1020-1066) ----
    for (const XmlNode& s : top("sources").kids) {
        mcb_source S;
        std::memset(&S, 0, sizeof(S));
        S.prob = 1.0;
        S.dir_kind = MCB_DIR_ISOTROPIC;        // default isotropic
        S.energy.kind = MCB_DIST_DELTA;        // default 2 MeV
        S.energy.a = 2e6;
        if (s.attribute("probability")) S.prob = s.attribute("probability").as_double();
        bool ok = true;
        if (s.attribute("direction")) {
            const NamedPoint* d = nullptr;
            for (auto& q : distp) if (q.name == s.attribute("direction").value()) { d = &q; break; }
            if (!d) ok = false;
            else { S.dir_kind = d->kind; std::memcpy(S.dir, d->xyz, sizeof(S.dir)); std::memcpy(S.dir_xyz, d->comp, sizeof(S.dir_xyz)); }
        }
        if (s.attribute("energy")) {
            const NamedDist1* d = nullptr;
            for (auto& q : dist1) if (q.name == s.attribute("energy").value()) { d = &q; break; }
            if (!d) ok = false;
            else S.energy = d->d;
        }
        if (!ok) { error = "[ERROR] unknown direction distribution in source."; return false; }
        if (s.name == "point") {
            S.pos[0] = s.attribute("x").as_double(); S.pos[1] = s.attribute("y").as_double(); S.pos[2] = s.attribute("z").as_double();
        } else if (s.name == "source" && s.attribute("position")) {
            // superset of the reference: <source position="name"> with a delta point distribution,
            // as written in examples/HEU_sphere_criticality/input.xml:46 (which the reference itself rejects)
            const NamedPoint* d = nullptr;
            for (auto& q : distp) if (q.name == s.attribute("position").value()) { d = &q; break; }
            if (!d || d->kind != MCB_DIR_DELTA) { error = "[INPUT ERROR] <source position=...> needs a delta point distribution"; return false; }
            std::memcpy(S.pos, d->xyz, sizeof(S.pos));
        } else if (s.name == "disk_z") {
            // superset of the reference: examples/sphere_detection/input.xml:106 (the deck of the MCNP6 integral test,
            // test/test_integral_Simulator.cpp:10-19) uses it, setup.cpp:1051-1063 rejects it.  Sampling: mcb200.h
            S.kind = MCB_SRC_DISK_Z;
            S.pos[0] = s.attribute("x").as_double(); S.pos[1] = s.attribute("y").as_double(); S.pos[2] = s.attribute("z").as_double();
            S.radius = s.attribute("r").as_double();
            if (!(S.radius > 0.0)) { error = "[INPUT ERROR] <disk_z> needs a radius r > 0"; return false; }
        } else {
            error = "[INPUT ERROR] Unknown source type: " + s.name; return false;
        }
        S.cell = search_cell(S.pos[0], S.pos[1], S.pos[2]);
        if (S.cell < 0) {
            std::ostringstream o;
            o << "[WARNING] A particle is lost:\n( x, y, z )  (" << S.pos[0] << ", " << S.pos[1] << ", " << S.pos[2] << " )";
            error = o.str(); return false;
        }
        sources.push_back(S);
    }
    if (sources.empty()) { error = "[ERROR] Source bank is empty..."; return false; }
    return true;
}

const mcb_problem* Deck::view()
{
    p.n_nuclides = (int32_t)nuclides.size();
    p.n_materials = (int32_t)material_names.size();
    p.nuclides = nuclides.data();
    p.xs_rows = xs_rows.data();
    p.n_xs_rows = (int64_t)(xs_rows.size() / MCB_XS_ROW);
    p.delayed_data = delayed_data.data();
    p.n_delayed_data = (int64_t)delayed_data.size();
    p.mat_begin = mat_begin.data();
    p.mat_nuclide = mat_nuclide.data();
    p.mat_density = mat_density.data();
    p.n_surfaces = (int32_t)surfaces.size();
    p.n_cells = (int32_t)cells.size();
    p.surfaces = surfaces.data();
    p.cells = cells.data();
    p.cell_surface = cell_surface.data();
    p.cell_sense = cell_sense.data();
    p.n_cell_surface = (int32_t)cell_surface.size();
    p.n_sources = (int32_t)sources.size();
    p.sources = sources.data();
    p.n_estimators = (int32_t)estimators.size();
    p.n_scores = (int32_t)scores.size();
    p.n_filters = (int32_t)filters.size();
    p.n_filter_grid = (int32_t)filter_grid.size();
    p.estimators = estimators.data();
    p.scores = scores.data();
    p.filters = filters.data();
    p.filter_grid = filter_grid.data();
    p.entropy_grid = entropy_grid.data();
    p.n_tdmc = (int32_t)tdmc_time.size();
    p.tdmc_time = tdmc_time.data();
    p.tdmc_interval = tdmc_interval.data();
    return &p;
}

}  // namespace mcb
