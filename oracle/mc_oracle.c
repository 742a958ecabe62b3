/* mc_oracle.c — TEST INFRASTRUCTURE ONLY (see mc_oracle.h).
 *
 * Plain-C restatement of the reference's history-by-history transport
 * (ilhamv/MC-old).  Each function cites the reference file:line it follows.
 * Random-number draw ORDER follows the g++ -O3 build of the reference exactly
 * (including the right-to-left evaluation of constructor arguments, SURVEY
 * App. D-4), so that in MCO_RNG_GLOBAL mode the whole run is bit-identical to
 * oracle/_ref/MC_ref.  Compile without FMA contraction (-ffp-contract=off).
 *
 * Parity: pinned (every fixture of tests/golden/ comes from the compiled reference) with ONE exception, stated here
 * as the rules ask: the <disk_z> source (deck_source, MCB_SRC_DISK_Z) has no reference implementation - the
 * reference's loader rejects the element (setup.cpp:1051-1063) - so that sampling is "parity unpinned"; it is
 * anchored on the MCNP6 number of the reference's own integral test (test/test_integral_Simulator.cpp:10-19).
 */
#include "mc_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* include/Constants.h:7-13 */
static const double EPSILON_float = 1.1920928955078125e-07;
static const double MAX_float = 3.4028234663852886e+38;
#define PI_ (acos(-1.0))

/* ------------------------------------------------------------------------------------------ */
/* Random (src/Random.cpp)                                                                    */
/* ------------------------------------------------------------------------------------------ */
#define RN_MULT 3512401965023503517ULL
#define RN_MASK ((~0ULL) >> 1)
#define RN_STRIDE 152917ULL

uint64_t mco_lcg_next(uint64_t seed) { return (RN_MULT * seed) & RN_MASK; } /* Random.cpp:123 */

/* RN_skip_ahead (Random.cpp:130-149), RN_ADD = 0 */
uint64_t mco_lcg_skip(uint64_t seed, uint64_t n)
{
    uint64_t nskip = n & RN_MASK;
    uint64_t gen = 1, g = RN_MULT;
    for (; nskip; nskip >>= 1) {
        if (nskip & 1) gen = (gen * g) & RN_MASK;
        g = (g * g) & RN_MASK;
    }
    return (gen * seed) & RN_MASK;
}

/* ------------------------------------------------------------------------------------------ */
/* types                                                                                      */
/* ------------------------------------------------------------------------------------------ */
typedef struct particle {
    double pos[3], dir[3];
    int alive;
    double E, E_old, speed, w, t, t_old;
    int cell, cell_old, surface_old;
    int tdmc; /* index of the next census time (Particle::tdmc, Particle.cpp:29,82) */
    uint64_t rng; /* MCO_RNG_HISTORY: this particle's stream */
} particle;

typedef struct tally { double hist, sum, squared, mean, uncer; } tally; /* include/Estimator.h:257-269 */

struct mco_ctx {
    const mcb_problem* p;
    int rng_mode, pick_mode;
    uint64_t seed;       /* MCO_RNG_GLOBAL: the one global stream */
    uint64_t n_draws;
    uint64_t shard_begin, shard_count;
    /* simulator state (include/simulator.h:100-127) */
    uint64_t icycle;
    int tally_on;
    double k;
    uint64_t Ntrack, Ncollision;
    /* banks (simulator.h:27-30) */
    particle* Pbank; size_t Pn, Pcap;
    particle* Fbank; size_t Fn, Fcap;   /* fission bank being filled */
    particle* Sbank; size_t Sn;         /* sample bank of SourceDelta sites (cycle > 0) */
    int S_is_deck;                      /* first cycle / fixed source: sample the deck's sources */
    int src_lost;                       /* a <disk_z> source particle fell outside every cell */
    double* cdf; size_t cdf_n;          /* SourceBank::p */
    /* estimators */
    tally* tallies;
    /* EstimatorK (include/Estimator.h:466-502) */
    double k_sum_C, k_sum_TL, k_sq_C, k_sq_TL, H_sum, k_C, k_TL;
    uint64_t Navg;
    double mean_accumulator, uncer_sq_accumulator;
    uint64_t cyc_tracks0, cyc_coll0, cyc_hist;
    /* ShannonEntropy (include/Entropy.h:19-36) */
    double* ent_p; int ent_I;
    uint32_t child_counter; /* secondaries spawned by the current event (stream derivation) */
    double *hist_kC, *hist_kTL; size_t hist_n; /* per-history k scores of the last cycle (shard-local order) */
};

/* Urand (Random.cpp:121-126) — global stream or the particle's own */
static double urand(mco_ctx* c, particle* P)
{
    uint64_t* s = (c->rng_mode == MCO_RNG_GLOBAL || !P) ? &c->seed : &P->rng;
    *s = (RN_MULT * (*s)) & RN_MASK;
    c->n_draws++;
    return (double)((*s) * (1. / (double)(1ULL << 63)));
}
static double urand_raw(uint64_t* s)
{
    *s = (RN_MULT * (*s)) & RN_MASK;
    return (double)((*s) * (1. / (double)(1ULL << 63)));
}

/* ------------------------------------------------------------------------------------------ */
/* Algorithm (src/Algorithm.cpp)                                                              */
/* ------------------------------------------------------------------------------------------ */
int mco_binary_search(double x, const double* v, int n) /* Algorithm.cpp:46-64 */
{
    int left = 0, right = n - 1, mid;
    while (left <= right) {
        mid = (left + right) / 2;
        if (v[mid] < x) left = mid + 1; else right = mid - 1;
    }
    return right;
}
/* same search over column 0 of xs rows */
static int row_search(double x, const double* rows, int n)
{
    int left = 0, right = n - 1, mid;
    while (left <= right) {
        mid = (left + right) / 2;
        if (rows[(size_t)mid * MCB_XS_ROW] < x) left = mid + 1; else right = mid - 1;
    }
    return right;
}
double mco_interpolate(double x, double x1, double x2, double y1, double y2) /* Algorithm.cpp:103-105 */
{
    return (x - x2) / (x1 - x2) * y1 + (x - x1) / (x2 - x1) * y2;
}
double mco_geometry_quad(double a, double b, double c) /* Algorithm.cpp:16-38 */
{
    const double D = b * b - 4.0 * a * c;
    if (D <= 0.0) return MAX_float;
    {
        const double sqrtD = sqrt(D);
        const double ai = 0.5 / a;
        double r1 = ai * (-1.0 * b - sqrtD);
        double r2 = ai * (-1.0 * b + sqrtD);
        if (r1 < 0) r1 = MAX_float;
        if (r2 < 0) r2 = MAX_float;
        return fmin(r1, r2);
    }
}
void mco_scatter_direction(const double* d, double mu0, double xi, double* f) /* Algorithm.cpp:67-101 */
{
    const double azi = 2.0 * PI_ * xi;
    const double cos_azi = cos(azi);
    const double sin_azi = sin(azi);
    const double Ac = sqrt(1.0 - mu0 * mu0);
    if (d[2] != 1.0) {
        const double B = sqrt(1.0 - d[2] * d[2]);
        const double C = Ac / B;
        f[0] = d[0] * mu0 + (d[0] * d[2] * cos_azi - d[1] * sin_azi) * C;
        f[1] = d[1] * mu0 + (d[1] * d[2] * cos_azi + d[0] * sin_azi) * C;
        f[2] = d[2] * mu0 - cos_azi * Ac * B;
    } else {
        const double B = sqrt(1.0 - d[1] * d[1]);
        const double C = Ac / B;
        f[0] = d[0] * mu0 + (d[0] * d[1] * cos_azi - d[2] * sin_azi) * C;
        f[2] = d[2] * mu0 + (d[2] * d[1] * cos_azi + d[0] * sin_azi) * C;
        f[1] = d[1] * mu0 - cos_azi * Ac * B;
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Particle (src/Particle.cpp)                                                                */
/* ------------------------------------------------------------------------------------------ */
double mco_speed_of_energy(double E) { return 13831.5926439 * sqrt(E) * 100.0; } /* Particle.cpp:46 */
double mco_energy_of_speed(double v) { return 5.2270376e-13 * v * v; }           /* Particle.cpp:53 */

static void p_set_energy(particle* P, double E) /* Particle.cpp:42-48 */
{
    P->E_old = P->E; P->E = E; P->speed = mco_speed_of_energy(E);
}
static void p_set_speed(particle* P, double v) /* Particle.cpp:49-56 */
{
    P->speed = v; P->E_old = P->E; P->E = mco_energy_of_speed(v);
}
static void p_move(particle* P, double d) /* Particle.cpp:66-76 */
{
    P->pos[0] += P->dir[0] * d; P->pos[1] += P->dir[1] * d; P->pos[2] += P->dir[2] * d;
    P->t_old = P->t; P->t += d / P->speed;
}
static void p_kill(particle* P) { P->alive = 0; P->w = 0.0; } /* Particle.cpp:77-81 */
static void p_set_cell(particle* P, int cnew) { P->cell_old = P->cell; P->cell = cnew; } /* Particle.cpp:37-41 */
/* Particle constructor (include/Particle.h:30-34). E_old is uninitialised in the reference; we define E_old = E */
static particle p_make(const double* pos, const double* dir, double E, double t, double w, int cell)
{
    particle P;
    memset(&P, 0, sizeof(P));
    memcpy(P.pos, pos, sizeof(P.pos)); memcpy(P.dir, dir, sizeof(P.dir));
    P.alive = 1; P.t = t; P.w = w; P.cell = cell; P.cell_old = cell; P.surface_old = -1;
    P.E = E; p_set_energy(&P, E);
    return P;
}

/* ------------------------------------------------------------------------------------------ */
/* Geometry (src/Geometry.cpp)                                                                */
/* ------------------------------------------------------------------------------------------ */
double mco_surface_eval(const mcb_problem* p, int s, const double* q) /* Geometry.cpp:29-69 */
{
    const mcb_surface* S = &p->surfaces[s];
    switch (S->type) {
    case MCB_SURF_PLANE_X: return q[0] - S->p[0];
    case MCB_SURF_PLANE_Y: return q[1] - S->p[0];
    case MCB_SURF_PLANE_Z: return q[2] - S->p[0];
    case MCB_SURF_PLANE: return S->p[0] * q[0] + S->p[1] * q[1] + S->p[2] * q[2] - S->p[3];
    case MCB_SURF_SPHERE: {
        const double x_t = q[0] - S->p[0], y_t = q[1] - S->p[1], z_t = q[2] - S->p[2];
        return x_t * x_t + y_t * y_t + z_t * z_t - S->p[4];
    }
    case MCB_SURF_CYL_X: {
        const double y_t = q[1] - S->p[0], z_t = q[2] - S->p[1];
        return y_t * y_t + z_t * z_t - S->p[3];
    }
    case MCB_SURF_CYL_Y: { /* Geometry.cpp:58-63: second coordinate is p.y - z0 */
        const double x_t = q[0] - S->p[0], z_t = q[1] - S->p[1];
        return x_t * x_t + z_t * z_t - S->p[3];
    }
    default: {
        const double x_t = q[0] - S->p[0], y_t = q[1] - S->p[1];
        return x_t * x_t + y_t * y_t - S->p[3];
    }
    }
}
static double axis_plane_distance(double loc, double pos, double dir) /* Geometry.cpp:76-126 */
{
    if (fabs(dir) > EPSILON_float) {
        const double dist = (loc - pos) / dir;
        if (dist > 0.0) return dist;
        return MAX_float;
    }
    return MAX_float;
}
double mco_surface_distance(const mcb_problem* p, int s, const double* q, const double* u) /* Geometry.cpp:76-188 */
{
    const mcb_surface* S = &p->surfaces[s];
    switch (S->type) {
    case MCB_SURF_PLANE_X: return axis_plane_distance(S->p[0], q[0], u[0]);
    case MCB_SURF_PLANE_Y: return axis_plane_distance(S->p[0], q[1], u[1]);
    case MCB_SURF_PLANE_Z: return axis_plane_distance(S->p[0], q[2], u[2]);
    case MCB_SURF_PLANE: {
        const double denom = S->p[0] * u[0] + S->p[1] * u[1] + S->p[2] * u[2];
        if (fabs(denom) > EPSILON_float) {
            const double dist = (S->p[3] - S->p[0] * q[0] - S->p[1] * q[1] - S->p[2] * q[2]) / denom;
            if (dist > 0.0) return dist;
            return MAX_float;
        }
        return MAX_float;
    }
    case MCB_SURF_SPHERE: {
        const double b = 2.0 * ((q[0] - S->p[0]) * u[0] + (q[1] - S->p[1]) * u[1] + (q[2] - S->p[2]) * u[2]);
        return mco_geometry_quad(1.0, b, mco_surface_eval(p, s, q));
    }
    case MCB_SURF_CYL_X: {
        const double a = 1.0 - u[0] * u[0];
        const double b = 2.0 * ((q[1] - S->p[0]) * u[1] + (q[2] - S->p[1]) * u[2]);
        return mco_geometry_quad(a, b, mco_surface_eval(p, s, q));
    }
    case MCB_SURF_CYL_Y: {
        const double a = 1.0 - u[1] * u[1];
        const double b = 2.0 * ((q[0] - S->p[0]) * u[0] + (q[2] - S->p[1]) * u[2]);
        return mco_geometry_quad(a, b, mco_surface_eval(p, s, q));
    }
    default: {
        const double a = 1.0 - u[2] * u[2];
        const double b = 2.0 * ((q[0] - S->p[0]) * u[0] + (q[1] - S->p[1]) * u[1]);
        return mco_geometry_quad(a, b, mco_surface_eval(p, s, q));
    }
    }
}
void mco_surface_reflect(const mcb_problem* p, int s, double* d) /* Geometry.cpp:195-222 */
{
    const mcb_surface* S = &p->surfaces[s];
    switch (S->type) {
    case MCB_SURF_PLANE_X: d[0] = -d[0]; break;
    case MCB_SURF_PLANE_Y: d[1] = -d[1]; break;
    case MCB_SURF_PLANE_Z: d[2] = -d[2]; break;
    case MCB_SURF_PLANE: {
        const double K = (S->p[0] * d[0] + S->p[1] * d[1] + S->p[2] * d[2]);
        const double qx = d[0] - S->p[4] * K, qy = d[1] - S->p[5] * K, qz = d[2] - S->p[6] * K;
        d[0] = qx; d[1] = qy; d[2] = qz;
        break;
    }
    default: break; /* sphere / cylinders: no-op */
    }
}
int mco_search_cell(const mcb_problem* p, const double* q) /* general.cpp:13-34 */
{
    int c, i;
    for (c = 0; c < p->n_cells; c++) {
        int inside = 1;
        for (i = p->cells[c].surf_begin; i < p->cells[c].surf_end; i++) {
            if (mco_surface_eval(p, p->cell_surface[i], q) * p->cell_sense[i] < 0) { inside = 0; break; }
        }
        if (inside) return c;
    }
    return -1;
}
int mco_surface_intersect(const mcb_problem* p, int cell, const double* q, const double* u, double* dist_out)
{ /* general.cpp:54-67 */
    double dist = MAX_float;
    int S = -1, i;
    for (i = p->cells[cell].surf_begin; i < p->cells[cell].surf_end; i++) {
        const double d = mco_surface_distance(p, p->cell_surface[i], q, u);
        if (d < dist) { dist = d; S = p->cell_surface[i]; }
    }
    *dist_out = dist;
    return S;
}

/* ------------------------------------------------------------------------------------------ */
/* XSec / Nuclide / Material (src/XSec.cpp, src/Nuclide.cpp, src/Material.cpp)                */
/* ------------------------------------------------------------------------------------------ */
enum { X_S = 0, X_C, X_F, X_A, X_T, X_NU, X_BETA };

/* value of derived table `col` at grid point i (setup.cpp:367-375: sigmaA=c3+c4, sigmaT=c2+c3+c4) */
static double grid_value(const double* r, int col)
{
    switch (col) {
    case X_S: return r[1];
    case X_C: return r[2];
    case X_F: return r[3];
    case X_A: return r[2] + r[3];
    case X_T: return r[1] + r[2] + r[3];
    case X_NU: return r[4];
    default: return r[5];
    }
}
/* Nuclide::checkE + XSTable::xs (Nuclide.cpp:18-24, XSec.cpp:9-40) without the memo (same values) */
static double micro(const mcb_problem* p, int n, int col, double E)
{
    const mcb_nuclide* N = &p->nuclides[n];
    const double* rows = p->xs_rows + (size_t)N->row_begin * MCB_XS_ROW;
    const int idx = row_search(E, rows, N->n_rows);
    if (idx == N->n_rows - 1) return grid_value(rows + (size_t)(N->n_rows - 1) * MCB_XS_ROW, col);
    if (idx == -1) return grid_value(rows, col);
    {
        const double* r1 = rows + (size_t)idx * MCB_XS_ROW;
        const double* r2 = r1 + MCB_XS_ROW;
        return mco_interpolate(E, r1[0], r2[0], grid_value(r1, col), grid_value(r2, col));
    }
}
static double micro_nusigmaF(const mcb_problem* p, int n, double E) /* Nuclide.cpp:57-61 */
{
    return micro(p, n, X_F, E) * micro(p, n, X_NU, E);
}
/* Material::Sigma* (Material.cpp:18-65): sum in deck order from 0.0 */
static double macro_(const mcb_problem* p, int m, int col, double E)
{
    double sum = 0.0;
    int i;
    for (i = p->mat_begin[m]; i < p->mat_begin[m + 1]; i++) sum += micro(p, p->mat_nuclide[i], col, E) * p->mat_density[i];
    return sum;
}
static double macro_nuSigmaF(const mcb_problem* p, int m, double E)
{
    double sum = 0.0;
    int i;
    for (i = p->mat_begin[m]; i < p->mat_begin[m + 1]; i++) sum += micro_nusigmaF(p, p->mat_nuclide[i], E) * p->mat_density[i];
    return sum;
}
/* Nuclide::nusigmaF_prompt / nusigmaF_delayed (Nuclide.cpp:62-73): (1-beta)*sf*nu and beta*fraction_i*sf*nu, left to right */
static double micro_nusigmaF_prompt(const mcb_problem* p, int n, double E)
{
    return (1.0 - micro(p, n, X_BETA, E)) * micro(p, n, X_F, E) * micro(p, n, X_NU, E);
}
static double micro_nusigmaF_delayed(const mcb_problem* p, int n, double E, int g)
{
    return micro(p, n, X_BETA, E) * p->nuclides[n].fraction[g] * micro(p, n, X_F, E) * micro(p, n, X_NU, E);
}
/* channel kinds: 0 scatter, 1 nu-fission, 2 prompt nu-fission, 3+g delayed nu-fission of group g */
static double micro_kind(const mcb_problem* p, int n, int kind, double E)
{
    if (kind == 0) return micro(p, n, X_S, E);
    if (kind == 1) return micro_nusigmaF(p, n, E);
    if (kind == 2) return micro_nusigmaF_prompt(p, n, E);
    return micro_nusigmaF_delayed(p, n, E, kind - 3);
}
/* Material::SigmaS / nuSigmaF / nuSigmaF_prompt / nuSigmaF_delayed (Material.cpp:26-82) */
static double macro_kind(const mcb_problem* p, int m, int kind, double E)
{
    double sum = 0.0;
    int i;
    for (i = p->mat_begin[m]; i < p->mat_begin[m + 1]; i++) sum += micro_kind(p, p->mat_nuclide[i], kind, E) * p->mat_density[i];
    return sum;
}
/* Material::nuSigmaF_delayed_decay (Material.cpp:83-91) */
static double macro_delayed_decay(const mcb_problem* p, int m, double E, int g)
{
    double sum = 0.0;
    int i;
    for (i = p->mat_begin[m]; i < p->mat_begin[m + 1]; i++) {
        const int n = p->mat_nuclide[i];
        sum += micro_nusigmaF_delayed(p, n, E, g) / p->nuclides[n].lambda[g] * p->mat_density[i];
    }
    return sum;
}
/* Material::nuclide_scatter / nuclide_nufission / _prompt / _delayed (Material.cpp:106-146); -1 = nullptr */
static int select_nuclide(const mcb_problem* p, int m, int kind, double E, double xi)
{
    const double u = macro_kind(p, m, kind, E) * xi;
    double s = 0.0;
    int i;
    for (i = p->mat_begin[m]; i < p->mat_begin[m + 1]; i++) {
        const int n = p->mat_nuclide[i];
        s += micro_kind(p, n, kind, E) * p->mat_density[i];
        if (s > u) return n;
    }
    return -1;
}

void mco_xs_lookup(const mcb_problem* p, int m, const double* E, int64_t n, double* out5)
{
    int64_t i;
    for (i = 0; i < n; i++) {
        out5[5 * i + 0] = macro_(p, m, X_T, E[i]);
        out5[5 * i + 1] = macro_(p, m, X_S, E[i]);
        out5[5 * i + 2] = macro_(p, m, X_C, E[i]);
        out5[5 * i + 3] = macro_(p, m, X_F, E[i]);
        out5[5 * i + 4] = macro_nuSigmaF(p, m, E[i]);
    }
}
void mco_select_channel(const mcb_problem* p, int m, int kind, const double* E, const double* xi, int64_t n, int32_t* out)
{
    int64_t i;
    for (i = 0; i < n; i++) out[i] = select_nuclide(p, m, kind, E[i], xi[i]);
}
void mco_beta(const mcb_problem* p, int nuc, const double* E, int64_t n, double* out)
{
    int64_t i;
    for (i = 0; i < n; i++) out[i] = micro(p, nuc, X_BETA, E[i]);
}

/* ------------------------------------------------------------------------------------------ */
/* Distribution / Reaction (src/Distribution.cpp, src/Reaction.cpp)                           */
/* ------------------------------------------------------------------------------------------ */
typedef double (*draw_fn)(void* a, void* b);
typedef struct rng_ref { mco_ctx* c; particle* P; uint64_t* raw; } rng_ref;
static double draw(rng_ref* r) { return r->raw ? urand_raw(r->raw) : urand(r->c, r->P); }

/* DistributionWatt::sample (Distribution.cpp:34-73) */
static double watt_sample(const double* va, const double* vb, const double* vg, double E, rng_ref* r)
{
    double a, b, g, xi, C, Eout;
    if (E <= 1.0) { a = va[0]; b = vb[0]; g = vg[0]; }
    else if (E <= 1.0e6) {
        a = mco_interpolate(E, 1.0, 1.0e6, va[0], va[1]);
        b = mco_interpolate(E, 1.0, 1.0e6, vb[0], vb[1]);
        g = mco_interpolate(E, 1.0, 1.0e6, vg[0], vg[1]);
    } else {
        a = mco_interpolate(E, 1.0e6, 14.0e6, va[1], va[2]);
        b = mco_interpolate(E, 1.0e6, 14.0e6, vb[1], vb[2]);
        g = mco_interpolate(E, 1.0e6, 14.0e6, vg[1], vg[2]);
    }
    do {
        double l2;
        xi = draw(r);
        Eout = -a * g * log(xi);
        l2 = log(draw(r));
        C = (1.0 - g) * (1.0 - log(xi)) - l2;
    } while (C * C > b * Eout);
    return (Eout * 1.0e6);
}
/* DistributionIsotropicDirection::sample (Distribution.cpp:78-92): x is the polar axis */
static void isotropic_direction(rng_ref* r, double* d)
{
    const double mu = 2.0 * draw(r) - 1.0;
    const double azi = 2.0 * PI_ * draw(r);
    const double c = sqrt(1.0 - mu * mu);
    d[1] = cos(azi) * c;
    d[2] = sin(azi) * c;
    d[0] = mu;
}
/* scalar distributions: Delta / Uniform (Distribution.cpp:30-33) / Watt at E = 0 */
static double dist1_sample(const mcb_dist1* d, rng_ref* r)
{
    switch (d->kind) {
    case MCB_DIST_DELTA: return d->a;
    case MCB_DIST_UNIFORM: return d->a + draw(r) * (d->b - d->a);
    default: return watt_sample(d->watt_a, d->watt_b, d->watt_g, 0.0, r);
    }
}

/* ReactionScatter::sample (Reaction.cpp:27-118): free-gas elastic scatter, isotropic in COM */
static void scatter_sample(double A, particle* P, rng_ref* r)
{
    const double mu0 = 2.0 * draw(r) - 1.0; /* DistributionIsotropicScatter (Distribution.cpp:74-77) */
    double V_tilda, mu_tilda, accept;
    const double beta = sqrt(2.0659834e-11 * A);
    const double y = beta * P->speed;
    const double PI_sqrt = sqrt(PI_), PI_half = 0.5 * PI_;
    double nuclide_dir[3], V_lab[3], v_lab[3], u[3], v_c[3], dir_c[3], dir_cNew[3], speed_c;
    do {
        double x;
        if (draw(r) < 2.0 / (2.0 + PI_sqrt * y)) {
            const double r1 = draw(r), r2 = draw(r);
            x = sqrt(-log(r1 * r2));
        } else {
            const double cos_arg = PI_half * draw(r);
            const double cos_val = cos(cos_arg);
            /* -log(Urand()) - log(Urand())*cos_val*cos_val : g++ calls the left Urand first */
            const double l1 = log(draw(r));
            const double l2 = log(draw(r));
            x = sqrt(-l1 - l2 * cos_val * cos_val);
        }
        V_tilda = x / beta;
        mu_tilda = 2.0 * draw(r) - 1.0;
        accept = draw(r);
    } while (accept > sqrt(P->speed * P->speed + V_tilda * V_tilda - 2.0 * P->speed * V_tilda * mu_tilda) / (P->speed + V_tilda));
    mco_scatter_direction(P->dir, mu_tilda, draw(r), nuclide_dir);
    V_lab[0] = nuclide_dir[0] * V_tilda; V_lab[1] = nuclide_dir[1] * V_tilda; V_lab[2] = nuclide_dir[2] * V_tilda;
    v_lab[0] = P->speed * P->dir[0]; v_lab[1] = P->speed * P->dir[1]; v_lab[2] = P->speed * P->dir[2];
    u[0] = (v_lab[0] + A * V_lab[0]) / (1.0 + A);
    u[1] = (v_lab[1] + A * V_lab[1]) / (1.0 + A);
    u[2] = (v_lab[2] + A * V_lab[2]) / (1.0 + A);
    v_c[0] = v_lab[0] - u[0]; v_c[1] = v_lab[1] - u[1]; v_c[2] = v_lab[2] - u[2];
    speed_c = sqrt(v_c[0] * v_c[0] + v_c[1] * v_c[1] + v_c[2] * v_c[2]);
    dir_c[0] = v_c[0] / speed_c; dir_c[1] = v_c[1] / speed_c; dir_c[2] = v_c[2] / speed_c;
    mco_scatter_direction(dir_c, mu0, draw(r), dir_cNew);
    v_c[0] = speed_c * dir_cNew[0]; v_c[1] = speed_c * dir_cNew[1]; v_c[2] = speed_c * dir_cNew[2];
    v_lab[0] = v_c[0] + u[0]; v_lab[1] = v_c[1] + u[1]; v_lab[2] = v_c[2] + u[2];
    p_set_speed(P, sqrt(v_lab[0] * v_lab[0] + v_lab[1] * v_lab[1] + v_lab[2] * v_lab[2]));
    P->dir[0] = v_lab[0] / P->speed; P->dir[1] = v_lab[1] / P->speed; P->dir[2] = v_lab[2] / P->speed;
}

void mco_scatter_sample(const mcb_problem* p, int nuc, uint64_t* seed, double* io)
{
    particle P;
    rng_ref r = {0, 0, seed};
    const double pos[3] = {0, 0, 0};
    P = p_make(pos, io, io[3], 0.0, 1.0, 0);
    scatter_sample(p->nuclides[nuc].A, &P, &r);
    io[0] = P.dir[0]; io[1] = P.dir[1]; io[2] = P.dir[2]; io[3] = P.E; io[4] = P.speed;
}
double mco_watt_sample(const mcb_problem* p, int nuc, uint64_t* seed, double E)
{
    rng_ref r = {0, 0, seed};
    const mcb_nuclide* N = &p->nuclides[nuc];
    return watt_sample(N->watt_a, N->watt_b, N->watt_g, E, &r);
}

/* ------------------------------------------------------------------------------------------ */
/* Estimator (src/Estimator.cpp)                                                              */
/* ------------------------------------------------------------------------------------------ */
static double kernel_value(const mco_ctx* c, int kernel, const particle* P, double l) /* Estimator.cpp:17-41 */
{
    switch (kernel) {
    case MCB_KERNEL_NEUTRON: return P->w;
    case MCB_KERNEL_TRACK: return P->w * l;
    case MCB_KERNEL_COLLISION: return P->w / macro_(c->p, c->p->cells[P->cell].material, X_T, P->E);
    case MCB_KERNEL_VELOCITY: return P->w * P->speed;
    default: return P->w * l * P->speed;
    }
}
static double score_value(const mco_ctx* c, const mcb_score* S, const particle* P, double l) /* Estimator.cpp:48-124 */
{
    const mcb_problem* p = c->p;
    const int m = p->cells[P->cell].material;
    const double kv = kernel_value(c, S->kernel, P, l);
    if (S->score == MCB_SCORE_FLUX) return kv;
    if (S->score == MCB_SCORE_INVERSE_VELOCITY) return kv / P->speed;
    if (m < 0) return 0.0; /* the reference would dereference a null material here */
    switch (S->score) {
    case MCB_SCORE_ABSORPTION: return macro_(p, m, X_A, P->E) * kv;
    case MCB_SCORE_SCATTER: return macro_(p, m, X_S, P->E) * kv;
    case MCB_SCORE_CAPTURE: return macro_(p, m, X_C, P->E) * kv;
    case MCB_SCORE_FISSION: return macro_(p, m, X_F, P->E) * kv;
    case MCB_SCORE_NU_FISSION: return macro_nuSigmaF(p, m, P->E) * kv;
    case MCB_SCORE_TOTAL: return macro_(p, m, X_T, P->E) * kv;
    case MCB_SCORE_SCATTER_OLD: return macro_(p, m, X_S, P->E_old) * kv;
    case MCB_SCORE_NU_FISSION_OLD: return macro_nuSigmaF(p, m, P->E_old) * kv;
    case MCB_SCORE_NU_FISSION_PROMPT_OLD: return macro_kind(p, m, 2, P->E_old) * kv;
    case MCB_SCORE_NU_FISSION_DELAYED_OLD: return macro_kind(p, m, 3 + S->group, P->E_old) * kv;
    case MCB_SCORE_NU_FISSION_DELAYED_DECAY_OLD: return macro_delayed_decay(p, m, P->E_old, S->group) * kv;
    default: return 0.0;
    }
}
#define MAX_FILTERS 4
#define MAX_SPAN 512
typedef struct idx_l { int n; int idx[MAX_SPAN]; double l[MAX_SPAN]; } idx_l;

/* Filter*::idx_l (Estimator.cpp:133-263) */
static void filter_idx_l(const mco_ctx* c, const mcb_filter* F, const particle* P, double l, idx_l* out)
{
    const double* g = c->p->filter_grid + F->grid_begin;
    const int Nbin = F->grid_n - 1;
    out->n = 0;
    switch (F->type) {
    case MCB_FILTER_SURFACE:
        out->idx[0] = mco_binary_search((double)P->surface_old, g, F->grid_n) + 1; out->l[0] = l; out->n = 1; break;
    case MCB_FILTER_CELL:
        out->idx[0] = mco_binary_search((double)P->cell, g, F->grid_n) + 1; out->l[0] = l; out->n = 1; break;
    case MCB_FILTER_ENERGY:
    case MCB_FILTER_ENERGY_OLD: {
        const int i = mco_binary_search(F->type == MCB_FILTER_ENERGY ? P->E : P->E_old, g, F->grid_n);
        if (i < 0 || i >= Nbin) return;
        out->idx[0] = i; out->l[0] = l; out->n = 1;
        break;
    }
    case MCB_FILTER_TDMC: /* FilterTDMC (Estimator.cpp:247-263): only a particle sitting exactly on its census time scores */
        if (P->t != g[P->tdmc]) return;
        out->idx[0] = P->tdmc; out->l[0] = l; out->n = 1;
        break;
    default: { /* FilterTime (Estimator.cpp:199-246) */
        const int loc1 = mco_binary_search(P->t_old, g, F->grid_n);
        const int loc2 = mco_binary_search(P->t, g, F->grid_n);
        if (loc1 == loc2) {
            if (loc1 >= 0 && loc1 < Nbin) { out->idx[0] = loc1; out->l[0] = l; out->n = 1; }
            return;
        } else {
            const int num_bin = loc2 - loc1 - 1;
            int i;
            if (loc1 >= 0) { out->idx[out->n] = loc1; out->l[out->n] = (g[loc1 + 1] - P->t_old) * P->speed; out->n++; }
            for (i = 1; i <= num_bin && out->n < MAX_SPAN - 1; i++) {
                out->idx[out->n] = loc1 + i; out->l[out->n] = (g[loc1 + i + 1] - g[loc1 + i]) * P->speed; out->n++;
            }
            if (loc2 < Nbin) { out->idx[out->n] = loc2; out->l[out->n] = (P->t - g[loc2]) * P->speed; out->n++; }
        }
    }
    }
}
/* Estimator::score (Estimator.cpp:298-336) */
static void estimator_score_plain(mco_ctx* c, int e, const particle* P, double l_in);
static double chid_sample(const mcb_problem* p, int nuc, int g, rng_ref* r);
/* EstimatorScatter / EstimatorFissionPrompt / EstimatorFissionDelayed ::score (Estimator.cpp:441-482): a copy of the
 * particle undergoes the event, drawn from the same stream as the transport, and is scored with the generic
 * estimator: energy_old = incident energy, energy = outgoing energy */
static void estimator_score(mco_ctx* c, int e, particle* P, double l_in)
{
    const mcb_problem* p = c->p;
    const mcb_estimator* E = &p->estimators[e];
    if (E->simulate == MCB_SIM_NONE) { estimator_score_plain(c, e, P, l_in); return; }
    {
        particle Q = *P;
        rng_ref r;
        const int m = p->cells[P->cell].material;
        int n;
        r.c = c; r.P = P; r.raw = 0;
        if (m < 0) return;
        if (E->simulate == MCB_SIM_SCATTER) {
            n = select_nuclide(p, m, 0, P->E, urand(c, P));
            if (n < 0) return; /* the reference dereferences a null nuclide here */
            scatter_sample(p->nuclides[n].A, &Q, &r);
        } else if (E->simulate == MCB_SIM_FISSION || E->simulate == MCB_SIM_FISSION_PROMPT) {
            const mcb_nuclide* N;
            n = select_nuclide(p, m, E->simulate == MCB_SIM_FISSION ? 1 : 2, P->E, urand(c, P));
            if (n < 0) return;
            N = &p->nuclides[n];
            p_set_energy(&Q, watt_sample(N->watt_a, N->watt_b, N->watt_g, P->E, &r));
        } else {
            const int g = E->simulate - MCB_SIM_FISSION_DELAYED;
            n = select_nuclide(p, m, 3 + g, P->E, urand(c, P));
            if (n < 0) return;
            p_set_energy(&Q, chid_sample(p, n, g, &r));
        }
        Q.rng = P->rng;
        estimator_score_plain(c, e, &Q, l_in);
    }
}
/* ReactionFission::ChiD -> DistributionDelayedNeutron::sample (Distribution.cpp:97-102): tabulated CDF, lin-lin */
static double chid_sample(const mcb_problem* p, int nuc, int g, rng_ref* r)
{
    const mcb_nuclide* N = &p->nuclides[nuc];
    const double* cdf = p->delayed_data + N->chid_cdf_begin[g];
    const double* v = p->delayed_data + N->chid_E_begin;
    const double xi = draw(r);
    const int idx = mco_binary_search(xi, cdf, N->chid_cdf_n[g]);
    return mco_interpolate(xi, cdf[idx], cdf[idx + 1], v[idx], v[idx + 1]);
}
static void estimator_score_plain(mco_ctx* c, int e, const particle* P, double l_in)
{
    const mcb_problem* p = c->p;
    const mcb_estimator* E = &p->estimators[e];
    static idx_l il[MAX_FILTERS];
    int idx[MAX_FILTERS] = {0, 0, 0, 0};
    double factor[MAX_FILTERS + 1];
    int i, s;
    tally* T = c->tallies + E->tally_begin;
    for (i = 0; i < E->n_filters; i++) {
        filter_idx_l(c, &p->filters[E->filter_begin + i], P, l_in, &il[i]);
        if (il[i].n == 0) return;
    }
    /* idx_factor (Estimator.cpp:288-295) */
    factor[E->n_filters] = 1.0;
    for (i = E->n_filters - 1; i >= 0; i--) factor[i] = factor[i + 1] * p->filters[E->filter_begin + i].size;
    for (;;) {
        double l = MAX_float;
        int idx_1D = 0;
        for (i = 0; i < E->n_filters; i++) l = fmin(l, il[i].l[idx[i]]);
        for (i = 0; i < E->n_filters; i++) idx_1D += (int)(il[i].idx[idx[i]] * factor[i + 1]);
        for (s = 0; s < E->n_scores; s++) {
            T[idx_1D].hist += score_value(c, &p->scores[E->score_begin + s], P, l);
            idx_1D += (int)factor[0];
        }
        for (i = 0; i < E->n_filters; i++) {
            il[i].l[idx[i]] -= l;
            if (il[i].l[idx[i]] < EPSILON_float) {
                if (idx[i] == il[i].n - 1) return;
                idx[i]++;
            }
        }
        if (E->n_filters == 0) return;
    }
}
/* estimators attached to a cell (TL or C) or to a surface, in deck order */
static void score_attached(mco_ctx* c, int attach, int id, particle* P, double l)
{
    const mcb_problem* p = c->p;
    int e, i;
    for (e = 0; e < p->n_estimators; e++) {
        const mcb_estimator* E = &p->estimators[e];
        const mcb_filter* F0;
        if (E->attach != attach) continue;
        F0 = &p->filters[E->filter_begin];
        if (F0->type == MCB_FILTER_TDMC) F0++; /* the geometry filter follows the TDMC filter (setup.cpp:703-741) */
        for (i = 0; i < F0->grid_n; i++) {
            if ((int)p->filter_grid[F0->grid_begin + i] == id) { estimator_score(c, e, P, l); break; }
        }
    }
}

/* ShannonEntropy::score (Entropy.cpp:27-35) */
static void entropy_score(mco_ctx* c, const double* pos, int N)
{
    const mcb_problem* p = c->p;
    const double* gx = p->entropy_grid;
    const double* gy = gx + p->entropy_n[0];
    const double* gz = gy + p->entropy_n[1];
    const int Iy = p->entropy_n[1] - 1, Iz = p->entropy_n[2] - 1;
    const int ix = mco_binary_search(pos[0], gx, p->entropy_n[0]);
    const int iy = mco_binary_search(pos[1], gy, p->entropy_n[1]);
    const int iz = mco_binary_search(pos[2], gz, p->entropy_n[2]);
    const int idx = ix * (Iz * Iy) + iy * Iz + iz;
    if (idx < 0 || idx >= c->ent_I) return; /* the reference indexes out of bounds here (quirk 14) */
    c->ent_p[idx] += N;
}
/* ShannonEntropy::calculate_H (Entropy.cpp:43-62) */
static double entropy_H(mco_ctx* c)
{
    double sum = 0.0;
    int i;
    if (!c->p->entropy_on) return 0.0;
    for (i = 0; i < c->ent_I; i++) sum += c->ent_p[i];
    if (sum == 0.0) return sum;
    for (i = 0; i < c->ent_I; i++) c->ent_p[i] /= sum;
    sum = 0;
    for (i = 0; i < c->ent_I; i++) { if (c->ent_p[i] != 0) sum -= c->ent_p[i] * log2(c->ent_p[i]); }
    memset(c->ent_p, 0, sizeof(double) * (size_t)c->ent_I);
    return sum;
}

/* ------------------------------------------------------------------------------------------ */
/* banks                                                                                      */
/* ------------------------------------------------------------------------------------------ */
static void push(particle** b, size_t* n, size_t* cap, const particle* P)
{
    if (*n == *cap) {
        *cap = *cap ? *cap * 2 : 1024;
        *b = (particle*)realloc(*b, *cap * sizeof(particle));
        if (!*b) { fprintf(stderr, "mc_oracle: out of memory\n"); exit(1); }
    }
    (*b)[(*n)++] = *P;
}
/* stream of a secondary (MCO_RNG_HISTORY): jump (j+1)*2^40 draws from the parent's state at the push */
static void give_child_stream(mco_ctx* c, const particle* parent, particle* child)
{
    if (c->rng_mode == MCO_RNG_HISTORY) child->rng = mco_lcg_skip(parent->rng, ((uint64_t)(c->child_counter + 1)) << 40);
    c->child_counter++;
}

/* ------------------------------------------------------------------------------------------ */
/* population control (src/simulator/population_control.cpp)                                  */
/* ------------------------------------------------------------------------------------------ */
static void weight_roulette(mco_ctx* c, particle* P) /* :9-15 */
{
    if (P->w < c->p->wr) {
        if (urand(c, P) < P->w / c->p->ws) P->w = c->p->ws;
        else p_kill(P);
    }
}
static void cell_importance(mco_ctx* c, particle* P) /* :21-49 */
{
    const double Iold = c->p->cells[P->cell_old].importance;
    const double Inew = c->p->cells[P->cell].importance;
    double rat;
    if (Inew == Iold) return;
    rat = Inew / Iold;
    if (rat < 1.0) {
        if (urand(c, P) < rat) P->w = P->w / rat;
        else p_kill(P);
    } else {
        const int n = (int)floor(rat + urand(c, P));
        int i;
        P->w = P->w / (double)n;
        for (i = 0; i < n - 1; i++) {
            particle Q = *P;
            give_child_stream(c, P, &Q);
            push(&c->Pbank, &c->Pn, &c->Pcap, &Q);
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* transport (src/simulator/general.cpp, ksearch.cpp, fixed_source.cpp)                       */
/* ------------------------------------------------------------------------------------------ */
static void move_particle(mco_ctx* c, particle* P, double l) /* general.cpp:73-83 */
{
    const mcb_problem* p = c->p;
    const int m = p->cells[P->cell].material;
    p_move(P, l);
    c->Ntrack++;
    if (p->ksearch && m >= 0) c->k_TL += macro_nuSigmaF(p, m, P->E) * P->w * l; /* Estimator.cpp:509-512 */
    if (c->tally_on) score_attached(c, MCB_ATTACH_CELL_TL, P->cell, P, l);
}
static int surface_hit(mco_ctx* c, particle* P, int S) /* general.cpp:89-115 */
{
    const mcb_problem* p = c->p;
    P->surface_old = S;
    if (p->surfaces[S].bc == 0) {
        int cn;
        p_move(P, EPSILON_float);
        cn = mco_search_cell(p, P->pos);
        if (cn < 0) {
            printf("[WARNING] A particle is lost:\n( x, y, z )  (%g, %g, %g )\n", P->pos[0], P->pos[1], P->pos[2]);
            return -1;
        }
        p_set_cell(P, cn);
    } else if (p->surfaces[S].bc == -1) {
        p_kill(P);
        p_set_cell(P, P->cell);
    } else {
        mco_surface_reflect(p, S, P->dir);
        p_move(P, EPSILON_float);
        p_set_cell(P, P->cell);
    }
    if (c->tally_on) score_attached(c, MCB_ATTACH_SURFACE, S, P, 0.0);
    cell_importance(c, P);
    return 0;
}
/* fission-site particle: Particle(P.pos(), isotropic.sample(), Chi(E), P.time(), 1.0, ..) — g++ evaluates the
 * arguments right to left: Watt energy first, then the direction (ksearch.cpp:41-46, fixed_source.cpp:16-21) */
/* MCO_RNG_HISTORY: the b-th neutron banked by a collision is sampled from its own stream, the parent's state at the
 * banking point jumped (b+1)*2^40 draws ahead; the parent's stream does not advance.  (With one global stream, as in
 * the reference, the neutrons are simply sampled from it in order.)  This lets the GPU sample all fission sites of
 * a generation in one dense pass instead of inside the divergent collision branch. */
static particle fission_neutron(mco_ctx* c, particle* P, int nuc, int b)
{
    const mcb_nuclide* N = &c->p->nuclides[nuc];
    double dir[3];
    if (c->rng_mode == MCO_RNG_HISTORY) {
        uint64_t s = mco_lcg_skip(P->rng, ((uint64_t)(b + 1)) << 40);
        rng_ref r = {0, 0, &s};
        particle Q;
        const double E = watt_sample(N->watt_a, N->watt_b, N->watt_g, P->E, &r);
        isotropic_direction(&r, dir);
        Q = p_make(P->pos, dir, E, P->t, 1.0, P->cell);
        Q.tdmc = P->tdmc;
        Q.rng = s;
        c->n_draws += 0;
        return Q;
    } else {
        rng_ref r = {c, P, 0};
        const double E = watt_sample(N->watt_a, N->watt_b, N->watt_g, P->E, &r);
        isotropic_direction(&r, dir);
        {
            particle Q = p_make(P->pos, dir, E, P->t, 1.0, P->cell);
            Q.tdmc = P->tdmc;
            return Q;
        }
    }
}
static void weight_roulette(mco_ctx* c, particle* P);
/* Simulator::forced_decay (time_dependent.cpp:16-45): a delayed neutron of the fission at P, forced to appear inside
 * [initial, initial + interval) with the weight of the expected emission there.  Draws, in order: emission time, precursor
 * group, ChiD, direction (2).  MCO_RNG_HISTORY: the neutron is child number `q` of this collision and everything about
 * it, its roulette included, is drawn from its own stream. */
static particle forced_decay(mco_ctx* c, particle* P, int nuc, double initial, double interval, int p_tdmc, int q)
{
    const mcb_nuclide* N = &c->p->nuclides[nuc];
    uint64_t s = 0;
    rng_ref r = {c, P, 0};
    double prob[6], p_weight = 0.0, total = 0.0, sum = 0.0, p_time, xi, E, dir[3];
    int k, cg = 0;
    particle Q;
    if (c->rng_mode == MCO_RNG_HISTORY) { s = mco_lcg_skip(P->rng, ((uint64_t)(q + 1)) << 40); r.c = 0; r.P = 0; r.raw = &s; }
    p_time = initial + draw(&r) * interval;
    for (k = 0; k < 6; k++) {
        prob[k] = N->fraction[k] * N->lambda[k] * exp(-N->lambda[k] * (p_time - P->t)); /* f_lambda(k) = fraction * lambda */
        p_weight += prob[k];
    }
    p_weight *= interval;
    for (k = 0; k < 6; k++) total += prob[k]; /* std::accumulate */
    xi = draw(&r) * total;
    for (k = 0; k < 6; k++) { sum += prob[k]; if (sum > xi) { cg = k; break; } }
    E = chid_sample(c->p, nuc, cg, &r);
    isotropic_direction(&r, dir);
    Q = p_make(P->pos, dir, E, p_time, p_weight, P->cell);
    Q.tdmc = p_tdmc;
    if (c->rng_mode == MCO_RNG_HISTORY) Q.rng = s;
    return Q;
}
static void collision(mco_ctx* c, particle* P) /* general.cpp:121-163 */
{
    const mcb_problem* p = c->p;
    const int m = p->cells[P->cell].material;
    double bank_nu, implicit;
    int N_fission, N_scatter, i;
    if (m < 0) { p_kill(P); return; }
    c->Ncollision++;
    if (c->tally_on) score_attached(c, MCB_ATTACH_CELL_C, P->cell, P, 0);
    {   /* floor( w/k * nuSigmaF / SigmaT + Urand() ) (general.cpp:135-136); one draw */
        const double a = P->w / c->k * macro_nuSigmaF(p, m, P->E) / macro_(p, m, X_T, P->E);
        bank_nu = floor(a + urand(c, P));
    }
    N_fission = select_nuclide(p, m, 1, P->E, urand(c, P));
    if (N_fission < 0) {
        /* oracle patch B: the reference dereferences a null nuclide here */
    } else if (p->ksearch) {
        /* implicit_fission_ksearch (ksearch.cpp:20-47) */
        if (urand(c, P) > micro(p, N_fission, X_BETA, P->E)) {
            /* prompt */
        } else {
            (void)urand(c, P); /* precursor group pick; the result is never used */
        }
        for (i = 0; i < bank_nu; i++) {
            particle Q = fission_neutron(c, P, N_fission, i);
            push(&c->Fbank, &c->Fn, &c->Fcap, &Q);
        }
        /* EstimatorK::estimate_C (Estimator.cpp:503-507) */
        c->k_C += macro_nuSigmaF(p, m, P->E) * P->w / macro_(p, m, X_T, P->E);
        if (bank_nu > 0 && p->entropy_on) entropy_score(c, P->pos, (int)bank_nu);
    } else {
        /* implicit_fission_fixed_source, prompt branch (fixed_source.cpp:12-22) */
        if (urand(c, P) > micro(p, N_fission, X_BETA, P->E)) {
            for (i = 0; i < bank_nu; i++) {
                particle Q = fission_neutron(c, P, N_fission, i); /* HISTORY mode: Q continues on its own stream */
                push(&c->Pbank, &c->Pn, &c->Pcap, &Q);
            }
        } else {
            if (p->tdmc_on) {
                /* combined and forced decay (fixed_source.cpp:25-40): per fission neutron one delayed neutron in the rest of
                 * the current interval and one in every later interval, each through the roulette on its own */
                int q = 0, j;
                for (i = 0; i < bank_nu; i++) {
                    particle Q = forced_decay(c, P, N_fission, P->t, p->tdmc_time[P->tdmc] - P->t, P->tdmc, q++);
                    weight_roulette(c, &Q);
                    if (Q.alive) push(&c->Pbank, &c->Pn, &c->Pcap, &Q);
                    for (j = P->tdmc; j < p->n_tdmc - 1; j++) {
                        Q = forced_decay(c, P, N_fission, p->tdmc_time[j], p->tdmc_interval[j + 1], j + 1, q++);
                        weight_roulette(c, &Q);
                        if (Q.alive) push(&c->Pbank, &c->Pn, &c->Pcap, &Q);
                    }
                }
            } else {
            /* delayed, non-TDMC branch (fixed_source.cpp:41-63): one draw for the precursor group, then per
             * neutron one draw for ChiD and one for the emission time; the loop over tdmc_time (:54-61) is empty
             * without a <tdmc> block, so NO particle is banked — delayed neutrons are dropped (reference bug 20) */
            (void)urand(c, P);
            for (i = 0; i < bank_nu; i++) { (void)urand(c, P); (void)urand(c, P); }
            }
        }
    }
    /* implicit absorption (general.cpp:154-156) */
    implicit = macro_(p, m, X_C, P->E) + macro_(p, m, X_F, P->E);
    P->w = P->w * (macro_(p, m, X_T, P->E) - implicit) / macro_(p, m, X_T, P->E);
    N_scatter = select_nuclide(p, m, 0, P->E, urand(c, P));
    if (N_scatter < 0) return;
    {
        rng_ref r = {c, P, 0};
        scatter_sample(p->nuclides[N_scatter].A, P, &r);
    }
}
/* particle_comb (population_control.cpp:55-84), called after every random walk (handler.cpp:27-28).  The comb's one
 * draw comes from the global stream (MCO_RNG_GLOBAL) or from the stream of the particle whose walk just ended
 * (MCO_RNG_HISTORY).  Reproduced as it is: `if`, not `while`, so a heavy particle takes one tooth; the new bank starts as
 * `teeth` copies of Pbank[0], and teeth the loop does not reach keep that copy with its ORIGINAL weight; a tooth past
 * the end (the reference writes out of bounds there) is dropped.  MCO_RNG_HISTORY: a combed particle keeps the stream
 * of the particle it copies; the leftover copies of Pbank[0] get streams of their own, (q + 1) * 2^40 draws on. */
static void particle_comb(mco_ctx* c, particle* done)
{
    const mcb_problem* p = c->p;
    const size_t teeth = (size_t)p->comb_teeth;
    double W = 0.0, w_avg, tooth, sum = 0.0;
    particle tmp[256];
    size_t i, j = 0, q;
    if (!p->comb_on || c->Pn < (size_t)p->comb_bank_max) return;
    for (i = 0; i < c->Pn; i++) W += c->Pbank[i].w;
    w_avg = W / (double)p->comb_teeth;
    tooth = urand(c, done) * w_avg;
    for (q = 0; q < teeth; q++) tmp[q] = c->Pbank[0];
    for (i = 0; i < c->Pn; i++) {
        sum += c->Pbank[i].w;
        if (sum > tooth) {
            if (j < teeth) { tmp[j] = c->Pbank[i]; tmp[j].w = w_avg; }
            tooth += w_avg; j++;
        }
    }
    if (c->rng_mode == MCO_RNG_HISTORY)
        for (q = j; q < teeth; q++) tmp[q].rng = mco_lcg_skip(c->Pbank[0].rng, ((uint64_t)(q + 1)) << 40);
    c->Pn = 0;
    for (q = 0; q < teeth; q++) push(&c->Pbank, &c->Pn, &c->Pcap, &tmp[q]);
}
static int random_walk(mco_ctx* c, particle* P) /* general.cpp:177-211 */
{
    const mcb_problem* p = c->p;
    while (P->alive) {
        double dsurf, dcol;
        const int S = mco_surface_intersect(p, P->cell, P->pos, P->dir, &dsurf);
        const int m = p->cells[P->cell].material;
        c->child_counter = 0;
        if (m >= 0) dcol = -log(urand(c, P)) / macro_(p, m, X_T, P->E); /* general.cpp:40-48, Algorithm.cpp:123-126 */
        else dcol = 0.9 * MAX_float;
        if (p->tdmc_on) { /* general.cpp:187-195; Simulator::time_hit (time_dependent.cpp:51-55); no roulette on this path */
            const double dbound = (p->tdmc_time[P->tdmc] - P->t) * P->speed;
            if ((dsurf < dcol ? dsurf : dcol) > dbound) {
                move_particle(c, P, dbound);
                P->tdmc++;
                if (P->tdmc == p->n_tdmc) p_kill(P);
                continue;
            }
        }
        if (dcol > dsurf) {
            move_particle(c, P, dsurf);
            if (S < 0) { p_kill(P); } /* cannot happen with Sigma_t > 0; the reference would dereference null */
            else if (surface_hit(c, P, S) != 0) return -1;
        } else {
            move_particle(c, P, dcol);
            collision(c, P);
        }
        weight_roulette(c, P);
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* source sampling (src/Source.cpp)                                                           */
/* ------------------------------------------------------------------------------------------ */
static size_t pick_source(mco_ctx* c, particle* stream, size_t n)
{
    const double xi = urand(c, stream);
    if (c->pick_mode == MCO_PICK_CDF) { /* Source.cpp:42-46; the reference would index out of range if xi > p[n] */
        const int j = mco_binary_search(xi, c->cdf, (int)c->cdf_n);
        return (size_t)(j < 0 ? 0 : (j >= (int)n ? (int)n - 1 : j));
    }
    {
        size_t j = (size_t)(xi * (double)n);
        return j < n ? j : n - 1;
    }
}
static void bank_set_up(mco_ctx* c, size_t n) /* SourceBank::set_up (Source.cpp:65-72) */
{
    size_t i;
    if (c->pick_mode != MCO_PICK_CDF) return;
    c->cdf = (double*)realloc(c->cdf, (n + 1) * sizeof(double));
    c->cdf_n = n + 1;
    c->cdf[0] = 0.0;
    {
        const double dp = 1.0 / n;
        for (i = 1; i < n + 1; i++) c->cdf[i] = c->cdf[i - 1] + dp;
    }
}
/* SourcePoint::get_source (Source.cpp:20-24): energy is sampled before the direction (right-to-left arguments) */
static particle deck_source(mco_ctx* c, const mcb_source* S, particle* stream)
{
    rng_ref r = {c, stream, 0};
    double dir[3];
    const double E = dist1_sample(&S->energy, &r);
    if (S->dir_kind == MCB_DIR_DELTA) { dir[0] = S->dir[0]; dir[1] = S->dir[1]; dir[2] = S->dir[2]; }
    else if (S->dir_kind == MCB_DIR_ISOTROPIC) isotropic_direction(&r, dir);
    else { /* DistributionIndepndentXYZ::sample: Point(x->sample(), y->sample(), z->sample()), right to left */
        dir[2] = dist1_sample(&S->dir_xyz[2], &r);
        dir[1] = dist1_sample(&S->dir_xyz[1], &r);
        dir[0] = dist1_sample(&S->dir_xyz[0], &r);
    }
    if (S->kind == MCB_SRC_DISK_Z) { /* no reference counterpart (setup.cpp:1051-1063 rejects the element): mcb200.h */
        double pos[3];
        const double rho = S->radius * sqrt(draw(&r));
        const double phi = 2.0 * PI_ * draw(&r);
        int cell;
        pos[0] = S->pos[0] + rho * cos(phi); pos[1] = S->pos[1] + rho * sin(phi); pos[2] = S->pos[2];
        cell = mco_search_cell(c->p, pos);
        if (cell < 0) {
            printf("[WARNING] A particle is lost:\n( x, y, z )  (%g, %g, %g )\n", pos[0], pos[1], pos[2]);
            c->src_lost = 1; cell = S->cell;
        }
        return p_make(pos, dir, E, 0.0, 1.0, cell);
    }
    return p_make(S->pos, dir, E, 0.0, 1.0, S->cell);
}

/* ------------------------------------------------------------------------------------------ */
/* driver (src/simulator/handler.cpp)                                                         */
/* ------------------------------------------------------------------------------------------ */
mco_ctx* mco_create(const mcb_problem* p, int rng_mode, int pick_mode)
{
    mco_ctx* c = (mco_ctx*)calloc(1, sizeof(mco_ctx));
    c->p = p; c->rng_mode = rng_mode; c->pick_mode = pick_mode;
    c->seed = p->seed ? p->seed : 1ULL;
    c->k = 1.0;
    c->shard_begin = 0; c->shard_count = p->n_sample;
    c->S_is_deck = 1;
    c->tallies = (tally*)calloc((size_t)(p->n_tallies > 0 ? p->n_tallies : 1), sizeof(tally));
    if (p->entropy_on) {
        c->ent_I = (p->entropy_n[0] - 1) * (p->entropy_n[1] - 1) * (p->entropy_n[2] - 1);
        c->ent_p = (double*)calloc((size_t)c->ent_I, sizeof(double));
    }
    return c;
}
void mco_destroy(mco_ctx* c)
{
    if (!c) return;
    free(c->Pbank); free(c->Fbank); free(c->Sbank); free(c->cdf); free(c->tallies); free(c->ent_p);
    free(c->hist_kC); free(c->hist_kTL); free(c);
}
void mco_set_shard(mco_ctx* c, uint64_t begin, uint64_t count) { c->shard_begin = begin; c->shard_count = count; }

int mco_transport_cycle(mco_ctx* c)
{
    const mcb_problem* p = c->p;
    uint64_t h;
    size_t nsrc;
    if (c->icycle == p->n_passive) c->tally_on = 1;       /* handler.cpp:15 */
    /* Sbank = Fbank; Sbank.set_up(); Fbank.reset() (handler.cpp:16) — done by mco_set_source_bank/close */
    nsrc = c->S_is_deck ? (size_t)p->n_sources : c->Sn;
    if (nsrc == 0) { fprintf(stderr, "[ERROR] Source bank is empty...\n"); return -1; }
    bank_set_up(c, nsrc);
    c->Fn = 0;
    c->cyc_tracks0 = c->Ntrack; c->cyc_coll0 = c->Ncollision; c->cyc_hist = 0;
    if (p->ksearch && c->hist_n != (size_t)c->shard_count) {
        c->hist_n = (size_t)c->shard_count;
        c->hist_kC = (double*)realloc(c->hist_kC, (c->hist_n ? c->hist_n : 1) * sizeof(double));
        c->hist_kTL = (double*)realloc(c->hist_kTL, (c->hist_n ? c->hist_n : 1) * sizeof(double));
    }
    for (h = c->shard_begin; h < c->shard_begin + c->shard_count; h++) {
        particle src, stream;
        size_t j;
        memset(&stream, 0, sizeof(stream));
        if (c->rng_mode == MCO_RNG_HISTORY) stream.rng = mco_lcg_skip(c->seed, (c->icycle * p->n_sample + h) * RN_STRIDE);
        j = pick_source(c, &stream, nsrc);                /* handler.cpp:20, Source.cpp:42-46 */
        if (c->S_is_deck) { src = deck_source(c, &p->sources[j], &stream); if (c->src_lost) return -1; }
        else src = c->Sbank[j];
        src.rng = stream.rng;
        c->child_counter = 0;
        push(&c->Pbank, &c->Pn, &c->Pcap, &src);
        while (c->Pn) {                                    /* handler.cpp:22-29 */
            particle P = c->Pbank[--c->Pn];
            if (random_walk(c, &P) != 0) return -1;
            particle_comb(c, &P);                          /* handler.cpp:27-28 */
        }
        if (c->tally_on) {                                 /* Estimator::end_history (Estimator.cpp:339-346) */
            int64_t t;
            for (t = 0; t < p->n_tallies; t++) {
                tally* T = &c->tallies[t];
                T->sum += T->hist; T->squared += T->hist * T->hist; T->hist = 0.0;
            }
        }
        if (p->ksearch) {                                  /* EstimatorK::end_history (Estimator.cpp:514-525) */
            c->H_sum += entropy_H(c);
            c->hist_kC[h - c->shard_begin] = c->k_C; c->hist_kTL[h - c->shard_begin] = c->k_TL;
            c->k_sum_C += c->k_C; c->k_sum_TL += c->k_TL;
            c->k_sq_C += c->k_C * c->k_C; c->k_sq_TL += c->k_TL * c->k_TL;
            c->k_C = 0; c->k_TL = 0;
        }
        c->cyc_hist++;
    }
    return 0;
}
void mco_get_partials(const mco_ctx* c, double* s, uint64_t* n)
{
    s[0] = c->k_sum_C; s[1] = c->k_sum_TL; s[2] = c->k_sq_C; s[3] = c->k_sq_TL; s[4] = c->H_sum;
    n[0] = c->Fn; n[1] = c->Ntrack - c->cyc_tracks0; n[2] = c->Ncollision - c->cyc_coll0; n[3] = c->cyc_hist;
}
int64_t mco_get_history_k(const mco_ctx* c, double* kC, double* kTL)
{
    if (!c->hist_kC) return 0;
    memcpy(kC, c->hist_kC, c->hist_n * sizeof(double)); memcpy(kTL, c->hist_kTL, c->hist_n * sizeof(double));
    return (int64_t)c->hist_n;
}
int64_t mco_bank_size(const mco_ctx* c) { return (int64_t)c->Fn; }
void mco_get_bank(const mco_ctx* c, double* s, int32_t* cells)
{
    size_t i;
    for (i = 0; i < c->Fn; i++) {
        const particle* P = &c->Fbank[i];
        s[8 * i + 0] = P->pos[0]; s[8 * i + 1] = P->pos[1]; s[8 * i + 2] = P->pos[2];
        s[8 * i + 3] = P->dir[0]; s[8 * i + 4] = P->dir[1]; s[8 * i + 5] = P->dir[2];
        s[8 * i + 6] = P->E; s[8 * i + 7] = P->t;
        cells[i] = P->cell;
    }
}
void mco_set_source_bank(mco_ctx* c, const double* s, const int32_t* cells, int64_t n)
{
    int64_t i;
    c->Sbank = (particle*)realloc(c->Sbank, (size_t)(n > 0 ? n : 1) * sizeof(particle));
    for (i = 0; i < n; i++) c->Sbank[i] = p_make(s + 8 * i, s + 8 * i + 3, s[8 * i + 6], s[8 * i + 7], 1.0, cells[i]);
    c->Sn = (size_t)n;
    c->S_is_deck = 0;
}
void mco_get_tally_partials(const mco_ctx* c, double* sum, double* squared)
{
    int64_t t;
    for (t = 0; t < c->p->n_tallies; t++) { sum[t] = c->tallies[t].sum; squared[t] = c->tallies[t].squared; }
}
void mco_close_cycle(mco_ctx* c, const double* s, const uint64_t* n, const double* tsum, const double* tsq, mco_cycle_result* out)
{
    const mcb_problem* p = c->p;
    const double Ns = (double)p->n_sample;
    mco_cycle_result r;
    memset(&r, 0, sizeof(r));
    if (c->tally_on) {                                     /* Estimator::end_cycle (Estimator.cpp:347-360) */
        int64_t t;
        for (t = 0; t < p->n_tallies; t++) {
            tally* T = &c->tallies[t];
            const double mean = tsum[t] / Ns;
            const double uncer_squared = (tsq[t] / Ns - mean * mean) / (Ns - 1.0);
            T->mean += mean; T->uncer += uncer_squared; T->sum = 0.0; T->squared = 0.0;
        }
    }
    if (p->ksearch) {                                      /* EstimatorK::report_cycle (Estimator.cpp:526-561) */
        const double mean_C = s[0] / Ns, mean_TL = s[1] / Ns;
        const double mean = (mean_C + mean_TL) / 2;
        r.H = s[4] / Ns;
        r.k_cycle = mean;
        c->k = mean;
        if (c->tally_on) {
            const double us_C = (s[2] / Ns - mean_C * mean_C) / (Ns - 1.0);
            const double us_TL = (s[3] / Ns - mean_TL * mean_TL) / (Ns - 1.0);
            c->Navg++;
            c->mean_accumulator += mean;
            c->uncer_sq_accumulator += us_C + us_TL;
            r.k_avg = c->mean_accumulator / c->Navg;
            r.k_uncer = sqrt(c->uncer_sq_accumulator) / c->Navg / 2;
        }
        c->k_sum_C = c->k_sum_TL = c->k_sq_C = c->k_sq_TL = c->H_sum = 0.0;
    }
    r.k_sum_C = s[0]; r.k_sum_TL = s[1]; r.k_sq_C = s[2]; r.k_sq_TL = s[3]; r.H_sum = s[4];
    r.n_sites = n[0]; r.n_tracks = n[1]; r.n_collisions = n[2]; r.n_histories = n[3]; r.n_draws = c->n_draws;
    c->icycle++;
    if (out) *out = r;
}
int mco_run_cycle(mco_ctx* c, mco_cycle_result* out)
{
    double s[5];
    uint64_t n[4];
    double *tsum, *tsq;
    const int64_t nt = c->p->n_tallies > 0 ? c->p->n_tallies : 1;
    if (mco_transport_cycle(c) != 0) return -1;
    mco_get_partials(c, s, n);
    tsum = (double*)malloc(sizeof(double) * (size_t)nt); tsq = (double*)malloc(sizeof(double) * (size_t)nt);
    mco_get_tally_partials(c, tsum, tsq);
    if (c->p->ksearch) { /* Sbank = Fbank (handler.cpp:16 of the next cycle) */
        particle* tmp = c->Sbank; c->Sbank = c->Fbank; c->Fbank = tmp;
        c->Sn = c->Fn; c->Fn = 0; c->Fcap = 0;
        if (c->Fbank) { free(c->Fbank); c->Fbank = 0; }
        c->S_is_deck = 0;
        n[0] = c->Sn;
    }
    mco_close_cycle(c, s, n, tsum, tsq, out);
    free(tsum); free(tsq);
    return 0;
}
void mco_end_simulation(mco_ctx* c) /* Estimator::end_simulation (Estimator.cpp:361-367) */
{
    const double Nactive = (double)(c->p->n_cycle - c->p->n_passive);
    int64_t t;
    for (t = 0; t < c->p->n_tallies; t++) {
        c->tallies[t].mean = c->tallies[t].mean / Nactive;
        c->tallies[t].uncer = sqrt(c->tallies[t].uncer) / Nactive;
    }
}
void mco_get_tallies(const mco_ctx* c, double* mean, double* uncer)
{
    int64_t t;
    for (t = 0; t < c->p->n_tallies; t++) { mean[t] = c->tallies[t].mean; uncer[t] = c->tallies[t].uncer; }
}
double mco_get_k(const mco_ctx* c) { return c->k; }
uint64_t mco_get_seed(const mco_ctx* c) { return c->seed; }
