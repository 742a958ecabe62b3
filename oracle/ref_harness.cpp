// TEST INFRASTRUCTURE ONLY (oracle/): function-level access to the compiled
// reference (ilhamv/MC-old) for differential tests.  Linked by
// oracle/build_ref.py against the reference's own objects (everything except
// Main.cpp and src/Random.cpp) into
// oracle/_ref/libref_harness.so.  Our code here only *calls* the reference.
//
// Urand() is re-provided here (instead of the reference's src/Random.cpp) so
// that tests can (a) set/read the LCG seed and (b) inject explicit xi values:
// same generator (src/Random.cpp:92-105,121-126 — 63-bit LCG, multiplier
// 3512401965023503517, mask 2^63-1, xi = seed * 2^-63), plus a FIFO that takes
// precedence while non-empty.  The unmodified src/Random.cpp is what
// oracle/_ref/MC_ref itself links.
#include <cstdio>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <vector>

#include "simulator.h"
#include "Algorithm.h"
#include "Random.h"

static unsigned long long g_seed = 1ULL;
static std::deque<double> g_inject;
static unsigned long long g_draws = 0;

double Urand( void )
{
    g_draws++;
    if( !g_inject.empty() ){ const double x = g_inject.front(); g_inject.pop_front(); return x; }
    g_seed = ( 3512401965023503517ULL * g_seed ) & ( (~0ULL) >> 1 );
    return (double)( g_seed * ( 1.0 / (double)( 1ULL << 63 ) ) );
}
void RN_init_particle( unsigned long long* nps ) { (void)nps; }

struct Harness { std::unique_ptr<Simulator> sim; };

static int nuclide_index( Simulator& S, const std::shared_ptr<Nuclide>& N )
{
    if( !N ) return -1;
    for( size_t i = 0; i < S.Nuclides.size(); i++ ){ if( S.Nuclides[i] == N ) return (int)i; }
    return -2;
}

extern "C" {

void refh_set_seed( unsigned long long s ) { g_seed = s; }
unsigned long long refh_get_seed( void ) { return g_seed; }
unsigned long long refh_draws( void ) { return g_draws; }
double refh_urand( void ) { return Urand(); }
void refh_inject( const double* xi, int n ) { for( int i = 0; i < n; i++ ) g_inject.push_back( xi[i] ); }
void refh_clear_inject( void ) { g_inject.clear(); }
int refh_inject_left( void ) { return (int)g_inject.size(); }

// deck_dir must end with '/', CWD must contain ./xs_library (src/simulator/setup.cpp:326)
void* refh_create( const char* deck_dir )
{
    Harness* h = new Harness;
    h->sim.reset( new Simulator( std::string( deck_dir ) ) );
    return h;
}
void refh_destroy( void* p ) { delete (Harness*)p; }

int refh_counts( void* p, int* out )
{
    Simulator& S = *((Harness*)p)->sim;
    out[0] = (int)S.Nuclides.size(); out[1] = (int)S.Materials.size();
    out[2] = (int)S.Surfaces.size(); out[3] = (int)S.Cells.size();
    out[4] = (int)S.Estimators.size();
    return 0;
}

// out = { SigmaT, SigmaS, SigmaC, SigmaF, nuSigmaF, SigmaA }   (src/Material.cpp:18-65)
void refh_sigma( void* p, int mat, const double* E, int n, double* out )
{
    Simulator& S = *((Harness*)p)->sim;
    Material& M = *S.Materials[mat];
    for( int i = 0; i < n; i++ ){
        out[6*i+0] = M.SigmaT( E[i] ); out[6*i+1] = M.SigmaS( E[i] ); out[6*i+2] = M.SigmaC( E[i] );
        out[6*i+3] = M.SigmaF( E[i] ); out[6*i+4] = M.nuSigmaF( E[i] ); out[6*i+5] = M.SigmaA( E[i] );
    }
}
// per-nuclide micro data: out = { sigmaT, sigmaS, sigmaC, sigmaF, nusigmaF, beta }   (src/Nuclide.cpp:36-77)
void refh_micro( void* p, int nuc, const double* E, int n, double* out )
{
    Simulator& S = *((Harness*)p)->sim;
    Nuclide& N = *S.Nuclides[nuc];
    for( int i = 0; i < n; i++ ){
        out[6*i+0] = N.sigmaT( E[i] ); out[6*i+1] = N.sigmaS( E[i] ); out[6*i+2] = N.sigmaC( E[i] );
        out[6*i+3] = N.sigmaF( E[i] ); out[6*i+4] = N.nusigmaF( E[i] ); out[6*i+5] = N.beta( E[i] );
    }
}
// kind 0: nuclide_scatter, 1: nuclide_nufission (src/Material.cpp:106-125); xi injected; returns index in Nuclides
void refh_select( void* p, int mat, int kind, const double* E, const double* xi, int n, int* out )
{
    Simulator& S = *((Harness*)p)->sim;
    Material& M = *S.Materials[mat];
    for( int i = 0; i < n; i++ ){
        g_inject.clear(); g_inject.push_back( xi[i] );
        std::shared_ptr<Nuclide> N = ( kind == 0 ) ? M.nuclide_scatter( E[i] ) : M.nuclide_nufission( E[i] );
        out[i] = nuclide_index( S, N );
    }
    g_inject.clear();
}

int refh_binary_search( double x, const double* v, int n )
{
    std::vector<double> vec( v, v + n );
    return binary_search( x, vec );
}
double refh_interpolate( double x, double x1, double x2, double y1, double y2 ) { return interpolate( x, x1, x2, y1, y2 ); }
double refh_geometry_quad( double a, double b, double c ) { return geometry_quad( a, b, c ); }
// consumes one Urand (azimuth), src/Algorithm.cpp:67-101
void refh_scatter_direction( const double* d, double mu0, double* out )
{
    Point r = scatter_direction( Point( d[0], d[1], d[2] ), mu0 );
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}

// surface: eval at pos, distance for (pos,dir), reflect dir (src/Geometry.cpp)
void refh_surface( void* p, int s, const double* pos, const double* dir, double* out )
{
    Simulator& S = *((Harness*)p)->sim;
    Particle P( Point( pos[0], pos[1], pos[2] ), Point( dir[0], dir[1], dir[2] ), 1.0e6, 0.0, 1.0, 0, S.Cells[0] );
    out[0] = S.Surfaces[s]->eval( P.pos() );
    out[1] = S.Surfaces[s]->distance( P );
    S.Surfaces[s]->reflect( P );
    out[2] = P.dir().x; out[3] = P.dir().y; out[4] = P.dir().z;
    out[5] = (double)S.Surfaces[s]->bc();
}

// free-gas elastic scatter off nuclide nuc (src/Reaction.cpp:27-118); randoms come from Urand (inject or LCG)
// io = { dir.x, dir.y, dir.z, E } in, overwritten with the outgoing state + { speed }
void refh_scatter_sample( void* p, int nuc, double* io )
{
    Simulator& S = *((Harness*)p)->sim;
    Particle P( Point( 0, 0, 0 ), Point( io[0], io[1], io[2] ), io[3], 0.0, 1.0, 0, S.Cells[0] );
    S.Nuclides[nuc]->scatter()->sample( P );
    io[0] = P.dir().x; io[1] = P.dir().y; io[2] = P.dir().z; io[3] = P.energy(); io[4] = P.speed();
}
// Watt spectrum of nuclide nuc at incident E (src/Distribution.cpp:34-73)
double refh_watt( void* p, int nuc, double E )
{
    Simulator& S = *((Harness*)p)->sim;
    return S.Nuclides[nuc]->fission()->Chi( E );
}
// isotropic direction (src/Distribution.cpp:78-92)
void refh_isotropic( double* out )
{
    DistributionIsotropicDirection d;
    Point q = d.sample();
    out[0] = q.x; out[1] = q.y; out[2] = q.z;
}
// particle energy<->speed constants (src/Particle.cpp:42-56)
void refh_particle_speed( void* p, double E, double* out )
{
    Simulator& S = *((Harness*)p)->sim;
    Particle P( Point( 0, 0, 0 ), Point( 1, 0, 0 ), E, 0.0, 1.0, 0, S.Cells[0] );
    out[0] = P.speed();
    P.set_speed( out[0] );
    out[1] = P.energy();
}

// run the whole reference simulation in-process and return k per cycle (ksearch) and tally means
int refh_run( void* p, double* k_cycle, int max_cycles )
{
    Simulator& S = *((Harness*)p)->sim;
    S.start();
    return 0;
}
double refh_tally( void* p, int est, int idx, int what )
{
    Simulator& S = *((Harness*)p)->sim;
    Tally t = S.Estimators[est]->tally( idx );
    return what == 0 ? t.mean : t.uncer;
}
unsigned long long refh_ntrack( void* p ) { return ((Harness*)p)->sim->Ntrack; }
double refh_k( void* p ) { return ((Harness*)p)->sim->k; }

}  // extern "C"
