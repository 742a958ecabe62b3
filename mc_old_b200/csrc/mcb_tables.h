// mcb_tables.h — per-material unionized energy grid with a hashed bin index.
//
// Replaces the per-nuclide binary search of the reference (Nuclide::checkE ->
// binary_search, src/Nuclide.cpp:18-24, src/Algorithm.cpp:46-64) by ONE search per
// (particle, energy) on the union of the material's nuclide grids:
//
//   u      = #{U < E} - 1                        (strict <, like Algorithm.cpp:46-64)
//   idx_n  = map[u*Nn + n] = #{n_E <= U[u]} - 1   (idx_n = -1 when u = -1)
//
// which equals the reference's idx_n = #{n_E < E} - 1 for every E, duplicates in n_E
// included.  The search range is narrowed by a hash on the bit
// pattern of E: key = (bits(E) >> shift) - key_min is an exactly monotone
// piecewise-linear stand-in for lethargy (no libm, no rounding), and
// hash[key] = #{U < lower edge of bin key}.
#ifndef MCB_TABLES_H
#define MCB_TABLES_H

#include <stdint.h>
#include <string.h>

#include "mcb200.h"

#define MCB_HASH_BITS_DEFAULT 16 /* mantissa bits of the hash key at most (bins per octave = 2^bits), see mcb_tables.cpp */
#define MCB_MAP_BISECT (-2) /* map entry of a nuclide whose grid is not ascending: bisect its rows like the reference */

#if defined(__CUDACC__)
#define MCB_THD __host__ __device__ __forceinline__
#else
#define MCB_THD inline
#endif

// #{U < E} restricted by the hash; works on host and device
MCB_THD int mcb_union_count_less(const double* U, const int32_t* hash, int64_t key_min, int32_t n_hash, int32_t shift,
                                 int32_t nU, double E)
{
    int64_t bits;
#if defined(__CUDA_ARCH__)
    bits = __double_as_longlong(E);
#else
    memcpy(&bits, &E, sizeof(bits));
#endif
    const int64_t key = (bits >> shift) - key_min;  // arithmetic shift: negative E -> negative key
    if (key < 0) return 0;
    if (key >= n_hash) return nU;
    int lo = hash[key], hi = hash[key + 1];
    while (lo < hi) {  // lower_bound: first element not < E
        const int mid = (lo + hi) >> 1;
        if (U[mid] < E) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// The same search through the bin RECORDS: record key = { u0 = #{U < lower edge of bin key}, cnt = points of U inside
// the bin, idx_n = map[(u0 - 1) * Nn + n] (-1 for u0 = 0) }, `stride` ints apiece; two more records stand for energies
// above the last bin (index n_hash) and below the first (index n_hash + 1).  Most bins hold no grid point (HEU: 204 k
// points in 672 k bins; none at all above the resolved resonances, where a fast system spends its collisions), and for
// those the per-nuclide indices come straight from the record: the chain hash -> union grid -> map -> rows of
// dependent loads becomes record -> rows.  Returns lo = #{U < E}; *rec_out = the record; *from_rec tells whether
// the record's indices apply (else read map[(lo - 1) * Nn + n]).
MCB_THD int mcb_union_lookup(const double* U, const int32_t* hrec, int32_t stride, int64_t key_min, int32_t n_hash, int32_t shift,
                             double E, const int32_t** rec_out, bool* from_rec)
{
    int64_t bits;
#if defined(__CUDA_ARCH__)
    bits = __double_as_longlong(E);
#else
    memcpy(&bits, &E, sizeof(bits));
#endif
    int64_t key = (bits >> shift) - key_min;
    if (key < 0) key = (int64_t)n_hash + 1;
    else if (key > n_hash) key = n_hash;
    const int32_t* rec = hrec + (size_t)key * stride;
    *rec_out = rec;
#if defined(__CUDA_ARCH__)
    const int2 h = __ldg(reinterpret_cast<const int2*>(rec));
    const int u0 = h.x, cnt = h.y;
#else
    const int u0 = rec[0], cnt = rec[1];
#endif
    int lo = u0, hi = u0 + cnt;
    while (lo < hi) {  // lower_bound inside the bin: first element not < E
        const int mid = (lo + hi) >> 1;
        if (U[mid] < E) lo = mid + 1; else hi = mid;
    }
    *from_rec = lo == u0;
    return lo;
}

// binary_search (Algorithm.cpp:46-64) over column 0 of n rows of MCB_XS_ROW doubles: the probe sequence of the
// reference, whatever the order of the grid
MCB_THD int mcb_row_bisect(const double* rows, int n, double E)
{
    int left = 0, right = n - 1;
    while (left <= right) {
        const int mid = (left + right) / 2;
        if (rows[(size_t)mid * MCB_XS_ROW] < E) left = mid + 1; else right = mid - 1;
    }
    return right;
}

#ifdef __cplusplus
#include <vector>
namespace mcb {

struct MaterialTables {
    std::vector<double> U;       // distinct energies of all nuclides of the material, ascending
    std::vector<int32_t> map;    // nU x n_nuc
    std::vector<int32_t> hash;   // n_hash + 1
    std::vector<int32_t> hrec;   // (n_hash + 2) bin records of hrec_stride ints: { u0, cnt, idx_0 .. idx_{Nn-1}, pad }
    int32_t hrec_stride = 0;
    int64_t key_min = 0;
    int32_t n_hash = 0, shift = 0, n_nuc = 0;
    int32_t n_bisect = 0;        // nuclides whose grid is not ascending
    int32_t max_bin = 0;         // largest number of grid points in one hash bin (search depth statistics)
};

// max_mant_bits: mantissa bits kept in the key at most (bins per octave = 2^bits)
void build_material_tables(const mcb_problem* p, int material, int max_mant_bits, MaterialTables& out);

// out[2 s + (side > 0)] = the cell search_cell (general.cpp:26-34) is bound to return for ANY point strictly on that side
// of surface s, or -1 where that cannot be told beforehand (see mcb_api.cu / ev_cross_pre); 2 * max(n_surfaces, 1) entries
void build_cross_neighbors(const mcb_problem* p, std::vector<int32_t>& out);

}  // namespace mcb
#endif
#endif
