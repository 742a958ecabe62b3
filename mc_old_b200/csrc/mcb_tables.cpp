#include "mcb_tables.h"

#include <algorithm>
#include <cstdlib>

namespace mcb {

static int64_t key_of(double E, int shift)
{
    int64_t bits;
    memcpy(&bits, &E, sizeof(bits));
    return bits >> shift;
}

void build_material_tables(const mcb_problem* p, int material, int max_mant_bits, MaterialTables& T)
{
    const int nb = p->mat_begin[material], ne = p->mat_begin[material + 1];
    T = MaterialTables();
    T.n_nuc = ne - nb;
    // union of the nuclide grids (column 0 of the xs rows)
    for (int i = nb; i < ne; i++) {
        const mcb_nuclide& N = p->nuclides[p->mat_nuclide[i]];
        const double* rows = p->xs_rows + (size_t)N.row_begin * MCB_XS_ROW;
        for (int r = 0; r < N.n_rows; r++) T.U.push_back(rows[(size_t)r * MCB_XS_ROW]);
    }
    std::sort(T.U.begin(), T.U.end());
    T.U.erase(std::unique(T.U.begin(), T.U.end()), T.U.end());
    const int nU = (int)T.U.size();
    // map[u][n] = #{n_E <= U[u]} - 1 : one merge pass per nuclide.  A grid that is not ascending (xs_library/005011.txt
    // has two descents) has no such count: the reference's bisection (Algorithm.cpp:46-64) then returns whatever its
    // probe sequence leads to, so the column is marked MCB_MAP_BISECT and the device repeats that bisection.
    T.map.assign((size_t)nU * T.n_nuc, -1);
    for (int i = nb; i < ne; i++) {
        const mcb_nuclide& N = p->nuclides[p->mat_nuclide[i]];
        const double* rows = p->xs_rows + (size_t)N.row_begin * MCB_XS_ROW;
        bool ascending = true;
        for (int r = 1; r < N.n_rows; r++) if (rows[(size_t)r * MCB_XS_ROW] < rows[(size_t)(r - 1) * MCB_XS_ROW]) ascending = false;
        if (!ascending) {
            for (int u = 0; u < nU; u++) T.map[(size_t)u * T.n_nuc + (i - nb)] = MCB_MAP_BISECT;
            T.n_bisect++;
            continue;
        }
        int r = 0;
        for (int u = 0; u < nU; u++) {
            while (r < N.n_rows && rows[(size_t)r * MCB_XS_ROW] <= T.U[u]) r++;
            T.map[(size_t)u * T.n_nuc + (i - nb)] = r - 1;
        }
    }
    if (nU == 0) {
        T.hash.assign(1, 0);
        T.hrec_stride = (2 + T.n_nuc + 1) & ~1;
        T.hrec.assign((size_t)2 * T.hrec_stride, 0);
        for (int r = 0; r < 2; r++) for (int n = 0; n < T.n_nuc; n++) T.hrec[(size_t)r * T.hrec_stride + 2 + n] = -1;
        return;
    }
    // hash on the bit pattern; the table may hold up to ~16 entries per grid point (measured on HEU, one B200: a walk step
    // costs more than a bin - 14 bits / 4 per point 5.75 ms per generation, 16 bits / 16 per point 5.65 ms, 18 / 64 the same,
    // 12 bits 5.90, 8 bits 6.25; profiles/r2f_hash_bits.txt).  MCB_HASH_PER_POINT / MCB_HASH_BITS are tuning knobs.
    const int64_t per_point = getenv("MCB_HASH_PER_POINT") ? std::max(1, atoi(getenv("MCB_HASH_PER_POINT"))) : 16;
    const int64_t cap = std::max<int64_t>(per_point * (int64_t)nU + 1024, 4096);
    int bits = std::min(std::max(max_mant_bits, 0), 20);
    for (;; bits--) {
        T.shift = 52 - bits;
        T.key_min = key_of(T.U.front(), T.shift);
        const int64_t n = key_of(T.U.back(), T.shift) - T.key_min + 1;
        if (n <= cap || bits == 0) { T.n_hash = (int32_t)std::min<int64_t>(n, INT32_MAX - 2); break; }
    }
    // hash[b] = #{U < lower edge of bin b} = #{U with key < b}; keys are monotone in U (all U >= 0 here; a
    // negative energy would sort before by key as well because the shift is arithmetic)
    T.hash.assign((size_t)T.n_hash + 1, 0);
    {
        int u = 0;
        for (int64_t b = 0; b <= T.n_hash; b++) {
            while (u < nU && key_of(T.U[u], T.shift) - T.key_min < b) u++;
            T.hash[(size_t)b] = u;
        }
    }
    for (int64_t b = 0; b < T.n_hash; b++) T.max_bin = std::max(T.max_bin, T.hash[b + 1] - T.hash[b]);
    // bin records (see mcb_union_lookup): the search window and the per-nuclide indices below it in one place
    T.hrec_stride = (2 + T.n_nuc + 1) & ~1;  // whole 8-byte words
    T.hrec.assign((size_t)(T.n_hash + 2) * T.hrec_stride, 0);
    auto fill = [&](int64_t r, int u0, int cnt) {
        int32_t* rec = T.hrec.data() + (size_t)r * T.hrec_stride;
        rec[0] = u0; rec[1] = cnt;
        for (int n = 0; n < T.n_nuc; n++) rec[2 + n] = u0 > 0 ? T.map[(size_t)(u0 - 1) * T.n_nuc + n] : -1;
    };
    for (int64_t b = 0; b < T.n_hash; b++) fill(b, T.hash[(size_t)b], T.hash[(size_t)b + 1] - T.hash[(size_t)b]);
    fill(T.n_hash, nU, 0);      // above the last bin
    fill(T.n_hash + 1, 0, 0);   // below the first
    for (int n = 0; n < T.n_nuc; n++)  // a grid that is not ascending is bisected whatever the energy
        if (nU && T.map[n] == MCB_MAP_BISECT)
            for (int64_t r = 0; r < T.n_hash + 2; r++)
                if (T.hrec[(size_t)r * T.hrec_stride] > 0) T.hrec[(size_t)r * T.hrec_stride + 2 + n] = MCB_MAP_BISECT;
}

void build_cross_neighbors(const mcb_problem* p, std::vector<int32_t>& nb)
{
    nb.assign((size_t)std::max(2 * p->n_surfaces, 2), -1);  // [2 s] side -1, [2 s + 1] side +1
    for (int s = 0; s < p->n_surfaces; s++)
        for (int k = 0; k < 2; k++) {
            const int side = k ? 1 : -1;
            for (int b = 0; b < p->n_cells; b++) {
                const mcb_cell& B = p->cells[b];
                bool ruled_out = false;  // the cell holds (s, -side): test_point puts every point of this side outside it
                for (int j = B.surf_begin; j < B.surf_end; j++) if (p->cell_surface[j] == s && p->cell_sense[j] != side) ruled_out = true;
                if (ruled_out) continue;
                // the first cell in deck order that the side does not rule out: the answer if it is (s, side) and nothing else
                if (B.surf_end - B.surf_begin == 1 && p->cell_surface[B.surf_begin] == s && p->cell_sense[B.surf_begin] == side) nb[(size_t)(2 * s + k)] = b;
                break;
            }
        }
}

}  // namespace mcb
