// Host check of controlar_b200/csrc/pk_plan.h (the per-token attention work split of the persistent decode kernel):
// the parts of all CTAs / warps must tile the flattened (pair, key) space exactly once, in order, and the per-segment
// records (key range, owner, warp range, part mask, first CTA) must describe exactly those parts.
// usage: pk_plan_check grid b_eff H n_lo n_hi   -> prints "ok <cases>" or the first violation, exit code 1
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../controlar_b200/csrc/pk_plan.h"

#define FAIL(...) do { printf("FAIL n=%d cta=%d: ", n, c); printf(__VA_ARGS__); printf("\n"); return 1; } while (0)

int main(int argc, char** argv) {
    if (argc < 6) return 2;
    const int grid = atoi(argv[1]), b_eff = atoi(argv[2]), H = atoi(argv[3]), n_lo = atoi(argv[4]), n_hi = atoi(argv[5]);
    const int nbh = b_eff * H;
    long long cases = 0;
    for (int n = n_lo; n <= n_hi; ++n) {
        const long long tot = (long long)nbh * n;
        long long cursor = 0;                       // next flat index that must be covered
        std::vector<int> first_cta_of(nbh, -1);
        int c = 0;
        for (c = 0; c < grid; ++c) {
            PkAttnPlan pl;
            pl.active = -1;
            for (int i = 0; i < 2 * PKP_WARPS + PKP_MAXSEG; ++i) pkp_fill(pl, i, c, grid, nbh, H, n);
            const bool should = (long long)c < (tot < grid ? tot : (long long)grid);
            if ((pl.active != 0) != should) FAIL("active %d", pl.active);
            if (!should) continue;
            if (pl.n != n) FAIL("n");
            const long long c_start = cursor;
            for (int w = 0; w < PKP_WARPS; ++w)
                for (int p = 0; p < 2; ++p) {
                    const PkPart& q = pl.part[w][p];
                    if (q.k0 >= q.k1) continue;
                    if (q.k0 < 0 || q.k1 > n || q.bh < 0 || q.bh >= nbh || q.b != q.bh / H) FAIL("part range w=%d p=%d", w, p);
                    if ((long long)q.bh * n + q.k0 != cursor) FAIL("gap/overlap at w=%d p=%d: start %lld cursor %lld", w, p, (long long)q.bh * n + q.k0, cursor);
                    cursor = (long long)q.bh * n + q.k1;
                }
            // CTA ranges are balanced: floor((c+1) tot / G) boundaries
            const long long G = tot < grid ? tot : grid;
            if (c_start != (long long)c * tot / G || cursor != (long long)(c + 1) * tot / G) FAIL("cta range");
            if (cursor == c_start) FAIL("empty cta");
            const int pair_lo = (int)(c_start / n), pair_hi = (int)((cursor - 1) / n);
            if (pl.pair_lo != pair_lo || pl.nseg != pair_hi - pair_lo + 1) FAIL("pairs %d %d nseg %d", pair_lo, pair_hi, pl.nseg);
            if (pl.nseg > PKP_MAXSEG) FAIL("nseg > max (host must reject this config)");
            for (int s = 0; s < pl.nseg; ++s) {
                const PkSegPlan& g = pl.seg[s];
                if (g.bh != pair_lo + s || g.b != g.bh / H || g.hd != g.bh % H) FAIL("seg ids");
                int kmin = 1 << 30, kmax = -1;
                for (int w = 0; w < PKP_WARPS; ++w) {
                    int hits = 0, which = -1;
                    for (int p = 0; p < 2; ++p) {
                        const PkPart& q = pl.part[w][p];
                        if (q.k0 < q.k1 && q.bh == g.bh) { ++hits; which = p; if (q.k0 < kmin) kmin = q.k0; if (q.k1 > kmax) kmax = q.k1; }
                    }
                    const bool in = w >= g.w0 && w <= g.w1;
                    if (hits != (in ? 1 : 0)) FAIL("seg %d warp %d: hits %d, w0 %d w1 %d", s, w, hits, g.w0, g.w1);
                    if (in && (int)((g.part_mask >> w) & 1u) != which) FAIL("seg %d warp %d: part mask", s, w);
                }
                if (g.part_mask >> (g.w1 + 1)) FAIL("mask bits above w1");
                if (kmin != g.ks || kmax != g.ke) FAIL("seg %d keys [%d,%d) vs parts [%d,%d)", s, g.ks, g.ke, kmin, kmax);
                if (g.owner != (g.ke == n)) FAIL("owner");
                if (first_cta_of[g.bh] < 0) { first_cta_of[g.bh] = c; if (g.ks != 0) FAIL("first cta of a pair must start at key 0"); }
                if (g.first_cta != first_cta_of[g.bh]) FAIL("first_cta %d vs %d", g.first_cta, first_cta_of[g.bh]);
                ++cases;
            }
        }
        c = -1;
        if (cursor != tot) FAIL("coverage ends at %lld of %lld", cursor, tot);
    }
    printf("ok %lld\n", cases);
    return 0;
}
