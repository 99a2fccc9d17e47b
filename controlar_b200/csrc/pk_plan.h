// pk_plan.h — per-token work split of the attention phase of the persistent decode kernel (decode_persistent.cuh).
// Plain integer code, host- and device-compilable: tests/test_pk_plan_cpu.py checks it on the CPU against a brute-force
// enumeration of the flattened (sequence, head, key) space.
//
// The split depends only on the context length n (= pos + 1), not on the layer, so it is computed ONCE per token into
// shared memory and read by all L attention phases (the 64-bit divisions it needs cost ~10 subroutine calls per warp).
//
//   flat space   f = (b * H + h) * n + key,  tot = b_eff * H * n,  G = min(grid, tot) participating CTAs
//   CTA c        [f0, f1) = [c tot / G, (c + 1) tot / G)
//   warp w       [wa, wb) = [min(f1, f0 + w Cw), min(f1, wa + Cw)),  Cw = ceil((f1 - f0) / 16)
//   part 0 / 1   the piece of the warp range inside the (b, h) pair of wa / inside the next pair (a warp range touches at
//                most two pairs when Cw <= n, which holds whenever a CTA range is at most 16 pairs long — host-checked)
//   segment s    pair pair_lo + s of the CTA: keys [ks, ke); the CTA owning the pair's LAST key is its owner and combines the
//                partials of the CTAs before it (first_cta .. c - 1)
#pragma once

#if defined(__CUDACC__)
#define PK_HD __host__ __device__ __forceinline__
#else
#define PK_HD inline
#endif

constexpr int PKP_WARPS = 16;
constexpr int PKP_MAXSEG = 6;

struct PkPart { int bh, b, k0, k1; };                    // keys [k0, k1) of pair bh (sequence b); k0 >= k1: empty
struct PkSegPlan {
    int bh, b, hd, ks, ke;                                // pair, sequence, head, key range of this CTA
    int owner;                                            // this CTA holds the pair's last key
    int w0, w1;                                           // warps whose range touches the pair
    unsigned part_mask;                                   // bit w: warp w's entry for this pair is its part 1 (else part 0)
    int first_cta;                                        // CTA holding the pair's first key
};
struct PkAttnPlan {
    int active, nseg, pair_lo, n;
    PkPart part[PKP_WARPS][2];
    PkSegPlan seg[PKP_MAXSEG];
};

PK_HD long long pkp_min(long long a, long long b) { return a < b ? a : b; }
PK_HD long long pkp_max(long long a, long long b) { return a > b ? a : b; }

struct PkSplit { long long tot, f0, f1, Cw; int G, pair_lo, pair_hi, nseg; };

PK_HD PkSplit pkp_split(int cta, int grid, int nbh, int n) {
    PkSplit s;
    s.tot = (long long)nbh * n;
    s.G = (int)pkp_min((long long)grid, s.tot);
    s.f0 = s.f1 = 0; s.Cw = 1; s.pair_lo = s.pair_hi = 0; s.nseg = 0;
    if (cta >= s.G) return s;
    s.f0 = ((long long)cta * s.tot) / s.G;
    s.f1 = ((long long)(cta + 1) * s.tot) / s.G;
    s.Cw = (s.f1 - s.f0 + PKP_WARPS - 1) / PKP_WARPS;
    s.pair_lo = (int)(s.f0 / n);
    s.pair_hi = (int)((s.f1 - 1) / n);
    s.nseg = (int)pkp_min(s.pair_hi - s.pair_lo + 1, PKP_MAXSEG);
    return s;
}

// CTA that holds flat index f
PK_HD int pkp_cta_of_flat(long long f, long long tot, int G) {
    int c = (int)((f * G) / tot);
    while (c + 1 < G && ((long long)(c + 1) * tot) / G <= f) ++c;
    while (c > 0 && ((long long)c * tot) / G > f) --c;
    return c;
}

PK_HD PkPart pkp_part(const PkSplit& s, int H, int n, int warp, int part) {
    PkPart r; r.bh = 0; r.b = 0; r.k0 = 0; r.k1 = 0;
    const long long wa = pkp_min(s.f1, s.f0 + (long long)warp * s.Cw), wb = pkp_min(s.f1, wa + s.Cw);
    const int bh0 = (int)(wa / n);
    const long long bound = pkp_min(wb, (long long)(bh0 + 1) * n);
    const long long pa = part == 0 ? wa : bound, pb = part == 0 ? bound : wb;
    if (pa >= pb) return r;
    r.bh = (int)(pa / n); r.b = r.bh / H;
    r.k0 = (int)(pa - (long long)r.bh * n); r.k1 = (int)(pb - (long long)r.bh * n);
    return r;
}

PK_HD PkSegPlan pkp_seg(const PkSplit& s, int H, int n, int cta, int sg) {
    PkSegPlan q;
    q.bh = s.pair_lo + sg; q.b = q.bh / H; q.hd = q.bh - q.b * H;
    q.ks = sg == 0 ? (int)(s.f0 - (long long)q.bh * n) : 0;
    q.ke = sg == s.nseg - 1 ? (int)(s.f1 - (long long)q.bh * n) : n;
    q.owner = q.ke == n;
    const long long pair_start = (long long)q.bh * n;
    const long long ps = pkp_max(s.f0, pair_start), pe = pkp_min(s.f1, pair_start + n);
    q.w0 = (int)((ps - s.f0) / s.Cw); q.w1 = (int)((pe - 1 - s.f0) / s.Cw);
    q.part_mask = 0u;
    for (int w = q.w0; w <= q.w1; ++w)
        if (s.f0 + (long long)w * s.Cw < pair_start) q.part_mask |= 1u << w;   // the warp range starts in the pair before
    q.first_cta = q.ks > 0 ? pkp_cta_of_flat(pair_start, s.tot, s.G) : cta;
    return q;
}

// the whole plan of one CTA, entry `i` of 32 + PKP_MAXSEG (one thread each in the kernel)
PK_HD void pkp_fill(PkAttnPlan& pl, int i, int cta, int grid, int nbh, int H, int n) {
    const PkSplit s = pkp_split(cta, grid, nbh, n);
    if (i == 0) { pl.active = cta < s.G; pl.nseg = s.nseg; pl.pair_lo = s.pair_lo; pl.n = n; }
    if (cta >= s.G) return;
    if (i < 2 * PKP_WARPS) pl.part[i >> 1][i & 1] = pkp_part(s, H, n, i >> 1, i & 1);
    else if (i - 2 * PKP_WARPS < s.nseg) pl.seg[i - 2 * PKP_WARPS] = pkp_seg(s, H, n, cta, i - 2 * PKP_WARPS);
}
