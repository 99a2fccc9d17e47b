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
//
// Helper piece first (r2): when the CTA's LAST segment does not hold its pair's last key, the CTA is only a helper for that
// pair.  That piece [fh, f1) — the "H range" — is processed FIRST by all 16 warps (16 equal warp ranges of it, `hk`), and its
// partial is published before the CTA starts on its owner segments — the "O range" [f0, fh), cut into the 16 warp ranges that
// `part` describes.  The owner of the pair (the next CTA) then finds the partial already in L2 when it finishes its own keys:
// the helper -> owner round trip is off the critical path of the phase (it cost ~5 us per layer when both finalised at the end).
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
    int has_h, h_w1, pad0_, pad1_;                        // (16-byte alignment of `part`) last segment is a helper piece (processed first); last warp with keys of it
    int hk[PKP_WARPS][2];                                 // H pass: keys [k0, k1) of pair seg[nseg - 1].bh per warp
    PkPart part[PKP_WARPS][2];                            // O pass: the CTA range minus the helper piece
    PkSegPlan seg[PKP_MAXSEG];                            // (seg[nseg - 1] is the helper piece when has_h)
};

PK_HD long long pkp_min(long long a, long long b) { return a < b ? a : b; }
PK_HD long long pkp_max(long long a, long long b) { return a > b ? a : b; }

struct PkSplit { long long tot, f0, f1, fh, Cw, Ch; int G, pair_lo, pair_hi, nseg, has_h; };

PK_HD PkSplit pkp_split(int cta, int grid, int nbh, int n) {
    PkSplit s;
    s.tot = (long long)nbh * n;
    s.G = (int)pkp_min((long long)grid, s.tot);
    s.f0 = s.f1 = s.fh = 0; s.Cw = 1; s.Ch = 1; s.pair_lo = s.pair_hi = 0; s.nseg = 0; s.has_h = 0;
    if (cta >= s.G) return s;
    s.f0 = ((long long)cta * s.tot) / s.G;
    s.f1 = ((long long)(cta + 1) * s.tot) / s.G;
    s.pair_lo = (int)(s.f0 / n);
    s.pair_hi = (int)((s.f1 - 1) / n);
    s.nseg = (int)pkp_min(s.pair_hi - s.pair_lo + 1, PKP_MAXSEG);
    s.has_h = s.f1 < (long long)(s.pair_hi + 1) * n;          // the last pair continues in the next CTA
    s.fh = s.has_h ? pkp_max(s.f0, (long long)s.pair_hi * n) : s.f1;
    s.Cw = pkp_max(1, (s.fh - s.f0 + PKP_WARPS - 1) / PKP_WARPS);
    s.Ch = pkp_max(1, (s.f1 - s.fh + PKP_WARPS - 1) / PKP_WARPS);
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
    const long long wa = pkp_min(s.fh, s.f0 + (long long)warp * s.Cw), wb = pkp_min(s.fh, wa + s.Cw);
    const int bh0 = (int)(wa / n);
    const long long bound = pkp_min(wb, (long long)(bh0 + 1) * n);
    const long long pa = part == 0 ? wa : bound, pb = part == 0 ? bound : wb;
    if (pa >= pb) return r;
    r.bh = (int)(pa / n); r.b = r.bh / H;
    r.k0 = (int)(pa - (long long)r.bh * n); r.k1 = (int)(pb - (long long)r.bh * n);
    return r;
}

// H pass: warp's keys [k0, k1) of pair s.pair_hi (empty when the CTA has no helper piece)
PK_HD void pkp_hpart(const PkSplit& s, int n, int warp, int& k0, int& k1) {
    const long long ha = pkp_min(s.f1, s.fh + (long long)warp * s.Ch), hb = pkp_min(s.f1, ha + s.Ch);
    k0 = (int)(ha - (long long)s.pair_hi * n); k1 = (int)(hb - (long long)s.pair_hi * n);
    if (!s.has_h) { k0 = 0; k1 = 0; }
}

PK_HD PkSegPlan pkp_seg(const PkSplit& s, int H, int n, int cta, int sg) {
    PkSegPlan q;
    q.bh = s.pair_lo + sg; q.b = q.bh / H; q.hd = q.bh - q.b * H;
    q.ks = sg == 0 ? (int)(s.f0 - (long long)q.bh * n) : 0;
    q.ke = sg == s.nseg - 1 ? (int)(s.f1 - (long long)q.bh * n) : n;
    q.owner = q.ke == n;
    const long long pair_start = (long long)q.bh * n;
    q.part_mask = 0u;
    if (s.has_h && sg == s.nseg - 1) {                     // the helper piece: all warps with keys of the H range
        q.w0 = 0; q.w1 = (int)((s.f1 - s.fh - 1) / s.Ch);
    } else {
        const long long ps = pkp_max(s.f0, pair_start), pe = pkp_min(s.fh, pair_start + n);
        q.w0 = (int)((ps - s.f0) / s.Cw); q.w1 = (int)((pe - 1 - s.f0) / s.Cw);
        for (int w = q.w0; w <= q.w1; ++w)
            if (s.f0 + (long long)w * s.Cw < pair_start) q.part_mask |= 1u << w;   // the warp range starts in the pair before
    }
    q.first_cta = q.ks > 0 ? pkp_cta_of_flat(pair_start, s.tot, s.G) : cta;
    return q;
}

// the whole plan of one CTA, entry `i` of 32 + PKP_MAXSEG (one thread each in the kernel)
PK_HD void pkp_fill(PkAttnPlan& pl, int i, int cta, int grid, int nbh, int H, int n) {
    const PkSplit s = pkp_split(cta, grid, nbh, n);
    if (i == 0) {
        pl.active = cta < s.G; pl.nseg = s.nseg; pl.pair_lo = s.pair_lo; pl.n = n;
        pl.has_h = s.has_h; pl.h_w1 = s.has_h ? (int)((s.f1 - s.fh - 1) / s.Ch) : -1;
    }
    if (cta >= s.G) return;
    if (i < PKP_WARPS) pkp_hpart(s, n, i, pl.hk[i][0], pl.hk[i][1]);
    if (i < 2 * PKP_WARPS) pl.part[i >> 1][i & 1] = pkp_part(s, H, n, i >> 1, i & 1);
    else if (i - 2 * PKP_WARPS < s.nseg) pl.seg[i - 2 * PKP_WARPS] = pkp_seg(s, H, n, cta, i - 2 * PKP_WARPS);
}
