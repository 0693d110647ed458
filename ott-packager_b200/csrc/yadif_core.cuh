// yadif mode 0 (one output frame per input frame, first field kept) for one sample, as the reference gets it from
// libavfilter "yadif=0:-1:{0|1}" (transvideo.c:2036-2052).  Follows upstream vf_yadif.c filter_line_c / filter_edges /
// filter_slice as summarised in SURVEY.md Appendix C.  Shared by the kernel and the CPU emulation used in tests.
#pragma once
#include "ladder_core.cuh"

namespace ottgpu {

OTT_DEV int ott_abs(int v) { return v < 0 ? -v : v; }

// Pointers address the sample; `up`/`dn` are element offsets to the rows above/below inside the same frame.
template <typename S>
OTT_DEV int yadif_sample(const S *prev, const S *cur, const S *next, const S *p2, const S *n2, int up, int dn,
                         bool search, bool spatial)
{
    const int c = cur[up], e = cur[dn];
    const int d = ((int)p2[0] + (int)n2[0]) >> 1;
    const int td0 = ott_abs((int)p2[0] - (int)n2[0]);
    const int td1 = (ott_abs((int)prev[up] - c) + ott_abs((int)prev[dn] - e)) >> 1;
    const int td2 = (ott_abs((int)next[up] - c) + ott_abs((int)next[dn] - e)) >> 1;
    int diff = ott_max(ott_max(td0 >> 1, td1), td2);
    int pred = (c + e) >> 1;
    if (search) {                                         // edge-directed interpolation, only 3 <= x < w-3
        const int um3 = cur[up - 3], um2 = cur[up - 2], um1 = cur[up - 1], up1 = cur[up + 1], up2 = cur[up + 2], up3 = cur[up + 3];
        const int dm3 = cur[dn - 3], dm2 = cur[dn - 2], dm1 = cur[dn - 1], dp1 = cur[dn + 1], dp2 = cur[dn + 2], dp3 = cur[dn + 3];
        int score = ott_abs(um1 - dm1) + ott_abs(c - e) + ott_abs(up1 - dp1) - 1;
        int s = ott_abs(um2 - e) + ott_abs(um1 - dp1) + ott_abs(c - dp2);                 // j = -1
        if (s < score) {
            score = s; pred = (um1 + dp1) >> 1;
            s = ott_abs(um3 - dp1) + ott_abs(um2 - dp2) + ott_abs(um1 - dp3);             // j = -2
            if (s < score) { score = s; pred = (um2 + dp2) >> 1; }
        }
        s = ott_abs(c - dm2) + ott_abs(up1 - dm1) + ott_abs(up2 - e);                     // j = +1
        if (s < score) {
            score = s; pred = (up1 + dm1) >> 1;
            s = ott_abs(up1 - dm3) + ott_abs(up2 - dm2) + ott_abs(up3 - dm1);             // j = +2
            if (s < score) { score = s; pred = (up2 + dm2) >> 1; }
        }
    }
    if (spatial) {                                        // skipped on rows 1 and h-2 (mode 2 there)
        const int b = ((int)p2[2 * up] + (int)n2[2 * up]) >> 1;
        const int f = ((int)p2[2 * dn] + (int)n2[2 * dn]) >> 1;
        const int mx = ott_max(ott_max(d - e, d - c), ott_min(b - c, f - e));
        const int mn = ott_min(ott_min(d - e, d - c), ott_max(b - c, f - e));
        diff = ott_max(ott_max(diff, mn), -mx);
    }
    return ott_min(ott_max(pred, d - diff), d + diff);
}

template <typename S>
OTT_DEV void yadif_plane_px(const YadifJob &j, size_t off, int w, int h, int x, int y)
{
    const S *prev = reinterpret_cast<const S *>(j.prev) + off, *cur = reinterpret_cast<const S *>(j.cur) + off;
    const S *next = reinterpret_cast<const S *>(j.next) + off;
    S *dst = reinterpret_cast<S *>(j.dst) + off;
    const size_t i = (size_t)y * w + x;
    if (!((y ^ j.parity) & 1)) { dst[i] = cur[i]; return; }
    const int dn = y + 1 < h ? w : -w, up = y ? -w : w;
    const bool first_field = (j.parity ^ j.tff) != 0;
    const S *p2 = first_field ? prev : cur, *n2 = first_field ? cur : next;
    dst[i] = (S)yadif_sample<S>(prev + i, cur + i, next + i, p2 + i, n2 + i, up, dn, x >= 3 && x < w - 3,
                                !(y == 1 || y + 2 == h));
}

// One thread position (x, y) of the luma grid; positions inside the chroma grid also produce one U and one V sample.
OTT_DEV void yadif_position(const YadifJob &j, int x, int y)
{
    const int w = j.width, h = j.height, cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    if (x < w && y < h) {
        if (j.bps == 1) yadif_plane_px<uint8_t>(j, 0, w, h, x, y);
        else yadif_plane_px<uint16_t>(j, 0, w, h, x, y);
    }
    if (x < cw && y < ch) {
        const size_t uo = (size_t)w * h, vo = uo + (size_t)cw * ch;
        if (j.bps == 1) { yadif_plane_px<uint8_t>(j, uo, cw, ch, x, y); yadif_plane_px<uint8_t>(j, vo, cw, ch, x, y); }
        else { yadif_plane_px<uint16_t>(j, uo, cw, ch, x, y); yadif_plane_px<uint16_t>(j, vo, cw, ch, x, y); }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Fast path (8-bit, plane widths multiples of 8, 4-byte aligned frames): one thread = 8 adjacent samples (two 32-bit
// words).  Every neighbour row is fetched as aligned words (4 words of `cur` above and below cover the +-3 sample search
// window of all eight samples) and the per-sample predictor is branch-free: |a-b|+c is one VABSDIFF, the nested
// direction search of filter_line_c becomes predicated selects.
OTT_DEV int ybyte(uint32_t w, int i) { return ott_byte(w, i); }

// kept-field row: 8 samples copied
OTT_DEV void yadif_keep8(const YadifJob &j, size_t off, int w, int x8, int y)
{
    const int i0 = y * (w >> 2) + 2 * x8;
    *reinterpret_cast<uint2 *>(reinterpret_cast<uint32_t *>(j.dst + off) + i0) =
        ott_ldg(reinterpret_cast<const uint2 *>(reinterpret_cast<const uint32_t *>(j.cur + off) + i0));
}

// per-byte floor((a+b)/2) of two words of four unsigned bytes
OTT_DEV uint32_t yavg4(uint32_t a, uint32_t b) { return (a & b) + (((a ^ b) & 0xfefefefeu) >> 1); }
// bytes (b0,b1) / (b2,b3) of a word as two 16-bit lanes
OTT_DEV uint32_t ylanes_lo(uint32_t w) { return ott_prmt(w, 0u, 0x4140u); }
OTT_DEV uint32_t ylanes_hi(uint32_t w) { return ott_prmt(w, 0u, 0x4342u); }

// interpolated row: 8 samples predicted.
// The edge-directed search (scores, the nested +-1/+-2 selection) runs per sample on scalars; everything that is
// "vertical" - temporal differences, the averages d/b/f, the spatial clamp of diff and the final clamp of the predictor -
// runs SIMD-in-register: four samples per VABSDIFF4 / byte average, then two samples per 16-bit-lane min/max/add.
// Lanes that can go negative carry a +256 bias (kBias) so plain 32-bit adds/subtracts never borrow across lanes.
OTT_DEV void yadif_oct8(const YadifJob &j, size_t off, int w, int h, int x8, int y)
{
    const uint32_t kBias = 0x01000100u;
    const uint32_t *prev = reinterpret_cast<const uint32_t *>(j.prev + off), *cur = reinterpret_cast<const uint32_t *>(j.cur + off);
    const uint32_t *next = reinterpret_cast<const uint32_t *>(j.next + off);
    uint32_t *dst = reinterpret_cast<uint32_t *>(j.dst + off);
    const int wq = w >> 2;                                 // words per row
    const int i0 = y * wq + 2 * x8;
    const int dn = y + 1 < h ? wq : -wq, up = y ? -wq : wq;
    const bool first_field = (j.parity ^ j.tff) != 0;
    const uint32_t *p2 = first_field ? prev : cur, *n2 = first_field ? cur : next;
    const bool spatial = !(y == 1 || y + 2 == h);
    const int xl = x8 > 0 ? -1 : 0, xr = 2 * x8 + 2 < wq ? 2 : 1;   // clamped neighbour words (values unused at the edges)
    const uint2 cum = ott_ldg(reinterpret_cast<const uint2 *>(cur + i0 + up)), cdm = ott_ldg(reinterpret_cast<const uint2 *>(cur + i0 + dn));
    const uint32_t cuw[4] = {ott_ldg(cur + i0 + up + xl), cum.x, cum.y, ott_ldg(cur + i0 + up + xr)};
    const uint32_t cdw[4] = {ott_ldg(cur + i0 + dn + xl), cdm.x, cdm.y, ott_ldg(cur + i0 + dn + xr)};
    const uint2 pu = ott_ldg(reinterpret_cast<const uint2 *>(prev + i0 + up)), pd = ott_ldg(reinterpret_cast<const uint2 *>(prev + i0 + dn));
    const uint2 nu = ott_ldg(reinterpret_cast<const uint2 *>(next + i0 + up)), nd = ott_ldg(reinterpret_cast<const uint2 *>(next + i0 + dn));
    const uint2 p20 = ott_ldg(reinterpret_cast<const uint2 *>(p2 + i0)), n20 = ott_ldg(reinterpret_cast<const uint2 *>(n2 + i0));
    uint2 p2u, n2u, p2d, n2d;
    p2u.x = p2u.y = n2u.x = n2u.y = p2d.x = p2d.y = n2d.x = n2d.y = 0;
    if (spatial) {
        p2u = ott_ldg(reinterpret_cast<const uint2 *>(p2 + i0 + 2 * up)); n2u = ott_ldg(reinterpret_cast<const uint2 *>(n2 + i0 + 2 * up));
        p2d = ott_ldg(reinterpret_cast<const uint2 *>(p2 + i0 + 2 * dn)); n2d = ott_ldg(reinterpret_cast<const uint2 *>(n2 + i0 + 2 * dn));
    }
    // ---- temporal difference and spatial check first (their inputs die here): clamp bounds d-diff / d+diff, two samples
    //      per register, lanes biased by +256.  Word hw covers samples 4hw .. 4hw+3.
    uint32_t lo[4], hi[4];
    OTT_UNROLL
    for (int hw = 0; hw < 2; ++hw) {
        const uint32_t cw4 = cuw[1 + hw], ew4 = cdw[1 + hw];
        const uint32_t a4 = hw ? p20.y : p20.x, b4 = hw ? n20.y : n20.x;
        const uint32_t td0h = (ott_vabsdiff4(a4, b4) >> 1) & 0x7f7f7f7fu;                                        // |p2-n2| >> 1
        const uint32_t td1 = yavg4(ott_vabsdiff4(hw ? pu.y : pu.x, cw4), ott_vabsdiff4(hw ? pd.y : pd.x, ew4));  // (|pu-c|+|pd-e|) >> 1
        const uint32_t td2 = yavg4(ott_vabsdiff4(hw ? nu.y : nu.x, cw4), ott_vabsdiff4(hw ? nd.y : nd.x, ew4));
        const uint32_t d4 = yavg4(a4, b4);
        uint32_t bq = 0, fq = 0;
        if (spatial) {
            bq = yavg4(hw ? p2u.y : p2u.x, hw ? n2u.y : n2u.x);
            fq = yavg4(hw ? p2d.y : p2d.x, hw ? n2d.y : n2d.x);
        }
        OTT_UNROLL
        for (int half = 0; half < 2; ++half) {             // two samples per 16-bit-lane register
            const uint32_t d = half ? ylanes_hi(d4) : ylanes_lo(d4);
            uint32_t diff = ott_vmax3_s16x2(half ? ylanes_hi(td0h) : ylanes_lo(td0h), half ? ylanes_hi(td1) : ylanes_lo(td1),
                                            half ? ylanes_hi(td2) : ylanes_lo(td2));
            if (spatial) {
                const uint32_t c = half ? ylanes_hi(cw4) : ylanes_lo(cw4), e = half ? ylanes_hi(ew4) : ylanes_lo(ew4);
                const uint32_t b = half ? ylanes_hi(bq) : ylanes_lo(bq), f = half ? ylanes_hi(fq) : ylanes_lo(fq);
                const uint32_t de = d + kBias - e, dc = d + kBias - c, bc = b + kBias - c, fe = f + kBias - e;   // each lane value + 256
                const uint32_t mx = ott_vmax3_s16x2(de, dc, ott_vmin_s16x2(bc, fe));
                const uint32_t mn = ott_vmin3_s16x2(de, dc, ott_vmax_s16x2(bc, fe));
                diff = ott_vmax3_s16x2(diff + kBias, mn, 2u * kBias - mx) - kBias;                               // max(diff, mn, -mx)
            }
            lo[2 * hw + half] = d + kBias - diff;
            hi[2 * hw + half] = d + kBias + diff;
        }
    }
    // ---- direction search.  For direction j the score of sample x is A_j[x-1] + A_j[x] + A_j[x+1] with
    //      A_j[x] = |u[x+j] - d[x-j]|  (u / d = the rows above / below in cur).  A_j is built four samples at a time: the
    //      rows are byte-shifted against each other with funnel shifts and differenced with one VABSDIFF4 per word.
    //      The three-term sums are dot products of the difference words with byte masks (IDP.4A, on the multiply pipe,
    //      which this kernel otherwise leaves idle); the mask value 8 and the addend turn the sum into the key
    //      8*score + rank (rank = the order vf_yadif tries the directions in: 0, -1, -2, +1, +2), so that "take the
    //      candidate only if strictly better" becomes a plain minimum over keys.
    //      Local byte L of the 16-byte windows cuw / cdw is sample x0 - 4 + L; the thread's samples are L = 4..11.
    uint32_t A[5][4], P[5][2];                                 // [direction rank][word]: |u-d| words 0..3, averages of words 1..2
    OTT_UNROLL
    for (int r = 0; r < 5; ++r) {
        const int j = r == 0 ? 0 : (r == 1 ? -1 : (r == 2 ? -2 : (r == 3 ? 1 : 2)));
        OTT_UNROLL
        for (int m = 0; m < 4; ++m) {
            uint32_t uw, dw;
            if (j == 0) {
                uw = cuw[m]; dw = cdw[m];
            } else if (j > 0) {                                 // u shifted up by j bytes, d down by j bytes
                uw = ott_funnel_r(cuw[m], cuw[m < 3 ? m + 1 : 3], 8u * j);
                dw = ott_funnel_r(cdw[m > 0 ? m - 1 : 0], cdw[m], 32u - 8u * j);
            } else {
                uw = ott_funnel_r(cuw[m > 0 ? m - 1 : 0], cuw[m], 32u + 8 * j);
                dw = ott_funnel_r(cdw[m], cdw[m < 3 ? m + 1 : 3], (uint32_t)(-8 * j));
            }
            A[r][m] = ott_vabsdiff4(uw, dw);
            if (m == 1 || m == 2) P[r][m - 1] = yavg4(uw, dw);  // candidate predictions (u[x+j] + d[x-j]) >> 1, four per word
        }
    }
    const bool first_group = x8 == 0, last_group = 8 * x8 + 8 >= w;
    // candidate bytes transposed per sample (byte r of q = candidate of rank r, rank 4 in byte 0 of r4): the winner is then
    // one byte permute with the rank as the selector - no branches, no per-candidate selects
    uint32_t q[8], r4[8];
    OTT_UNROLL
    for (int wsel = 0; wsel < 2; ++wsel) {
        const uint32_t ab_lo = ott_prmt(P[0][wsel], P[1][wsel], 0x5140u), ab_hi = ott_prmt(P[0][wsel], P[1][wsel], 0x7362u);
        const uint32_t cd_lo = ott_prmt(P[2][wsel], P[3][wsel], 0x5140u), cd_hi = ott_prmt(P[2][wsel], P[3][wsel], 0x7362u);
        q[4 * wsel + 0] = ott_prmt(ab_lo, cd_lo, 0x5410u);
        q[4 * wsel + 1] = ott_prmt(ab_lo, cd_lo, 0x7632u);
        q[4 * wsel + 2] = ott_prmt(ab_hi, cd_hi, 0x5410u);
        q[4 * wsel + 3] = ott_prmt(ab_hi, cd_hi, 0x7632u);
        OTT_UNROLL
        for (int bsel = 0; bsel < 4; ++bsel) r4[4 * wsel + bsel] = ott_prmt(P[4][wsel], 0u, 0x4440u | (uint32_t)bsel);
    }
    uint32_t o16[4];
    OTT_UNROLL
    for (int i = 0; i < 8; ++i) {
        // bytes L-1, L, L+1 of A with L = 4 + i: one or two dot products against masks of 8s
        int key[5];
        OTT_UNROLL
        for (int r = 0; r < 5; ++r) {
            const uint32_t c0 = r == 0 ? 0xfffffff8u : (uint32_t)r;          // j = 0 carries the "- 1" of vf_yadif: 8*(s-1)
            uint32_t k;
            switch (i) {
            case 0: k = ott_dp4a_uu(A[r][1], 0x00000808u, ott_dp4a_uu(A[r][0], 0x08000000u, c0)); break;
            case 1: k = ott_dp4a_uu(A[r][1], 0x00080808u, c0); break;
            case 2: k = ott_dp4a_uu(A[r][1], 0x08080800u, c0); break;
            case 3: k = ott_dp4a_uu(A[r][2], 0x00000008u, ott_dp4a_uu(A[r][1], 0x08080000u, c0)); break;
            case 4: k = ott_dp4a_uu(A[r][2], 0x00000808u, ott_dp4a_uu(A[r][1], 0x08000000u, c0)); break;
            case 5: k = ott_dp4a_uu(A[r][2], 0x00080808u, c0); break;
            case 6: k = ott_dp4a_uu(A[r][2], 0x08080800u, c0); break;
            default: k = ott_dp4a_uu(A[r][3], 0x00000008u, ott_dp4a_uu(A[r][2], 0x08080000u, c0)); break;
            }
            key[r] = (int)k;
        }
        // vf_yadif only tries +-2 when +-1 improved the score; ties keep the earlier candidate (the rank in the low bits)
        const int k2 = ott_select(key[1] < key[0], key[2], 0x7fffffff);
        const int kl = ott_min(ott_min(key[0], key[1]), k2);
        const int best = ott_select(key[3] < kl, ott_min(key[3], key[4]), kl);
        // the search is skipped on the first and last three samples of a row (filter_edges): with w a multiple of 8 those
        // are samples 0-2 of the first group and 5-7 of the last one
        const bool in = i < 3 ? !first_group : (i >= 5 ? !last_group : true);
        const uint32_t sel = (uint32_t)ott_select(in, (best & 7) | 0x5550, 0x5550);      // bytes 1-3 of the result: zero (r4 byte 1)
        const uint32_t p = ott_prmt(q[i], r4[i], sel);
        if (i & 1) {
            const uint32_t pp = o16[i >> 1] + (p << 16);                                                         // (pred[i-1], pred[i]) + bias
            o16[i >> 1] = ott_vmin_s16x2(ott_vmax_s16x2(pp, lo[i >> 1]), hi[i >> 1]) - kBias;                   // 0..255 per lane
        } else {
            o16[i >> 1] = p + kBias;
        }
    }
    const uint32_t outw[2] = {ott_prmt(o16[0], o16[1], 0x6420u), ott_prmt(o16[2], o16[3], 0x6420u)};
    uint2 o;
    o.x = outw[0]; o.y = outw[1];
    *reinterpret_cast<uint2 *>(dst + i0) = o;
}

// Fast-path thread position: (x8, rp) over a grid of (w/8) x (row pairs of luma + row pairs of chroma) positions; a
// thread produces the kept row AND the interpolated row of its pair, so every warp of a block carries the same load
// (with one row per thread half of the warps would only copy and leave their slots idle until the block retires).
// Chroma row pairs hold the U samples in the left half and the V samples in the right half.
OTT_DEV void yadif_pair8(const YadifJob &j, size_t off, int w, int h, int x8, int rp)
{
    const int kept = 2 * rp + (j.parity & 1), other = 2 * rp + 1 - (j.parity & 1);      // rows with (y ^ parity) & 1 == 0 are kept
    if (kept < h) yadif_keep8(j, off, w, x8, kept);
    if (other < h) yadif_oct8(j, off, w, h, x8, other);
}

OTT_DEV int yadif_fast_rows(int h) { return ((h + 1) >> 1) + ((((h + 1) >> 1) + 1) >> 1); }   // row pairs: luma + chroma

OTT_DEV void yadif_position_fast(const YadifJob &j, int x8, int rp)
{
    const int w = j.width, h = j.height, cw = w >> 1, ch = (h + 1) >> 1;
    if (x8 >= (w >> 3)) return;
    const int lp = (h + 1) >> 1;
    if (rp < lp) { yadif_pair8(j, 0, w, h, x8, rp); return; }
    rp -= lp;
    if (rp >= ((ch + 1) >> 1)) return;
    const int co = cw >> 3;
    const size_t uo = (size_t)w * h, vo = uo + (size_t)cw * ch;
    if (x8 < co) yadif_pair8(j, uo, cw, ch, x8, rp);
    else yadif_pair8(j, vo, cw, ch, x8 - co, rp);
}

}  // namespace ottgpu
