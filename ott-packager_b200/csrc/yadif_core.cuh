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
// Fast path (8-bit, plane widths multiples of 4, 4-byte aligned frames): one thread = 4 adjacent samples.  Every
// neighbour row is fetched as aligned 32-bit words (3 words of `cur` above and below cover the +-3 sample search
// window of all four samples) instead of ~35 byte loads per sample.
OTT_DEV int ybyte(uint32_t w, int i) { return (int)((w >> (8 * i)) & 0xffu); }

OTT_DEV void yadif_quad8(const YadifJob &j, size_t off, int w, int h, int x4, int y)
{
    const uint32_t *prev = reinterpret_cast<const uint32_t *>(j.prev + off), *cur = reinterpret_cast<const uint32_t *>(j.cur + off);
    const uint32_t *next = reinterpret_cast<const uint32_t *>(j.next + off);
    uint32_t *dst = reinterpret_cast<uint32_t *>(j.dst + off);
    const int wq = w >> 2;                                 // words per row
    const int i0 = y * wq + x4;
    if (!((y ^ j.parity) & 1)) { dst[i0] = ott_ldg(cur + i0); return; }
    const int dn = y + 1 < h ? wq : -wq, up = y ? -wq : wq;
    const bool first_field = (j.parity ^ j.tff) != 0;
    const uint32_t *p2 = first_field ? prev : cur, *n2 = first_field ? cur : next;
    const bool spatial = !(y == 1 || y + 2 == h);
    const int xl = x4 > 0 ? -1 : 0, xr = x4 + 1 < wq ? 1 : 0;   // clamped neighbour words (values unused at the edges)
    const uint32_t cuw[3] = {ott_ldg(cur + i0 + up + xl), ott_ldg(cur + i0 + up), ott_ldg(cur + i0 + up + xr)};
    const uint32_t cdw[3] = {ott_ldg(cur + i0 + dn + xl), ott_ldg(cur + i0 + dn), ott_ldg(cur + i0 + dn + xr)};
    const uint32_t pu = ott_ldg(prev + i0 + up), pd = ott_ldg(prev + i0 + dn);
    const uint32_t nu = ott_ldg(next + i0 + up), nd = ott_ldg(next + i0 + dn);
    const uint32_t p20 = ott_ldg(p2 + i0), n20 = ott_ldg(n2 + i0);
    uint32_t p2u = 0, n2u = 0, p2d = 0, n2d = 0;
    if (spatial) {
        p2u = ott_ldg(p2 + i0 + 2 * up); n2u = ott_ldg(n2 + i0 + 2 * up);
        p2d = ott_ldg(p2 + i0 + 2 * dn); n2d = ott_ldg(n2 + i0 + 2 * dn);
    }
    int cu[12], cd[12];
    OTT_UNROLL
    for (int k = 0; k < 12; ++k) { cu[k] = ybyte(cuw[k >> 2], k & 3); cd[k] = ybyte(cdw[k >> 2], k & 3); }
    uint32_t out = 0;
    OTT_UNROLL
    for (int i = 0; i < 4; ++i) {
        const int x = 4 * x4 + i, q = i + 4;               // q indexes cu/cd
        const int c = cu[q], e = cd[q];
        const int a0 = ybyte(p20, i), b0 = ybyte(n20, i);
        const int d = (a0 + b0) >> 1;
        const int td0 = ott_abs(a0 - b0);
        const int td1 = (ott_abs(ybyte(pu, i) - c) + ott_abs(ybyte(pd, i) - e)) >> 1;
        const int td2 = (ott_abs(ybyte(nu, i) - c) + ott_abs(ybyte(nd, i) - e)) >> 1;
        int diff = ott_max(ott_max(td0 >> 1, td1), td2);
        int pred = (c + e) >> 1;
        if (x >= 3 && x < w - 3) {
            int score = ott_abs(cu[q - 1] - cd[q - 1]) + ott_abs(c - e) + ott_abs(cu[q + 1] - cd[q + 1]) - 1;
            int sc = ott_abs(cu[q - 2] - e) + ott_abs(cu[q - 1] - cd[q + 1]) + ott_abs(c - cd[q + 2]);                  // j = -1
            if (sc < score) {
                score = sc; pred = (cu[q - 1] + cd[q + 1]) >> 1;
                sc = ott_abs(cu[q - 3] - cd[q + 1]) + ott_abs(cu[q - 2] - cd[q + 2]) + ott_abs(cu[q - 1] - cd[q + 3]);  // j = -2
                if (sc < score) { score = sc; pred = (cu[q - 2] + cd[q + 2]) >> 1; }
            }
            sc = ott_abs(c - cd[q - 2]) + ott_abs(cu[q + 1] - cd[q - 1]) + ott_abs(cu[q + 2] - e);                     // j = +1
            if (sc < score) {
                score = sc; pred = (cu[q + 1] + cd[q - 1]) >> 1;
                sc = ott_abs(cu[q + 1] - cd[q - 3]) + ott_abs(cu[q + 2] - cd[q - 2]) + ott_abs(cu[q + 3] - cd[q - 1]);  // j = +2
                if (sc < score) { score = sc; pred = (cu[q + 2] + cd[q - 2]) >> 1; }
            }
        }
        if (spatial) {
            const int b = (ybyte(p2u, i) + ybyte(n2u, i)) >> 1, f = (ybyte(p2d, i) + ybyte(n2d, i)) >> 1;
            const int mx = ott_max(ott_max(d - e, d - c), ott_min(b - c, f - e));
            const int mn = ott_min(ott_min(d - e, d - c), ott_max(b - c, f - e));
            diff = ott_max(ott_max(diff, mn), -mx);
        }
        const int v = ott_min(ott_max(pred, d - diff), d + diff);
        out |= (uint32_t)v << (8 * i);
    }
    dst[i0] = out;
}

// Fast-path thread position: (x4, r) over a grid of (w/4) x (h + ch) words; rows >= h are chroma rows with the U words
// in the left half and the V words in the right half.
OTT_DEV void yadif_position_fast(const YadifJob &j, int x4, int r)
{
    const int w = j.width, h = j.height, cw = w >> 1, ch = (h + 1) >> 1;
    if (x4 >= (w >> 2)) return;
    if (r < h) { yadif_quad8(j, 0, w, h, x4, r); return; }
    r -= h;
    if (r >= ch) return;
    const int cq = cw >> 2;
    const size_t uo = (size_t)w * h, vo = uo + (size_t)cw * ch;
    if (x4 < cq) yadif_quad8(j, uo, cw, ch, x4, r);
    else yadif_quad8(j, vo, cw, ch, x4 - cq, r);
}

}  // namespace ottgpu
