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

}  // namespace ottgpu
