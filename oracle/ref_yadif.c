/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the yadif half of the transvideo hot path.
 *
 * PARITY UNPINNED: the reference (cannonbeach/ott-packager) deinterlaces by pushing frames through libavfilter's
 * "yadif=0:-1:0" / "yadif=0:-1:1" graph:
 *   /root/reference/source/transvideo.c:2036-2052  filter strings (mode 0 = one frame per frame, parity auto,
 *                                                  deint all / deint flagged-only when fps_num is 60000|50000)
 *   /root/reference/source/transvideo.c:2100-2101  interlaced_frame / top_field_first taken from the message
 *   /root/reference/source/transvideo.c:2132-2141  av_buffersrc_add_frame_flags / av_buffersink_get_frame
 * libavfilter is NOT under /root/reference (un-vendored FFmpeg fork, setuptranscode.sh:97) and no libavfilter build
 * exists in this image, so this restatement of upstream libavfilter/vf_yadif.c (filter_line_c, filter_edges,
 * filter_slice) and yadif_common.c (ff_yadif_filter_frame, return_frame) cannot be checked against an executable
 * reference here.  It is pinned only by hand-derived known-answer tests and structural properties
 * (tests/test_oracle_yadif.py); SURVEY.md Appendix C is the summary it follows.
 *
 * Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() may use this file.
 */
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

static inline int iabs(int v) { return v < 0 ? -v : v; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int imin(int a, int b) { return a < b ? a : b; }

/*
 * One interpolated sample.  Pointers address the sample itself; pitches are in samples.
 *   up/dn  : offsets (in samples) to the rows above / below inside this field's neighbours (mrefs / prefs)
 *   p2/n2  : the two temporal neighbours of the missing field (prev2 / next2)
 *   search : 1 if the +-2 diagonal search may run (3 <= x < w-3)
 *   spat   : 1 if the 2-rows-away spatial interlacing check runs (mode bit 1 clear)
 */
#define YADIF_SAMPLE(T)                                                                                        \
    static inline int yadif_px_##T(const T *prev, const T *cur, const T *next, const T *p2, const T *n2,       \
                                   int up, int dn, int search, int spat)                                       \
    {                                                                                                          \
        int c = cur[up], e = cur[dn];                                                                          \
        int d = (p2[0] + n2[0]) >> 1;                                                                          \
        int td0 = iabs(p2[0] - n2[0]);                                                                         \
        int td1 = (iabs(prev[up] - c) + iabs(prev[dn] - e)) >> 1;                                              \
        int td2 = (iabs(next[up] - c) + iabs(next[dn] - e)) >> 1;                                              \
        int diff = imax(imax(td0 >> 1, td1), td2);                                                             \
        int pred = (c + e) >> 1;                                                                               \
        if (search) {                                                                                          \
            int score = iabs(cur[up - 1] - cur[dn - 1]) + iabs(c - e) + iabs(cur[up + 1] - cur[dn + 1]) - 1;   \
            int s;                                                                                             \
            s = iabs(cur[up - 2] - cur[dn]) + iabs(cur[up - 1] - cur[dn + 1]) + iabs(cur[up] - cur[dn + 2]);   \
            if (s < score) {                                                                                   \
                score = s; pred = (cur[up - 1] + cur[dn + 1]) >> 1;                                            \
                s = iabs(cur[up - 3] - cur[dn + 1]) + iabs(cur[up - 2] - cur[dn + 2]) +                        \
                    iabs(cur[up - 1] - cur[dn + 3]);                                                           \
                if (s < score) { score = s; pred = (cur[up - 2] + cur[dn + 2]) >> 1; }                         \
            }                                                                                                  \
            s = iabs(cur[up] - cur[dn - 2]) + iabs(cur[up + 1] - cur[dn - 1]) + iabs(cur[up + 2] - cur[dn]);   \
            if (s < score) {                                                                                   \
                score = s; pred = (cur[up + 1] + cur[dn - 1]) >> 1;                                            \
                s = iabs(cur[up + 1] - cur[dn - 3]) + iabs(cur[up + 2] - cur[dn - 2]) +                        \
                    iabs(cur[up + 3] - cur[dn - 1]);                                                           \
                if (s < score) { score = s; pred = (cur[up + 2] + cur[dn - 2]) >> 1; }                         \
            }                                                                                                  \
        }                                                                                                      \
        if (spat) {                                                                                            \
            int b = (p2[2 * up] + n2[2 * up]) >> 1;                                                            \
            int f = (p2[2 * dn] + n2[2 * dn]) >> 1;                                                            \
            int mx = imax(imax(d - e, d - c), imin(b - c, f - e));                                             \
            int mn = imin(imin(d - e, d - c), imax(b - c, f - e));                                             \
            diff = imax(imax(diff, mn), -mx);                                                                  \
        }                                                                                                      \
        if (pred > d + diff) pred = d + diff;                                                                  \
        else if (pred < d - diff) pred = d - diff;                                                             \
        return pred;                                                                                           \
    }

YADIF_SAMPLE(uint8_t)
YADIF_SAMPLE(uint16_t)

/*
 * One plane.  parity = which rows are interpolated: rows with ((y ^ parity) & 1) != 0.
 * tff as resolved by return_frame(): cur.interlaced ? cur.tff : 1.  First-field call => parity = tff ^ 1.
 */
#define YADIF_PLANE(T)                                                                                         \
    static void yadif_plane_##T(const T *prev, const T *cur, const T *next, T *dst, int pitch, int dpitch,     \
                                int w, int h, int parity, int tff, int row0, int row1)                         \
    {                                                                                                          \
        for (int y = row0; y < row1; y++) {                                                                    \
            const T *pr = prev + (size_t)y * pitch, *cu = cur + (size_t)y * pitch, *nx = next + (size_t)y * pitch; \
            T *o = dst + (size_t)y * dpitch;                                                                   \
            if (!((y ^ parity) & 1)) { memcpy(o, cu, sizeof(T) * (size_t)w); continue; }                       \
            int dn = y + 1 < h ? pitch : -pitch;                                                               \
            int up = y ? -pitch : pitch;                                                                       \
            int spat = !(y == 1 || y + 2 == h);                                                                \
            const T *p2 = (parity ^ tff) ? pr : cu;                                                            \
            const T *n2 = (parity ^ tff) ? cu : nx;                                                            \
            for (int x = 0; x < w; x++)                                                                        \
                o[x] = (T)yadif_px_##T(pr + x, cu + x, nx + x, p2 + x, n2 + x, up, dn, x >= 3 && x < w - 3, spat); \
        }                                                                                                      \
    }

YADIF_PLANE(uint8_t)
YADIF_PLANE(uint16_t)

/* bytes_per_sample 1 or 2; pitches in samples; rows [row0,row1) so callers can slice-thread like nb_threads=4. */
void ora_yadif_plane(const void *prev, const void *cur, const void *next, void *dst, int pitch, int dpitch,
                     int w, int h, int parity, int tff, int bytes_per_sample, int row0, int row1)
{
    if (bytes_per_sample == 1)
        yadif_plane_uint8_t((const uint8_t *)prev, (const uint8_t *)cur, (const uint8_t *)next, (uint8_t *)dst,
                            pitch, dpitch, w, h, parity, tff, row0, row1);
    else
        yadif_plane_uint16_t((const uint16_t *)prev, (const uint16_t *)cur, (const uint16_t *)next, (uint16_t *)dst,
                             pitch, dpitch, w, h, parity, tff, row0, row1);
}

/*
 * Whole planar 4:2:0 frame, contiguous (Y then U then V, strides w / ceil(w/2)) exactly like the packed buffers the
 * reference moves around (transvideo.c:1476-1502).  interlaced/tff are the message flags of `cur`.
 */
void ora_yadif_frame_420(const void *prev, const void *cur, const void *next, void *dst,
                         int w, int h, int interlaced, int tff_flag, int bytes_per_sample)
{
    const int tff = interlaced ? (tff_flag ? 1 : 0) : 1;
    const int parity = tff ^ 1;
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    const size_t bps = (size_t)bytes_per_sample;
    size_t off[3] = { 0, (size_t)w * h * bps, ((size_t)w * h + (size_t)cw * ch) * bps };
    int pw[3] = { w, cw, cw }, ph[3] = { h, ch, ch };
    for (int p = 0; p < 3; p++)
        ora_yadif_plane((const char *)prev + off[p], (const char *)cur + off[p], (const char *)next + off[p],
                        (char *)dst + off[p], pw[p], pw[p], pw[p], ph[p], parity, tff, bytes_per_sample, 0, ph[p]);
}
