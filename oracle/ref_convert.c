/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the decode-side pixel-format converter.
 *
 * The reference converts whatever the decoder emits to packed 8-bit 4:2:0 before the prepare thread sees it:
 *   /root/reference/source/transvideo.c:2797-2808  sws_getContext(w,h,source_format, w,h,AV_PIX_FMT_YUV420P, SWS_BICUBIC)
 *                                                   + sws_scale()  ("422 to 420 conversion", "10 to 8 bit conversion")
 *   /root/reference/source/transvideo.c:2823-2840  three row-copy loops into the contiguous w*h*3/2 frame
 * The arithmetic lives in the un-vendored FFmpeg fork (libswscale); this file restates what libswscale does for the
 * same-size conversions yuv422p / yuv420p10le / yuv422p10le -> yuv420p under the reference's flags (target A) and under
 * SWS_BITEXACT (target B):
 *   - equal chroma subsampling (yuv420p10le): swscale_unscaled.c planarCopyWrapper, depth reduction with the 2x2
 *     ordered dither {{1,2},{3,0}}: dst = clip_u8((src + d[y&1][x&1]) >> 2)
 *   - 4:2:2 sources: the generic scaler.  Luma and the horizontal chroma direction are 1-tap (identity) filters, the
 *     vertical chroma direction is the 2:1 bicubic polyphase filter of ref_swscale.c; >8-bit sources switch on the
 *     8x8 Bayer dither (ff_dither_8x8_128, row = output row & 7, column = (x + offset) & 7 with offset 3 for V).
 *
 * PINNING: tests/test_oracle_convert.py checks this byte-for-byte against the real libswscale 8.0.1 bundled in the
 * image (oracle/swscale_so.py) and against fixtures it produced (tests/golden/convert_small.npz).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORA_PIX_YUV420P   0
#define ORA_PIX_YUV422P   1
#define ORA_PIX_YUV420P10 2
#define ORA_PIX_YUV422P10 3

typedef struct ora_filter ora_filter;
ora_filter *ora_filter_make(int src_n, int dst_n, int kernel, int align, int one, int zero_tail);
void ora_filter_free(ora_filter *f);
int ora_filter_taps(const ora_filter *f);
const int32_t *ora_filter_pos(const ora_filter *f);
const int16_t *ora_filter_coef(const ora_filter *f);

static const uint8_t bayer_8x8_128[8][8] = {
    { 36, 68,  60, 92,  34, 66,  58, 90 }, { 100,  4, 124, 28,  98,  2, 122, 26 },
    { 52, 84,  44, 76,  50, 82,  42, 74 }, { 116, 20, 108, 12, 114, 18, 106, 10 },
    { 32, 64,  56, 88,  38, 70,  62, 94 }, {  96,  0, 120, 24, 102,  6, 126, 30 },
    { 48, 80,  40, 72,  54, 86,  46, 78 }, { 112, 16, 104,  8, 118, 22, 110, 14 },
};
static const uint8_t dither_2x2_4[2][2] = { { 1, 2 }, { 3, 0 } };

static inline uint8_t clip_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

static void copy_plane(const uint8_t *src, int sstride, uint8_t *dst, int w, int h)
{
    for (int y = 0; y < h; y++) memcpy(dst + (size_t)y * w, src + (size_t)y * sstride, (size_t)w);
}

/* planarCopyWrapper, 10 -> 8 bit */
static void reduce_plane_2x2(const uint8_t *src, int sstride, uint8_t *dst, int w, int h)
{
    for (int y = 0; y < h; y++) {
        const uint16_t *s = (const uint16_t *)(src + (size_t)y * sstride);
        for (int x = 0; x < w; x++) dst[(size_t)y * w + x] = clip_u8((s[x] + dither_2x2_4[y & 1][x & 1]) >> 2);
    }
}

/* generic path, 1-tap vertical filter on a >8-bit source: hScale16To15 (<<5) then yuv2plane1_8 with the Bayer row */
static void reduce_plane_bayer(const uint8_t *src, int sstride, uint8_t *dst, int w, int h)
{
    for (int y = 0; y < h; y++) {
        const uint16_t *s = (const uint16_t *)(src + (size_t)y * sstride);
        for (int x = 0; x < w; x++) dst[(size_t)y * w + x] = clip_u8(((s[x] << 5) + bayer_8x8_128[y & 7][x & 7]) >> 7);
    }
}

/*
 * Chroma plane of a 4:2:2 source (cw x h) -> cw x ch, vertical polyphase only.
 * bps 1: intermediates src<<7, constant dither 64.  bps 2: intermediates src<<5, Bayer dither with column offset.
 * target 0 (A): x86 yuv2yuvX on all rows but the ones produced once dstY >= dstH-2 (exact_from), which use the C code.
 */
static void chroma_422_to_420(const uint8_t *src, int sstride, int bps, uint8_t *dst, int cw, int h, int ch,
                              int target, int exact_from, int offset)
{
    ora_filter *vf = ora_filter_make(h, ch, 0 /* bicubic */, target ? 1 : 2, 1 << 12, target);
    const int vs = ora_filter_taps(vf);
    const int32_t *pos = ora_filter_pos(vf);
    const int16_t *coef = ora_filter_coef(vf);
    for (int y = 0; y < ch; y++) {
        const int16_t *c = coef + (size_t)y * vs;
        uint8_t *o = dst + (size_t)y * cw;
        for (int x = 0; x < cw; x++) {
            const int d = bps == 2 ? bayer_8x8_128[y & 7][(x + offset) & 7] : 64;
            int t[64];
            for (int j = 0; j < vs; j++) {
                const uint8_t *row = src + (size_t)(pos[y] + j) * sstride;
                t[j] = bps == 2 ? ((const uint16_t *)row)[x] << 5 : row[x] << 7;
            }
            if (vs == 1) {
                o[x] = clip_u8((t[0] + d) >> 7);
            } else if (target == 1 || y >= exact_from) {
                int val = d << 12;
                for (int j = 0; j < vs; j++) val += t[j] * c[j];
                o[x] = clip_u8(val >> 19);
            } else {
                int16_t acc = (int16_t)((d + 8 * (vs - 1)) >> 4);
                for (int j = 0; j < vs; j++) acc = (int16_t)(acc + (int16_t)((t[j] * c[j]) >> 16));
                o[x] = clip_u8(acc >> 3);
            }
        }
    }
    ora_filter_free(vf);
}

/*
 * planes/strides (bytes) as the decoder hands them over; dst = contiguous I420 frame (w*h + 2*cw*ch bytes).
 * Returns 0, or -1 for an unknown format.
 */
int ora_convert_to_i420(const void *const planes[3], const int strides[3], int fmt, int w, int h, uint8_t *dst, int target)
{
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    uint8_t *dy = dst, *du = dy + (size_t)w * h, *dv = du + (size_t)cw * ch;
    const uint8_t *sy = (const uint8_t *)planes[0], *su = (const uint8_t *)planes[1], *sv = (const uint8_t *)planes[2];
    int uv_exact = (h - 2 + 1) / 2;            /* chroma row uv is produced while dstY == 2*uv */
    if (uv_exact < 0) uv_exact = 0;
    switch (fmt) {
    case ORA_PIX_YUV420P:
        copy_plane(sy, strides[0], dy, w, h);
        copy_plane(su, strides[1], du, cw, ch);
        copy_plane(sv, strides[2], dv, cw, ch);
        return 0;
    case ORA_PIX_YUV420P10:
        reduce_plane_2x2(sy, strides[0], dy, w, h);
        reduce_plane_2x2(su, strides[1], du, cw, ch);
        reduce_plane_2x2(sv, strides[2], dv, cw, ch);
        return 0;
    case ORA_PIX_YUV422P:
        copy_plane(sy, strides[0], dy, w, h);
        chroma_422_to_420(su, strides[1], 1, du, cw, h, ch, target, uv_exact, 0);
        chroma_422_to_420(sv, strides[2], 1, dv, cw, h, ch, target, uv_exact, 3);
        return 0;
    case ORA_PIX_YUV422P10:
        reduce_plane_bayer(sy, strides[0], dy, w, h);
        chroma_422_to_420(su, strides[1], 2, du, cw, h, ch, target, uv_exact, 0);
        chroma_422_to_420(sv, strides[2], 2, dv, cw, h, ch, target, uv_exact, 3);
        return 0;
    }
    return -1;
}
