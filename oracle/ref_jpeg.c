/* TEST INFRASTRUCTURE ONLY -- never linked into libottgpu.
 *
 * CPU restatement of the thumbnail JPEG encoder (the reference's save_frame_as_jpeg,
 * /root/reference/source/transvideo.c:164-198, hands one 176x144 YUVJ420P frame to libavcodec's MJPEG encoder).
 * The reference's encoder arithmetic lives in its un-vendored FFmpeg fork and cannot be restated bit-exactly (SIMD DCT,
 * rate control, optimal Huffman tables): PARITY UNPINNED against libavcodec's BYTES.  What IS pinned to the real library of
 * this image (oracle/avcodec_so.py, tests/test_parity_jpeg.py): the file structure and the quantisation table
 * (q[i] = (ff_mpeg1_default_intra_matrix[i] * qscale) >> 3, q[0] = 8, one table for all components) and the decoded
 * picture quality.  This file is written the way a sequential encoder is written (a bit writer that stuffs as it goes,
 * blocks visited MCU by MCU with running DC predictors): it shares no code with the parallel GPU formulation in
 * ott-packager_b200/csrc/jpeg_core.cuh and must produce the same bytes.
 * Public algorithms followed: IJG jfdctint.c (integer "islow" DCT, results scaled by 8), libavcodec's quantiser constants
 * (measured against the real library: every DC and 99.8 % of the AC levels come out identical),
 * ITU-T T.81 Annex C / F / K (Huffman code assignment, entropy coding, standard tables).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "jpeg_std_tables.h"

static const unsigned char mpeg1_intra[64] = {
    8,  16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34, 34, 38, 22, 22, 26, 27, 29, 34, 37, 40,
    22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32, 35, 40, 48, 58, 26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83};

static const unsigned char zz[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                     41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                     30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

typedef struct { uint8_t *p; size_t n, cap; uint32_t acc; int nbits; } writer;

static void wbyte(writer *w, int b) { if (w->n < w->cap) w->p[w->n] = (uint8_t)b; w->n++; }
static void w16(writer *w, int v) { wbyte(w, v >> 8); wbyte(w, v & 255); }
static void wbits(writer *w, unsigned code, int len)
{
    for (int i = len - 1; i >= 0; --i) {
        w->acc = (w->acc << 1) | ((code >> i) & 1u);
        if (++w->nbits == 8) {
            wbyte(w, (int)(w->acc & 255u));
            if ((w->acc & 255u) == 255u) wbyte(w, 0);
            w->acc = 0; w->nbits = 0;
        }
    }
}
static void wflush(writer *w) { while (w->nbits) wbits(w, 1, 1); }

/* jfdctint.c */
#define DESCALE(x, n) (((x) + (1 << ((n) - 1))) >> (n))
static void fdct_islow(int *data)
{
    int *d = data;
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < 8; ++i) {
            int *e = pass == 0 ? data + 8 * i : data + i;
            const int st = pass == 0 ? 1 : 8;
            int t0 = e[0] + e[7 * st], t7 = e[0] - e[7 * st], t1 = e[st] + e[6 * st], t6 = e[st] - e[6 * st];
            int t2 = e[2 * st] + e[5 * st], t5 = e[2 * st] - e[5 * st], t3 = e[3 * st] + e[4 * st], t4 = e[3 * st] - e[4 * st];
            int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
            const int sh = pass == 0 ? 13 - 2 : 13 + 2;
            if (pass == 0) { e[0] = (t10 + t11) * 4; e[4 * st] = (t10 - t11) * 4; }
            else { e[0] = DESCALE(t10 + t11, 2); e[4 * st] = DESCALE(t10 - t11, 2); }
            int z1 = (t12 + t13) * 4433;
            e[2 * st] = DESCALE(z1 + t13 * 6270, sh);
            e[6 * st] = DESCALE(z1 + t12 * -15137, sh);
            z1 = t4 + t7;
            int z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7, z5 = (z3 + z4) * 9633;
            t4 *= 2446; t5 *= 16819; t6 *= 25172; t7 *= 12299;
            z1 *= -7373; z2 *= -20995; z3 *= -16069; z4 *= -3196;
            z3 += z5; z4 += z5;
            e[7 * st] = DESCALE(t4 + z1 + z3, sh);
            e[5 * st] = DESCALE(t5 + z2 + z4, sh);
            e[3 * st] = DESCALE(t6 + z2 + z3, sh);
            e[st] = DESCALE(t7 + z1 + z4, sh);
        }
    }
    (void)d;
}

typedef struct { unsigned code[256]; int len[256]; } htab;
static void mktab(const unsigned char *bits, const unsigned char *vals, htab *t)
{
    memset(t, 0, sizeof(*t));
    unsigned code = 0;
    int k = 0;
    for (int l = 1; l <= 16; ++l) {
        for (int i = 0; i < bits[l - 1]; ++i) { t->code[vals[k]] = code++; t->len[vals[k]] = l; ++k; }
        code <<= 1;
    }
}
static int nbits_of(int v) { int a = v < 0 ? -v : v, n = 0; while (a) { ++n; a >>= 1; } return n; }

static void encode_block(writer *w, const uint8_t *plane, int pitch, int pw, int ph, int bx, int by, const int *q, int *pred,
                         const htab *dc, const htab *ac)
{
    int blk[64], qz[64];
    for (int r = 0; r < 8; ++r)
        for (int c = 0; c < 8; ++c) {
            int y = 8 * by + r, x = 8 * bx + c;
            if (y > ph - 1) y = ph - 1;
            if (x > pw - 1) x = pw - 1;
            blk[8 * r + c] = (int)plane[(size_t)y * pitch + x] - 128;
        }
    fdct_islow(blk);
    /* libavcodec's quantiser for MJPEG (mpegvideo_enc.c: convert_matrix + the 16-bit dct_quantize): DC scale 8 with floor
     * rounding, AC with qmat16 = 65536 / (8 q) and the 3/8 intra bias carried as bias16 = round(96 * 256 / qmat16) */
    for (int k = 0; k < 64; ++k) {
        int v = blk[zz[k]];
        if (k == 0) {
            v = (v + 32) >> 6;
            if (v > 1023) v = 1023;
            if (v < -1023) v = -1023;
            qz[k] = v;
            continue;
        }
        unsigned qmat16 = (2u << 16) / (16u * (unsigned)q[zz[k]]);
        if (qmat16 == 0 || qmat16 == 128 * 256) qmat16 = 128 * 256 - 1;
        const unsigned bias16 = (96u * 256u + qmat16 / 2) / qmat16;
        const int neg = v < 0;
        unsigned a = (unsigned)(neg ? -v : v);
        a = ((a + bias16) * qmat16) >> 16;
        if (a > 1023) a = 1023;
        qz[k] = neg ? -(int)a : (int)a;
    }
    int diff = qz[0] - *pred, s = nbits_of(diff);
    *pred = qz[0];
    wbits(w, dc->code[s], dc->len[s]);
    if (s) wbits(w, (unsigned)(diff < 0 ? diff - 1 : diff) & ((1u << s) - 1u), s);
    int run = 0;
    for (int k = 1; k < 64; ++k) {
        if (!qz[k]) { ++run; continue; }
        for (; run > 15; run -= 16) wbits(w, ac->code[0xF0], ac->len[0xF0]);
        s = nbits_of(qz[k]);
        wbits(w, ac->code[(run << 4) | s], ac->len[(run << 4) | s]);
        wbits(w, (unsigned)(qz[k] < 0 ? qz[k] - 1 : qz[k]) & ((1u << s) - 1u), s);
        run = 0;
    }
    if (run) wbits(w, ac->code[0], ac->len[0]);
}

/* returns the file size (may exceed cap: then nothing beyond cap was written) */
size_t ora_jpeg_encode(const uint8_t *y, int ypitch, const uint8_t *u, const uint8_t *v, int cpitch, int width, int height, int qscale,
                       uint8_t *out, size_t cap)
{
    writer w = {out, 0, cap, 0, 0};
    int q[64];
    for (int i = 0; i < 64; ++i) { q[i] = (mpeg1_intra[i] * qscale) >> 3; if (q[i] < 1) q[i] = 1; if (q[i] > 255) q[i] = 255; }
    q[0] = 8;
    htab dcl, dcc, acl, acc;
    mktab(ref_bits_dc_luma, ref_val_dc_luma, &dcl); mktab(ref_bits_dc_chroma, ref_val_dc_chroma, &dcc);
    mktab(ref_bits_ac_luma, ref_val_ac_luma, &acl); mktab(ref_bits_ac_chroma, ref_val_ac_chroma, &acc);
    w16(&w, 0xFFD8);
    w16(&w, 0xFFFE); w16(&w, 2 + 9); for (const char *c = "libottgpu"; *c; ++c) wbyte(&w, *c);
    w16(&w, 0xFFDB); w16(&w, 67); wbyte(&w, 0); for (int k = 0; k < 64; ++k) wbyte(&w, q[zz[k]]);
    w16(&w, 0xFFC0); w16(&w, 17); wbyte(&w, 8); w16(&w, height); w16(&w, width); wbyte(&w, 3);
    wbyte(&w, 1); wbyte(&w, 0x22); wbyte(&w, 0); wbyte(&w, 2); wbyte(&w, 0x11); wbyte(&w, 0); wbyte(&w, 3); wbyte(&w, 0x11); wbyte(&w, 0);
    const struct { int id; const unsigned char *bits, *vals; int n; } tabs[4] = {
        {0x00, ref_bits_dc_luma, ref_val_dc_luma, 12}, {0x10, ref_bits_ac_luma, ref_val_ac_luma, 162},
        {0x01, ref_bits_dc_chroma, ref_val_dc_chroma, 12}, {0x11, ref_bits_ac_chroma, ref_val_ac_chroma, 162}};
    for (int t = 0; t < 4; ++t) {
        w16(&w, 0xFFC4); w16(&w, 19 + tabs[t].n); wbyte(&w, tabs[t].id);
        for (int i = 0; i < 16; ++i) wbyte(&w, tabs[t].bits[i]);
        for (int i = 0; i < tabs[t].n; ++i) wbyte(&w, tabs[t].vals[i]);
    }
    w16(&w, 0xFFDA); w16(&w, 12); wbyte(&w, 3); wbyte(&w, 1); wbyte(&w, 0x00); wbyte(&w, 2); wbyte(&w, 0x11); wbyte(&w, 3); wbyte(&w, 0x11);
    wbyte(&w, 0); wbyte(&w, 63); wbyte(&w, 0);
    const int mx = (width + 15) / 16, my = (height + 15) / 16, cw = (width + 1) / 2, ch = (height + 1) / 2;
    int py = 0, pu = 0, pv = 0;
    for (int m = 0; m < mx * my; ++m) {
        const int ax = m % mx, ay = m / mx;
        for (int k = 0; k < 4; ++k) encode_block(&w, y, ypitch, width, height, 2 * ax + (k & 1), 2 * ay + (k >> 1), q, &py, &dcl, &acl);
        encode_block(&w, u, cpitch, cw, ch, ax, ay, q, &pu, &dcc, &acc);
        encode_block(&w, v, cpitch, cw, ch, ax, ay, q, &pv, &dcc, &acc);
    }
    wflush(&w);
    w16(&w, 0xFFD9);
    return w.n;
}
