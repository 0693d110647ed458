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
 * Like libavcodec's MJPEG encoder (default huffman=optimal) the picture is coded with its own optimal Huffman tables, built
 * from the symbol statistics by the procedure of T.81 Annex K.2 (IJG jpeg_gen_optimal_table), so this restatement makes
 * two passes: quantise all blocks and count symbols, then write the tables and the entropy-coded data.
 * Public algorithms followed: IJG jfdctint.c (integer "islow" DCT, results scaled by 8), libavcodec's quantiser constants
 * (measured against the real library: every DC and 99.8 % of the AC levels come out identical),
 * ITU-T T.81 Annex C / F / K (Huffman code assignment, entropy coding, standard tables).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

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

/* quantised block in zigzag order */
static void quantise_block(const uint8_t *plane, int pitch, int pw, int ph, int bx, int by, const int *q, int *qz)
{
    int blk[64];
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
}

/* one block's symbols: either counted (w == NULL) or written */
static void code_block(writer *w, const int *qz, int *pred, long *dc_freq, long *ac_freq, const htab *dc, const htab *ac)
{
    int diff = qz[0] - *pred, s = nbits_of(diff);
    *pred = qz[0];
    if (!w) dc_freq[s]++;
    else {
        wbits(w, dc->code[s], dc->len[s]);
        if (s) wbits(w, (unsigned)(diff < 0 ? diff - 1 : diff) & ((1u << s) - 1u), s);
    }
    int run = 0;
    for (int k = 1; k < 64; ++k) {
        if (!qz[k]) { ++run; continue; }
        for (; run > 15; run -= 16) {
            if (!w) ac_freq[0xF0]++;
            else wbits(w, ac->code[0xF0], ac->len[0xF0]);
        }
        s = nbits_of(qz[k]);
        if (!w) ac_freq[(run << 4) | s]++;
        else {
            wbits(w, ac->code[(run << 4) | s], ac->len[(run << 4) | s]);
            wbits(w, (unsigned)(qz[k] < 0 ? qz[k] - 1 : qz[k]) & ((1u << s) - 1u), s);
        }
        run = 0;
    }
    if (run) {
        if (!w) ac_freq[0]++;
        else wbits(w, ac->code[0], ac->len[0]);
    }
}

/* T.81 K.2 / IJG jpeg_gen_optimal_table: optimal code lengths (<= 16) for the counted symbols; one code point is reserved
 * so that no symbol gets the all-ones code */
static void optimal_table(const long *count, unsigned char *bits /*16*/, unsigned char *vals, int *nvals)
{
    long freq[257];
    int size[257], other[257];
    int per_len[33];
    memset(per_len, 0, sizeof(per_len));
    for (int i = 0; i < 256; ++i) { freq[i] = count[i]; size[i] = 0; other[i] = -1; }
    freq[256] = 1; size[256] = 0; other[256] = -1;
    for (;;) {
        int lo1 = -1, lo2 = -1;
        long best = 1000000000L;
        for (int i = 0; i <= 256; ++i) if (freq[i] && freq[i] <= best) { best = freq[i]; lo1 = i; }
        best = 1000000000L;
        for (int i = 0; i <= 256; ++i) if (freq[i] && freq[i] <= best && i != lo1) { best = freq[i]; lo2 = i; }
        if (lo2 < 0) break;
        freq[lo1] += freq[lo2];
        freq[lo2] = 0;
        size[lo1]++;
        while (other[lo1] >= 0) { lo1 = other[lo1]; size[lo1]++; }
        other[lo1] = lo2;
        size[lo2]++;
        while (other[lo2] >= 0) { lo2 = other[lo2]; size[lo2]++; }
    }
    for (int i = 0; i <= 256; ++i) if (size[i]) per_len[size[i] > 32 ? 32 : size[i]]++;
    for (int i = 32; i > 16; --i)
        while (per_len[i] > 0) {
            int j = i - 2;
            while (per_len[j] == 0) --j;
            per_len[i] -= 2; per_len[i - 1]++; per_len[j + 1] += 2; per_len[j]--;
        }
    int longest = 16;
    while (longest > 0 && per_len[longest] == 0) --longest;
    if (longest > 0) per_len[longest]--;
    for (int l = 1; l <= 16; ++l) bits[l - 1] = (unsigned char)per_len[l];
    int p = 0;
    for (int l = 1; l <= 32; ++l)
        for (int sym = 0; sym < 256; ++sym) if (size[sym] == l) vals[p++] = (unsigned char)sym;
    *nvals = p;
}

/* returns the file size (may exceed cap: then nothing beyond cap was written) */
size_t ora_jpeg_encode(const uint8_t *y, int ypitch, const uint8_t *u, const uint8_t *v, int cpitch, int width, int height, int qscale,
                       uint8_t *out, size_t cap)
{
    writer w = {out, 0, cap, 0, 0};
    int q[64];
    for (int i = 0; i < 64; ++i) { q[i] = (mpeg1_intra[i] * qscale) >> 3; if (q[i] < 1) q[i] = 1; if (q[i] > 255) q[i] = 255; }
    q[0] = 8;
    const int mx = (width + 15) / 16, my = (height + 15) / 16, cw = (width + 1) / 2, ch = (height + 1) / 2;
    const int nblocks = 6 * mx * my;
    int *coef = (int *)malloc((size_t)nblocks * 64 * sizeof(int));
    if (!coef) return 0;
    /* pass 1: all blocks in scan order (MCU by MCU: four luma blocks, Cb, Cr), symbol statistics */
    for (int m = 0, b = 0; m < mx * my; ++m) {
        const int ax = m % mx, ay = m / mx;
        for (int k = 0; k < 4; ++k, ++b) quantise_block(y, ypitch, width, height, 2 * ax + (k & 1), 2 * ay + (k >> 1), q, coef + 64 * b);
        quantise_block(u, cpitch, cw, ch, ax, ay, q, coef + 64 * b++);
        quantise_block(v, cpitch, cw, ch, ax, ay, q, coef + 64 * b++);
    }
    long dcl_f[256], dcc_f[256], acl_f[256], acc_f[256];
    memset(dcl_f, 0, sizeof(dcl_f)); memset(dcc_f, 0, sizeof(dcc_f)); memset(acl_f, 0, sizeof(acl_f)); memset(acc_f, 0, sizeof(acc_f));
    int py = 0, pu = 0, pv = 0;
    for (int m = 0; m < mx * my; ++m) {
        for (int k = 0; k < 4; ++k) code_block(NULL, coef + 64 * (6 * m + k), &py, dcl_f, acl_f, NULL, NULL);
        code_block(NULL, coef + 64 * (6 * m + 4), &pu, dcc_f, acc_f, NULL, NULL);
        code_block(NULL, coef + 64 * (6 * m + 5), &pv, dcc_f, acc_f, NULL, NULL);
    }
    unsigned char bits[4][16], vals[4][256];
    int nvals[4];
    optimal_table(dcl_f, bits[0], vals[0], &nvals[0]); optimal_table(acl_f, bits[1], vals[1], &nvals[1]);
    optimal_table(dcc_f, bits[2], vals[2], &nvals[2]); optimal_table(acc_f, bits[3], vals[3], &nvals[3]);
    htab dcl, dcc, acl, acc;
    mktab(bits[0], vals[0], &dcl); mktab(bits[1], vals[1], &acl); mktab(bits[2], vals[2], &dcc); mktab(bits[3], vals[3], &acc);
    /* header */
    w16(&w, 0xFFD8);
    w16(&w, 0xFFFE); w16(&w, 2 + 9); for (const char *c = "libottgpu"; *c; ++c) wbyte(&w, *c);
    w16(&w, 0xFFDB); w16(&w, 67); wbyte(&w, 0); for (int k = 0; k < 64; ++k) wbyte(&w, q[zz[k]]);
    /* segment order and the single four-table DHT segment as libavcodec writes them: DQT, DHT, SOF0, SOS */
    int dht_len = 2;
    for (int t = 0; t < 4; ++t) dht_len += 17 + nvals[t];
    w16(&w, 0xFFC4); w16(&w, dht_len);
    const int ids[4] = {0x00, 0x01, 0x10, 0x11};
    const int order[4] = {0, 2, 1, 3};                         /* bits[]/vals[] are DC luma, AC luma, DC chroma, AC chroma */
    for (int t = 0; t < 4; ++t) {
        const int k = order[t];
        wbyte(&w, ids[t]);
        for (int i = 0; i < 16; ++i) wbyte(&w, bits[k][i]);
        for (int i = 0; i < nvals[k]; ++i) wbyte(&w, vals[k][i]);
    }
    w16(&w, 0xFFC0); w16(&w, 17); wbyte(&w, 8); w16(&w, height); w16(&w, width); wbyte(&w, 3);
    wbyte(&w, 1); wbyte(&w, 0x22); wbyte(&w, 0); wbyte(&w, 2); wbyte(&w, 0x11); wbyte(&w, 0); wbyte(&w, 3); wbyte(&w, 0x11); wbyte(&w, 0);
    w16(&w, 0xFFDA); w16(&w, 12); wbyte(&w, 3); wbyte(&w, 1); wbyte(&w, 0x00); wbyte(&w, 2); wbyte(&w, 0x11); wbyte(&w, 3); wbyte(&w, 0x11);
    wbyte(&w, 0); wbyte(&w, 63); wbyte(&w, 0);
    /* pass 2: the entropy-coded segment */
    py = pu = pv = 0;
    for (int m = 0; m < mx * my; ++m) {
        for (int k = 0; k < 4; ++k) code_block(&w, coef + 64 * (6 * m + k), &py, NULL, NULL, &dcl, &acl);
        code_block(&w, coef + 64 * (6 * m + 4), &pu, NULL, NULL, &dcc, &acc);
        code_block(&w, coef + 64 * (6 * m + 5), &pv, NULL, NULL, &dcc, &acc);
    }
    wflush(&w);
    w16(&w, 0xFFD9);
    free(coef);
    return w.n;
}
