/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the swscale half of the transvideo hot path.
 *
 * The reference (cannonbeach/ott-packager) does its scaling by calling libswscale:
 *   /root/reference/source/transvideo.c:1563-1565  sws_getContext(w,h,YUV420P, ow,oh,YUV420P, SWS_BICUBIC,...)
 *   /root/reference/source/transvideo.c:1576-1581  sws_scale(ctx, planes, strides, 0, h, out, outstrides)
 *   /root/reference/source/transvideo.c:2155-2174  same call for the 176x144 thumbnail
 * The arithmetic itself is NOT in /root/reference: it lives in the un-vendored, un-pinned FFmpeg fork
 * (setuptranscode.sh:97 `git clone https://github.com/cannonbeach/FFmpeg.git ./cbffmpeg`).  This file restates the
 * published legacy-swscale algorithm (libswscale/utils.c initFilter, swscale.c hScale8To15_c/hScale16To15_c,
 * output.c yuv2planeX_8_c / yuv2plane1_8_c / yuv2p010lX_c, x86/yuv2yuvX.asm) as summarised in SURVEY.md Appendix A.
 *
 * PINNING: tests/test_oracle_swscale.py checks this restatement byte-for-byte against the real libswscale 8.0.1
 * bundled in this image (oracle/swscale_so.py) and against the fixtures it produced (tests/golden/).
 *
 * Two parity targets:
 *   A ("x86 default")  = flags SWS_BICUBIC / SWS_LANCZOS alone on any x86 with MMX+: filterAlign H=4,V=2, padded taps
 *                        kept, 8-bit vertical pass with per-tap (a*b)>>16 truncation (pmulhw), C path on last rows.
 *   B ("bitexact")     = SWS_BITEXACT / SWS_ACCURATE_RND / C-only: filterAlign 1, exact vertical pass everywhere.
 *
 * Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() may use this file.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

#if defined(__GNUC__)
#pragma GCC optimize("fp-contract=off") /* lanczos taps must round exactly like a plain x86-64 libswscale build */
#endif

#define ORA_BICUBIC 0
#define ORA_LANCZOS 1

typedef struct {
    int     taps;     /* filter size after reduction + alignment */
    int     n;        /* number of output samples               */
    int32_t *pos;     /* [n]       first source sample          */
    int16_t *coef;    /* [n*taps]                               */
} ora_filter;

static int64_t i64abs(int64_t v) { return v < 0 ? -v : v; }
static int ilog2(unsigned v) { int r = 0; while (v >>= 1) r++; return r; }
static int64_t rounded_div(int64_t a, int64_t b) { return (a >= 0 ? a + (b >> 1) : a - (b >> 1)) / b; }

void ora_filter_free(ora_filter *f)
{
    if (!f) return;
    free(f->pos); free(f->coef); free(f);
}

/* utils.c initFilter(), bicubic (B=0,C=0.6) and lanczos (p=3) branches, srcPos=dstPos=128 (centre sampling). */
ora_filter *ora_filter_make(int src_n, int dst_n, int kernel, int align, int one, int zero_tail)
{
    const int xinc = (int)((((int64_t)src_n << 16) + (dst_n >> 1)) / dst_n);
    int ratio_log = ilog2((unsigned)(src_n / dst_n > 0 ? src_n / dst_n : 1));
    if (src_n / dst_n == 0) ratio_log = 0;
    if (ratio_log > 8) ratio_log = 8;
    const int64_t fone = 1LL << (54 - ratio_log);
    int fsize, i, j;
    int64_t *raw;
    int32_t *pos = (int32_t *)malloc(sizeof(int32_t) * (size_t)(dst_n + 3));

    int d10 = xinc - 0x10000; if (d10 < 0) d10 = -d10;
    if (d10 < 10) {
        fsize = 1;
        raw = (int64_t *)calloc((size_t)dst_n, sizeof(int64_t));
        for (i = 0; i < dst_n; i++) { raw[i] = fone; pos[i] = i; }
    } else {
        const int size_factor = kernel == ORA_LANCZOS ? 6 : 4;
        int64_t x_dst_in_src;
        if (xinc <= (1 << 16)) fsize = 1 + size_factor;
        else fsize = 1 + (int)(((int64_t)size_factor * src_n + dst_n - 1) / dst_n);
        if (fsize > src_n - 2) fsize = src_n - 2;
        if (fsize < 1) fsize = 1;
        raw = (int64_t *)malloc(sizeof(int64_t) * (size_t)dst_n * (size_t)fsize);
        x_dst_in_src = ((128 * (int64_t)xinc) >> 7) - ((128 * 0x10000LL) >> 7);
        for (i = 0; i < dst_n; i++) {
            int xx = (int)((x_dst_in_src - (int64_t)(fsize - 2) * (1LL << 16)) / (1 << 17));
            pos[i] = xx;
            for (j = 0; j < fsize; j++) {
                int64_t d = i64abs(((int64_t)xx * (1 << 17)) - x_dst_in_src) << 13;
                int64_t coeff;
                double floatd;
                if (xinc > (1 << 16)) d = d * dst_n / src_n;
                floatd = d * (1.0 / (1 << 30));
                if (kernel == ORA_BICUBIC) {
                    const int64_t B = 0;
                    const int64_t C = (int64_t)(0.6 * (1 << 24));
                    if (d >= (1LL << 31)) {
                        coeff = 0;
                    } else {
                        int64_t dd = (d * d) >> 30;
                        int64_t ddd = (dd * d) >> 30;
                        if (d < (1LL << 30))
                            coeff = (12 * (1 << 24) - 9 * B - 6 * C) * ddd +
                                    (-18 * (1 << 24) + 12 * B + 6 * C) * dd +
                                    (6 * (1 << 24) - 2 * B) * (1LL << 30);
                        else
                            coeff = (-B - 6 * C) * ddd + (6 * B + 30 * C) * dd +
                                    (-12 * B - 48 * C) * d + (8 * B + 24 * C) * (1LL << 30);
                    }
                    coeff /= (1LL << 54) / fone;
                } else {
                    const double p = 3.0;
                    coeff = (int64_t)((d ? sin(floatd * M_PI) * sin(floatd * M_PI / p) /
                                           (floatd * floatd * M_PI * M_PI / p)
                                         : 1.0) * fone);
                    if (floatd > p) coeff = 0;
                }
                raw[(size_t)i * fsize + j] = coeff;
                xx++;
            }
            x_dst_in_src += 2LL * xinc;
        }
    }

    /* reduce: drop near-zero heads (shifting left) and count near-zero tails */
    int min_size = 0;
    const double cutoff_lim = 0.002 * (double)fone;
    for (i = dst_n - 1; i >= 0; i--) {
        int64_t *row = raw + (size_t)i * fsize;
        int keep = fsize;
        int64_t cut = 0;
        for (j = 0; j < fsize; j++) {
            int k;
            cut += i64abs(row[0]);
            if ((double)cut > cutoff_lim) break;
            if (i < dst_n - 1 && pos[i] >= pos[i + 1]) break;
            for (k = 1; k < fsize; k++) row[k - 1] = row[k];
            row[fsize - 1] = 0;
            pos[i]++;
        }
        cut = 0;
        for (j = fsize - 1; j > 0; j--) {
            cut += i64abs(row[j]);
            if ((double)cut > cutoff_lim) break;
            keep--;
        }
        if (keep > min_size) min_size = keep;
    }
    if (min_size == 1 && align == 2) align = 1;
    const int osize = (min_size + (align - 1)) & ~(align - 1);

    int64_t *fin = (int64_t *)malloc(sizeof(int64_t) * (size_t)dst_n * (size_t)osize);
    for (i = 0; i < dst_n; i++)
        for (j = 0; j < osize; j++) {
            int64_t v = j >= fsize ? 0 : raw[(size_t)i * fsize + j];
            if (zero_tail && j >= min_size) v = 0;
            fin[(size_t)i * osize + j] = v;
        }
    free(raw);

    /* borders: fold taps that fall outside [0, src_n) onto the edge sample */
    for (i = 0; i < dst_n; i++) {
        int64_t *row = fin + (size_t)i * osize;
        if (pos[i] < 0) {
            for (j = 1; j < osize; j++) {
                int left = j + pos[i] > 0 ? j + pos[i] : 0;
                row[left] += row[j];
                row[j] = 0;
            }
            pos[i] = 0;
        }
        if (pos[i] + osize > src_n) {
            int shift = pos[i] + (osize - src_n < 0 ? osize - src_n : 0);
            int64_t acc = 0;
            for (j = osize - 1; j >= 0; j--)
                if (pos[i] + j >= src_n) { acc += row[j]; row[j] = 0; }
            for (j = osize - 1; j >= 0; j--)
                row[j] = j < shift ? 0 : row[j - shift];
            pos[i] -= shift;
            row[src_n - 1 - pos[i]] += acc;
        }
    }

    /* normalise to `one` with error diffusion */
    ora_filter *f = (ora_filter *)calloc(1, sizeof(*f));
    f->taps = osize; f->n = dst_n; f->pos = pos;
    f->coef = (int16_t *)calloc((size_t)osize * (size_t)(dst_n + 3), sizeof(int16_t));
    for (i = 0; i < dst_n; i++) {
        int64_t *row = fin + (size_t)i * osize;
        int64_t sum = 0, err = 0;
        for (j = 0; j < osize; j++) sum += row[j];
        sum = (sum + one / 2) / one;
        if (!sum) sum = 1;
        for (j = 0; j < osize; j++) {
            int64_t v = row[j] + err;
            int q = (int)rounded_div(v, sum);
            f->coef[(size_t)i * osize + j] = (int16_t)q;
            err = v - (int64_t)q * sum;
        }
    }
    free(fin);
    return f;
}

int ora_filter_taps(const ora_filter *f) { return f->taps; }
int ora_filter_n(const ora_filter *f) { return f->n; }
const int32_t *ora_filter_pos(const ora_filter *f) { return f->pos; }
const int16_t *ora_filter_coef(const ora_filter *f) { return f->coef; }

static inline uint8_t clip_u8(int v) { return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v)); }

/*
 * One 8-bit plane, sw x sh -> dw x dh.  target: 0 = A (x86 default flags), 1 = B (bitexact).
 * exact_from: first destination row that uses the exact (C) vertical function under target A
 *             (swscale.c switches to the C output functions once dstY >= dstH-2; for chroma the caller passes the
 *              first chroma row whose luma row 2*uv satisfies that).  Ignored for target B.
 * Unscaled plane (sw==dw && sh==dh) is the caller's job (plain copy; libswscale takes its unscaled fast path).
 */
void ora_scale_plane_u8(const uint8_t *src, int sstride, int sw, int sh,
                        uint8_t *dst, int dstride, int dw, int dh,
                        int kernel, int target, int exact_from)
{
    ora_filter *hf = ora_filter_make(sw, dw, kernel, target ? 1 : 4, 1 << 14, target);
    ora_filter *vf = ora_filter_make(sh, dh, kernel, target ? 1 : 2, 1 << 12, target);
    int16_t *tmp = (int16_t *)malloc(sizeof(int16_t) * (size_t)dw * (size_t)sh);
    int x, y, j;
    /* hScale8To15_c */
    for (y = 0; y < sh; y++) {
        const uint8_t *row = src + (size_t)y * sstride;
        int16_t *t = tmp + (size_t)y * dw;
        for (x = 0; x < dw; x++) {
            const int16_t *c = hf->coef + (size_t)x * hf->taps;
            const uint8_t *p = row + hf->pos[x];
            int val = 0;
            for (j = 0; j < hf->taps; j++) val += (int)p[j] * c[j];
            val >>= 7;
            t[x] = (int16_t)(val > 32767 ? 32767 : val);
        }
    }
    const int vs = vf->taps;
    for (y = 0; y < dh; y++) {
        const int16_t *c = vf->coef + (size_t)y * vs;
        const int16_t *t0 = tmp + (size_t)vf->pos[y] * dw;
        uint8_t *o = dst + (size_t)y * dstride;
        if (vs == 1) { /* yuv2plane1_8_c with the constant dither 64 */
            for (x = 0; x < dw; x++) o[x] = clip_u8((t0[x] + 64) >> 7);
        } else if (target == 1 || y >= exact_from) { /* yuv2planeX_8_c */
            for (x = 0; x < dw; x++) {
                int val = 64 << 12;
                for (j = 0; j < vs; j++) val += (int)t0[(size_t)j * dw + x] * c[j];
                o[x] = clip_u8(val >> 19);
            }
        } else { /* x86 yuv2yuvX: pmulhw / paddw / psraw 3 / packuswb */
            const int16_t init = (int16_t)((64 + 8 * (vs - 1)) >> 4);
            for (x = 0; x < dw; x++) {
                int16_t acc = init;
                for (j = 0; j < vs; j++)
                    acc = (int16_t)(acc + (int16_t)(((int)t0[(size_t)j * dw + x] * c[j]) >> 16));
                o[x] = clip_u8(acc >> 3);
            }
        }
    }
    free(tmp);
    ora_filter_free(hf); ora_filter_free(vf);
}

/*
 * One 10-bit plane held as plain integers 0..1023 in uint16 (i.e. P010 samples already >>6, chroma already
 * de-interleaved: p010LEToY_c / p010LEToUV_c).  Output again 0..1023 (caller stores <<6).
 * hScale16To15_c with sh = depth-1 = 9; yuv2p010lX_c / yuv2p010l1_c with shift 17 / 5.
 * Targets differ only by the coefficient tables.
 */
void ora_scale_plane_u10(const uint16_t *src, int sstride_px, int sw, int sh,
                         uint16_t *dst, int dstride_px, int dw, int dh,
                         int kernel, int target)
{
    ora_filter *hf = ora_filter_make(sw, dw, kernel, target ? 1 : 4, 1 << 14, target);
    ora_filter *vf = ora_filter_make(sh, dh, kernel, target ? 1 : 2, 1 << 12, target);
    int16_t *tmp = (int16_t *)malloc(sizeof(int16_t) * (size_t)dw * (size_t)sh);
    int x, y, j;
    for (y = 0; y < sh; y++) {
        const uint16_t *row = src + (size_t)y * sstride_px;
        int16_t *t = tmp + (size_t)y * dw;
        for (x = 0; x < dw; x++) {
            const int16_t *c = hf->coef + (size_t)x * hf->taps;
            const uint16_t *p = row + hf->pos[x];
            int val = 0;
            for (j = 0; j < hf->taps; j++) val += (int)p[j] * c[j];
            val >>= 9;
            t[x] = (int16_t)(val > 32767 ? 32767 : val);
        }
    }
    const int vs = vf->taps;
    for (y = 0; y < dh; y++) {
        const int16_t *c = vf->coef + (size_t)y * vs;
        const int16_t *t0 = tmp + (size_t)vf->pos[y] * dw;
        uint16_t *o = dst + (size_t)y * dstride_px;
        for (x = 0; x < dw; x++) {
            int val;
            if (vs == 1) {
                val = (t0[x] + (1 << 4)) >> 5;
            } else {
                val = 1 << 16;
                for (j = 0; j < vs; j++) val += (int)t0[(size_t)j * dw + x] * c[j];
                val >>= 17;
            }
            o[x] = (uint16_t)(val < 0 ? 0 : (val > 1023 ? 1023 : val));
        }
    }
    free(tmp);
    ora_filter_free(hf); ora_filter_free(vf);
}

static int chroma_exact_from(int dh)
{
    /* chroma row uv is produced while dstY == 2*uv; C functions are in use once dstY >= dstH-2 */
    int uv = (dh - 2 + 1) / 2;
    return uv < 0 ? 0 : uv;
}

/*
 * Whole-frame yuv420p -> yuv420p as the reference calls it (contiguous planes, strides w and ceil(w/2)).
 * Same size => identity copy (libswscale unscaled path).
 */
void ora_scale_i420(const uint8_t *src, int sw, int sh, uint8_t *dst, int dw, int dh, int kernel, int target)
{
    const int scw = (sw + 1) >> 1, sch = (sh + 1) >> 1, dcw = (dw + 1) >> 1, dch = (dh + 1) >> 1;
    if (sw == dw && sh == dh) { memcpy(dst, src, (size_t)sw * sh + 2 * (size_t)scw * sch); return; }
    const uint8_t *sy = src, *su = sy + (size_t)sw * sh, *sv = su + (size_t)scw * sch;
    uint8_t *dy = dst, *du = dy + (size_t)dw * dh, *dv = du + (size_t)dcw * dch;
    ora_scale_plane_u8(sy, sw, sw, sh, dy, dw, dw, dh, kernel, target, dh - 2);
    ora_scale_plane_u8(su, scw, scw, sch, du, dcw, dcw, dch, kernel, target, chroma_exact_from(dh));
    ora_scale_plane_u8(sv, scw, scw, sch, dv, dcw, dcw, dch, kernel, target, chroma_exact_from(dh));
}

/* Reference NV12 (transvideo.c:532-557): Y rows of the I420 result + byte-interleaved U,V of the I420 result. */
void ora_i420_to_nv12(const uint8_t *i420, int w, int h, uint8_t *nv12)
{
    const int cw = (w + 1) >> 1, ch = (h + 1) >> 1;
    const uint8_t *u = i420 + (size_t)w * h, *v = u + (size_t)cw * ch;
    uint8_t *uv = nv12 + (size_t)w * h;
    memcpy(nv12, i420, (size_t)w * h);
    for (size_t i = 0; i < (size_t)cw * ch; i++) { uv[2 * i] = u[i]; uv[2 * i + 1] = v[i]; }
}

/* P010 (Y plane then interleaved UV plane, little-endian 16-bit, 10 significant MSBs) -> P010. */
void ora_scale_p010(const uint16_t *src, int sw, int sh, uint16_t *dst, int dw, int dh, int kernel, int target)
{
    const int scw = (sw + 1) >> 1, sch = (sh + 1) >> 1, dcw = (dw + 1) >> 1, dch = (dh + 1) >> 1;
    if (sw == dw && sh == dh) { memcpy(dst, src, 2 * ((size_t)sw * sh + 2 * (size_t)scw * sch)); return; }
    size_t i;
    uint16_t *sy = (uint16_t *)malloc(2 * (size_t)sw * sh);
    uint16_t *su = (uint16_t *)malloc(2 * (size_t)scw * sch), *sv = (uint16_t *)malloc(2 * (size_t)scw * sch);
    uint16_t *dy = (uint16_t *)malloc(2 * (size_t)dw * dh);
    uint16_t *du = (uint16_t *)malloc(2 * (size_t)dcw * dch), *dv = (uint16_t *)malloc(2 * (size_t)dcw * dch);
    const uint16_t *suv = src + (size_t)sw * sh;
    for (i = 0; i < (size_t)sw * sh; i++) sy[i] = src[i] >> 6;
    for (i = 0; i < (size_t)scw * sch; i++) { su[i] = suv[2 * i] >> 6; sv[i] = suv[2 * i + 1] >> 6; }
    ora_scale_plane_u10(sy, sw, sw, sh, dy, dw, dw, dh, kernel, target);
    ora_scale_plane_u10(su, scw, scw, sch, du, dcw, dcw, dch, kernel, target);
    ora_scale_plane_u10(sv, scw, scw, sch, dv, dcw, dcw, dch, kernel, target);
    uint16_t *duv = dst + (size_t)dw * dh;
    for (i = 0; i < (size_t)dw * dh; i++) dst[i] = (uint16_t)(dy[i] << 6);
    for (i = 0; i < (size_t)dcw * dch; i++) { duv[2 * i] = (uint16_t)(du[i] << 6); duv[2 * i + 1] = (uint16_t)(dv[i] << 6); }
    free(sy); free(su); free(sv); free(dy); free(du); free(dv);
}
