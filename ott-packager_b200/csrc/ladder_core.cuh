// The ladder scaler's per-CTA work, written as barrier-separated PHASES: phase(tid, ctx) is called by every thread of
// the CTA (kernel: threadIdx.x, with __syncthreads() between phases).  The same text compiles as plain C++ when
// OTTGPU_EMULATE is defined, which tests/emu uses to run the CTA logic thread-by-thread on the CPU *as a test* of the
// kernel source before GPU time is spent; that build is never part of libottgpu.
//
// Arithmetic (bit-exact to libswscale's legacy scaler):
//   H  8-bit : t = min((sum src*c14) >> 7, 32767)      16-bit(10 significant): t = min((sum v10*c14) >> 9, 32767)
//   V  8-bit exact : clip_u8((64<<12 + sum t*c12) >> 19)
//   V  8-bit x86   : clip_u8(int16((64+8(vs-1))>>4 + sum (t*c12)>>16) >> 3)      (rows < exact_from, target A)
//   V  10-bit      : clip10((1<<16 + sum t*c12) >> 17)
//   one vertical tap (axis unscaled): 8-bit clip_u8((t+64)>>7), 10-bit clip10((t+16)>>5)
#pragma once
#include "device_types.h"

#if defined(OTTGPU_EMULATE)
#include <cstring>
#define OTT_DEV inline
#if !defined(__CUDACC__) && !defined(OTTGPU_EMU_VECTOR_TYPES)
#define OTTGPU_EMU_VECTOR_TYPES
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
#endif
namespace ottgpu {
inline uint32_t ott_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
inline int ott_dp4a_us(uint32_t a, uint32_t b, int c)
{
    for (int k = 0; k < 4; ++k) c += (int)((a >> (8 * k)) & 255) * (int)(int8_t)((b >> (8 * k)) & 255);
    return c;
}
inline uint32_t ott_dp4a_uu(uint32_t a, uint32_t b, uint32_t c)
{
    for (int k = 0; k < 4; ++k) c += ((a >> (8 * k)) & 255) * ((b >> (8 * k)) & 255);
    return c;
}
template <typename T> inline T ott_ldg(const T *p) { return *p; }
}  // namespace ottgpu
#else
#define OTT_DEV __device__ __forceinline__
namespace ottgpu {
OTT_DEV uint32_t ott_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
OTT_DEV int ott_dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
OTT_DEV uint32_t ott_dp4a_uu(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
template <typename T> OTT_DEV T ott_ldg(const T *p) { return __ldg(p); }
}  // namespace ottgpu
#endif

namespace ottgpu {

OTT_DEV int ott_min(int a, int b) { return a < b ? a : b; }
OTT_DEV int ott_max(int a, int b) { return a > b ? a : b; }
OTT_DEV int ott_clip(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// Everything one tile needs, resolved once per CTA (thread-uniform).
struct TileCtx {
    const PlaneTables *T;
    int bps;                 // source/destination bytes per sample
    int nplanes;             // 1 luma, 2 chroma pair
    int x0, y0, tw, th;      // output tile
    int sy0, sr;             // source rows [sy0, sy0+sr)
    int sx0, nwords;         // source columns start (sample index, word aligned) and 32-bit words per staged row
    int src_pitch_words;     // smem pitch of the staged source
    int mid_pitch;           // smem pitch (int16 elements) of the H-pass result, even
    int tw_log2;             // T->tw is 32 or 64
    int approx_v;
    // shared memory carve-up
    uint32_t *src_s;         // [nplanes*sr][src_pitch_words]
    int16_t *mid_s;          // [nplanes*sr][mid_pitch]
    int16_t *vc_s;           // [th][vtaps]
    int32_t *vpos_s;         // [th]
    // global operands
    const uint8_t *src[2];   // plane bases (chroma: U,V or the interleaved UV plane twice)
    int src_pitch[2];        // bytes
    int src_interleaved;     // chroma samples interleaved in the source (P010)
    int src_shift;           // 6 for P010 (MSB aligned), else 0
    uint8_t *dst[2];
    int dst_pitch[2];
    int dst_interleaved;     // NV12 / P010 chroma writer
    int dst_shift;           // 6 for P010
};

OTT_DEV size_t tile_smem_bytes(const PlaneTables &T, int nplanes)
{
    const int mid_pitch = (T.tw + 1) & ~1;
    size_t b = (size_t)nplanes * T.sr_max * T.sw_words * 4;      // staged source
    b += (size_t)nplanes * T.sr_max * mid_pitch * 2;             // H-pass result
    b += (size_t)T.th * T.vtaps * 2;                             // vertical coefficients of the tile's rows
    b = (b + 3) & ~(size_t)3;
    b += (size_t)T.th * 4;                                       // vpos
    return b;
}

// ------------------------------------------------------------------------------------------------ phase 0: tables
OTT_DEV void phase_tables(int tid, const TileCtx &c)
{
    const PlaneTables &T = *c.T;
    for (int i = tid; i < c.th * T.vtaps; i += kThreads) c.vc_s[i] = ott_ldg(T.vc + (size_t)c.y0 * T.vtaps + i);
    for (int i = tid; i < c.th; i += kThreads) c.vpos_s[i] = ott_ldg(T.vpos + c.y0 + i);
}

// ------------------------------------------------------------------------------------------------ phase 1: staging
// Copies the tile's source window into shared memory, one 32-bit word (4 or 2 samples) per thread step, rows spread
// over warps so consecutive lanes read consecutive global words.  Samples right of the plane edge read as zero (they
// only ever meet zero padding coefficients).
OTT_DEV uint32_t load_word8(const uint8_t *row, int x, int srcW, bool aligned)
{
    if (aligned && x + 4 <= srcW) return ott_ldg(reinterpret_cast<const uint32_t *>(row + x));
    uint32_t v = 0;
    for (int k = 0; k < 4; ++k)
        if (x + k < srcW) v |= (uint32_t)ott_ldg(row + x + k) << (8 * k);
    return v;
}

OTT_DEV uint32_t load_word16(const uint8_t *row, int x, int srcW, bool aligned, int shift)
{
    uint32_t v;
    if (aligned && x + 2 <= srcW) {
        v = ott_ldg(reinterpret_cast<const uint32_t *>(row + 2 * x));
    } else {
        v = 0;
        for (int k = 0; k < 2; ++k)
            if (x + k < srcW) v |= (uint32_t)ott_ldg(reinterpret_cast<const uint16_t *>(row + 2 * (x + k))) << (16 * k);
    }
    return (v >> shift) & (0xffffu >> shift) * 0x10001u;
}

// interleaved 16-bit chroma (P010): one 32-bit global word = (V<<16 | U) of one chroma sample
OTT_DEV void load_uv16(const uint8_t *row, int x, int srcW, int shift, uint32_t &u, uint32_t &v)
{
    uint32_t a = 0, b = 0;   // samples x and x+1
    if (x < srcW) a = ott_ldg(reinterpret_cast<const uint32_t *>(row + 4 * x));
    if (x + 1 < srcW) b = ott_ldg(reinterpret_cast<const uint32_t *>(row + 4 * (x + 1)));
    const uint32_t m = (0xffffu >> shift) * 0x10001u;
    a = (a >> shift) & m;
    b = (b >> shift) & m;
    u = (a & 0xffffu) | (b << 16);
    v = (a >> 16) | (b & 0xffff0000u);
}

OTT_DEV void phase_stage(int tid, const TileCtx &c)
{
    const PlaneTables &T = *c.T;
    const int warp = tid >> 5, lane = tid & 31;
    const int spw = c.bps == 1 ? 4 : 2;                  // samples per word
    if (c.src_interleaved) {                             // P010 chroma: de-interleave into the two staged planes
        const bool ok = (((uintptr_t)c.src[0] | (uintptr_t)c.src_pitch[0]) & 3) == 0;
        (void)ok;                                        // packed P010 rows are always 4-byte aligned (2*2*cw)
        for (int r = warp; r < c.sr; r += kThreads / 32) {
            const uint8_t *row = c.src[0] + (size_t)(c.sy0 + r) * c.src_pitch[0];
            for (int w = lane; w < c.nwords; w += 32) {
                uint32_t u, v;
                load_uv16(row, c.sx0 + 2 * w, T.srcW, c.src_shift, u, v);
                c.src_s[(size_t)r * c.src_pitch_words + w] = u;
                c.src_s[(size_t)(c.sr + r) * c.src_pitch_words + w] = v;
            }
        }
        return;
    }
    for (int p = 0; p < c.nplanes; ++p) {
        const bool aligned = (((uintptr_t)c.src[p] | (uintptr_t)c.src_pitch[p]) & 3) == 0;
        for (int r = warp; r < c.sr; r += kThreads / 32) {
            const uint8_t *row = c.src[p] + (size_t)(c.sy0 + r) * c.src_pitch[p];
            uint32_t *srow = c.src_s + (size_t)(p * c.sr + r) * c.src_pitch_words;
            for (int w = lane; w < c.nwords; w += 32) {
                const int x = c.sx0 + spw * w;
                srow[w] = c.bps == 1 ? load_word8(row, x, T.srcW, aligned)
                                     : load_word16(row, x, T.srcW, aligned, c.src_shift);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ phase 2: H pass
// Thread = one output column of the tile (coefficients live in registers), striding over the staged rows.
// 8-bit: four taps per dp4a, 14-bit coefficient split into signed high byte and unsigned low byte:
//        sum p*c = 256*sum p*(c>>8) + sum p*(c&255)   (exact)
template <int HQ>
OTT_DEV void hpass8(int tid, const TileCtx &c)
{
    const PlaneTables &T = *c.T;
    const int col = tid & (T.tw - 1), grp = tid >> c.tw_log2, ngroups = kThreads >> c.tw_log2;
    if (col >= c.tw) return;
    const int gx = c.x0 + col;
    const int off = ott_ldg(T.hpos + gx) - c.sx0;
    const int wi = off >> 2;
    const uint32_t sh = (uint32_t)(off & 3) * 8;
    uint32_t hi[HQ], lo[HQ];
#pragma unroll
    for (int q = 0; q < HQ; ++q) {
        hi[q] = ott_ldg(T.hc_hi + (size_t)gx * HQ + q);
        lo[q] = ott_ldg(T.hc_lo + (size_t)gx * HQ + q);
    }
    const int nrows = c.nplanes * c.sr;
    for (int row = grp; row < nrows; row += ngroups) {
        const uint32_t *w = c.src_s + (size_t)row * c.src_pitch_words + wi;
        uint32_t cur = w[0];
        int ahi = 0;
        uint32_t alo = 0;
#pragma unroll
        for (int q = 0; q < HQ; ++q) {
            const uint32_t nxt = w[q + 1];
            const uint32_t a = ott_funnel_r(cur, nxt, sh);
            ahi = ott_dp4a_us(a, hi[q], ahi);
            alo = ott_dp4a_uu(a, lo[q], alo);
            cur = nxt;
        }
        const int val = ahi * 2 + (int)(alo >> 7);       // == (256*ahi + alo) >> 7
        c.mid_s[(size_t)row * c.mid_pitch + col] = (int16_t)ott_min(val, 32767);
    }
}

// any tap count: coefficients re-read per row (L1-resident); used for very long filters (thumbnail, 4K->360p lanczos)
OTT_DEV void hpass8_any(int tid, const TileCtx &c)
{
    const PlaneTables &T = *c.T;
    const int col = tid & (T.tw - 1), grp = tid >> c.tw_log2, ngroups = kThreads >> c.tw_log2;
    if (col >= c.tw) return;
    const int gx = c.x0 + col, hq = T.hq;
    const int off = ott_ldg(T.hpos + gx) - c.sx0;
    const int wi = off >> 2;
    const uint32_t sh = (uint32_t)(off & 3) * 8;
    const uint32_t *phi = T.hc_hi + (size_t)gx * hq, *plo = T.hc_lo + (size_t)gx * hq;
    const int nrows = c.nplanes * c.sr;
    for (int row = grp; row < nrows; row += ngroups) {
        const uint32_t *w = c.src_s + (size_t)row * c.src_pitch_words + wi;
        uint32_t cur = w[0];
        int ahi = 0;
        uint32_t alo = 0;
        for (int q = 0; q < hq; ++q) {
            const uint32_t nxt = w[q + 1];
            const uint32_t a = ott_funnel_r(cur, nxt, sh);
            ahi = ott_dp4a_us(a, ott_ldg(phi + q), ahi);
            alo = ott_dp4a_uu(a, ott_ldg(plo + q), alo);
            cur = nxt;
        }
        const int val = ahi * 2 + (int)(alo >> 7);
        c.mid_s[(size_t)row * c.mid_pitch + col] = (int16_t)ott_min(val, 32767);
    }
}

// 16-bit samples (10 significant bits, already right-aligned by the staging): two samples per word, plain IMAD.
template <int HQ>
OTT_DEV void hpass16(int tid, const TileCtx &c)
{
    const PlaneTables &T = *c.T;
    const int col = tid & (T.tw - 1), grp = tid >> c.tw_log2, ngroups = kThreads >> c.tw_log2;
    if (col >= c.tw) return;
    const int gx = c.x0 + col;
    const int off = ott_ldg(T.hpos + gx) - c.sx0;
    const int wi = off >> 1;
    const uint32_t sh = (uint32_t)(off & 1) * 16;
    int cf[4 * HQ];
#pragma unroll
    for (int j = 0; j < 4 * HQ; ++j) cf[j] = ott_ldg(T.hc16 + (size_t)gx * (4 * HQ) + j);
    const int nrows = c.nplanes * c.sr;
    for (int row = grp; row < nrows; row += ngroups) {
        const uint32_t *w = c.src_s + (size_t)row * c.src_pitch_words + wi;
        uint32_t cur = w[0];
        int acc = 0;
#pragma unroll
        for (int k = 0; k < 2 * HQ; ++k) {
            const uint32_t nxt = w[k + 1];
            const uint32_t a = ott_funnel_r(cur, nxt, sh);
            acc += (int)(a & 0xffffu) * cf[2 * k] + (int)(a >> 16) * cf[2 * k + 1];
            cur = nxt;
        }
        c.mid_s[(size_t)row * c.mid_pitch + col] = (int16_t)ott_min(acc >> 9, 32767);
    }
}

OTT_DEV void hpass16_any(int tid, const TileCtx &c)
{
    const PlaneTables &T = *c.T;
    const int col = tid & (T.tw - 1), grp = tid >> c.tw_log2, ngroups = kThreads >> c.tw_log2;
    if (col >= c.tw) return;
    const int gx = c.x0 + col, taps = 4 * T.hq;
    const int off = ott_ldg(T.hpos + gx) - c.sx0;
    const int wi = off >> 1;
    const uint32_t sh = (uint32_t)(off & 1) * 16;
    const int16_t *pc = T.hc16 + (size_t)gx * taps;
    const int nrows = c.nplanes * c.sr;
    for (int row = grp; row < nrows; row += ngroups) {
        const uint32_t *w = c.src_s + (size_t)row * c.src_pitch_words + wi;
        uint32_t cur = w[0];
        int acc = 0;
        for (int k = 0; 2 * k < taps; ++k) {
            const uint32_t nxt = w[k + 1];
            const uint32_t a = ott_funnel_r(cur, nxt, sh);
            acc += (int)(a & 0xffffu) * ott_ldg(pc + 2 * k) + (int)(a >> 16) * ott_ldg(pc + 2 * k + 1);
            cur = nxt;
        }
        c.mid_s[(size_t)row * c.mid_pitch + col] = (int16_t)ott_min(acc >> 9, 32767);
    }
}

OTT_DEV void phase_hpass(int tid, const TileCtx &c)
{
    const int hq = c.T->hq;
    if (c.bps == 1) {
        switch (hq) {
        case 1: hpass8<1>(tid, c); break;
        case 2: hpass8<2>(tid, c); break;
        case 3: hpass8<3>(tid, c); break;
        case 4: hpass8<4>(tid, c); break;
        case 5: hpass8<5>(tid, c); break;
        case 6: hpass8<6>(tid, c); break;
        default: hpass8_any(tid, c); break;
        }
    } else {
        switch (hq) {
        case 1: hpass16<1>(tid, c); break;
        case 2: hpass16<2>(tid, c); break;
        case 3: hpass16<3>(tid, c); break;
        case 4: hpass16<4>(tid, c); break;
        case 5: hpass16<5>(tid, c); break;
        case 6: hpass16<6>(tid, c); break;
        default: hpass16_any(tid, c); break;
        }
    }
}

// ------------------------------------------------------------------------------------------------ phase 3: V pass + store
// Thread = two adjacent output columns (one 32-bit shared load per tap gives both), one warp-row at a time.
// mode: 0 exact 8-bit, 1 x86-truncating 8-bit, 2 10-bit, 3 single tap 8-bit, 4 single tap 10-bit
OTT_DEV void vcolumn2(const TileCtx &c, const uint32_t *m, int step_words, const int16_t *cf, int vtaps, int mode,
                      int &o0, int &o1)
{
    if (mode == 3 || mode == 4) {
        const uint32_t w = m[0];
        const int t0 = (int)(int16_t)(w & 0xffffu), t1 = (int)w >> 16;
        if (mode == 3) { o0 = ott_clip((t0 + 64) >> 7, 255); o1 = ott_clip((t1 + 64) >> 7, 255); }
        else { o0 = ott_clip((t0 + 16) >> 5, 1023); o1 = ott_clip((t1 + 16) >> 5, 1023); }
        return;
    }
    if (mode == 1) {
        int a0 = (64 + 8 * (vtaps - 1)) >> 4, a1 = a0;
        for (int j = 0; j < vtaps; ++j) {
            const uint32_t w = m[(size_t)j * step_words];
            const int cj = cf[j];
            a0 += ((int)(int16_t)(w & 0xffffu) * cj) >> 16;
            a1 += (((int)w >> 16) * cj) >> 16;
        }
        o0 = ott_clip((int)(int16_t)a0 >> 3, 255);         // 16-bit wraparound (paddw) then psraw 3, packuswb
        o1 = ott_clip((int)(int16_t)a1 >> 3, 255);
        return;
    }
    int a0 = mode == 0 ? 64 << 12 : 1 << 16, a1 = a0;
    for (int j = 0; j < vtaps; ++j) {
        const uint32_t w = m[(size_t)j * step_words];
        const int cj = cf[j];
        a0 += (int)(int16_t)(w & 0xffffu) * cj;
        a1 += ((int)w >> 16) * cj;
    }
    if (mode == 0) { o0 = ott_clip(a0 >> 19, 255); o1 = ott_clip(a1 >> 19, 255); }
    else { o0 = ott_clip(a0 >> 17, 1023); o1 = ott_clip(a1 >> 17, 1023); }
}

template <typename S>
OTT_DEV void store_run(uint8_t *row, int elem, const int *v, int n, int total)
{
    // n = 2 (planar) or 4 (interleaved pair) values starting at element index `elem`; `total` = elements in the row
    S *p = reinterpret_cast<S *>(row) + elem;
    const int live = ott_min(n, total - elem);
    constexpr int kVecBytes = (int)sizeof(S);
    const bool vec = (((uintptr_t)p) & (uintptr_t)(n * kVecBytes - 1)) == 0 && live == n;
    if (vec) {
        if (n * kVecBytes == 2) {
            *reinterpret_cast<uint16_t *>(p) = (uint16_t)(v[0] | (v[1] << 8));
        } else if (n * kVecBytes == 4) {
            if (kVecBytes == 1) *reinterpret_cast<uint32_t *>(p) = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24);
            else *reinterpret_cast<uint32_t *>(p) = (uint32_t)v[0] | ((uint32_t)v[1] << 16);
        } else {   // 4 x 16-bit
            uint2 q;
            q.x = (uint32_t)v[0] | ((uint32_t)v[1] << 16);
            q.y = (uint32_t)v[2] | ((uint32_t)v[3] << 16);
            *reinterpret_cast<uint2 *>(p) = q;
        }
    } else {
        for (int k = 0; k < live; ++k) p[k] = (S)v[k];
    }
}

OTT_DEV void phase_vpass(int tid, const TileCtx &c)
{
    const PlaneTables &T = *c.T;
    const int lanes_per_row = T.tw >> 1, lpr_log2 = c.tw_log2 - 1;
    const int lx = tid & (lanes_per_row - 1), ry = tid >> lpr_log2, rows_per_pass = kThreads >> lpr_log2;
    const int x = 2 * lx;
    if (x >= c.tw) return;
    const int step_words = c.mid_pitch >> 1;
    const int vtaps = T.vtaps;
    for (int y = ry; y < c.th; y += rows_per_pass) {
        const int gy = c.y0 + y;
        const int r0 = c.vpos_s[y] - c.sy0;
        const int16_t *cf = c.vc_s + (size_t)y * vtaps;
        int mode;
        if (c.bps == 1) mode = vtaps == 1 ? 3 : ((c.approx_v && gy < T.exact_from) ? 1 : 0);
        else mode = vtaps == 1 ? 4 : 2;
        int o[4];
        for (int p = 0; p < c.nplanes; ++p) {
            const uint32_t *m = reinterpret_cast<const uint32_t *>(c.mid_s + (size_t)(p * c.sr + r0) * c.mid_pitch + x);
            vcolumn2(c, m, step_words, cf, vtaps, mode, o[2 * p], o[2 * p + 1]);
        }
        const int gx = c.x0 + x;
        if (c.nplanes == 1) {
            int v[2] = { o[0] << c.dst_shift, o[1] << c.dst_shift };
            uint8_t *row = c.dst[0] + (size_t)gy * c.dst_pitch[0];
            if (c.bps == 1) store_run<uint8_t>(row, gx, v, 2, T.dstW);
            else store_run<uint16_t>(row, gx, v, 2, T.dstW);
        } else if (c.dst_interleaved) {
            int v[4] = { o[0] << c.dst_shift, o[2] << c.dst_shift, o[1] << c.dst_shift, o[3] << c.dst_shift };
            uint8_t *row = c.dst[0] + (size_t)gy * c.dst_pitch[0];
            if (c.bps == 1) store_run<uint8_t>(row, 2 * gx, v, 4, 2 * T.dstW);
            else store_run<uint16_t>(row, 2 * gx, v, 4, 2 * T.dstW);
        } else {
            for (int p = 0; p < 2; ++p) {
                int v[2] = { o[2 * p] << c.dst_shift, o[2 * p + 1] << c.dst_shift };
                uint8_t *row = c.dst[p] + (size_t)gy * c.dst_pitch[p];
                if (c.bps == 1) store_run<uint8_t>(row, gx, v, 2, T.dstW);
                else store_run<uint16_t>(row, gx, v, 2, T.dstW);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ copy items
// Same-size rung: rows [y0, y0+rows) of one plane (luma) or of the chroma planes.
OTT_DEV void copy_rows(int tid, const uint8_t *src, int spitch, uint8_t *dst, int dpitch, int row_bytes, int y0, int rows)
{
    const bool vec = ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)spitch | (uintptr_t)dpitch) & 15) == 0);
    if (vec) {
        const int nvec = row_bytes >> 4;
        const int warp = tid >> 5, lane = tid & 31;
        for (int r = warp; r < rows; r += kThreads / 32) {
            const uint4 *s = reinterpret_cast<const uint4 *>(src + (size_t)(y0 + r) * spitch);
            uint4 *d = reinterpret_cast<uint4 *>(dst + (size_t)(y0 + r) * dpitch);
            for (int i = lane; i < nvec; i += 32) d[i] = ott_ldg(s + i);
            for (int i = (nvec << 4) + lane; i < row_bytes; i += 32)
                dst[(size_t)(y0 + r) * dpitch + i] = src[(size_t)(y0 + r) * spitch + i];
        }
    } else {
        for (int r = 0; r < rows; ++r)
            for (int i = tid; i < row_bytes; i += kThreads)
                dst[(size_t)(y0 + r) * dpitch + i] = ott_ldg(src + (size_t)(y0 + r) * spitch + i);
    }
}

// I420 chroma -> NV12 interleaved (byte-wise U,V interleave of transvideo.c:543-557), 4 chroma samples per thread step
OTT_DEV void interleave_rows8(int tid, const uint8_t *u, const uint8_t *v, int spitch, uint8_t *dst, int dpitch, int cw,
                              int y0, int rows)
{
    const bool vec = ((((uintptr_t)u | (uintptr_t)v | (uintptr_t)spitch) & 3) == 0) &&
                     ((((uintptr_t)dst | (uintptr_t)dpitch) & 7) == 0);
    const int warp = tid >> 5, lane = tid & 31;
    for (int r = warp; r < rows; r += kThreads / 32) {
        const uint8_t *ur = u + (size_t)(y0 + r) * spitch, *vr = v + (size_t)(y0 + r) * spitch;
        uint8_t *d = dst + (size_t)(y0 + r) * dpitch;
        int done = 0;
        if (vec) {
            const int n4 = cw >> 2;
            for (int i = lane; i < n4; i += 32) {
                const uint32_t a = ott_ldg(reinterpret_cast<const uint32_t *>(ur) + i);
                const uint32_t b = ott_ldg(reinterpret_cast<const uint32_t *>(vr) + i);
                uint2 q;
                q.x = (a & 0xffu) | ((b & 0xffu) << 8) | ((a & 0xff00u) << 8) | ((b & 0xff00u) << 16);
                q.y = ((a >> 16) & 0xffu) | (((b >> 16) & 0xffu) << 8) | ((a >> 24) << 16) | ((b >> 24) << 24);
                reinterpret_cast<uint2 *>(d)[i] = q;
            }
            done = n4 << 2;
        }
        for (int i = done + lane; i < cw; i += 32) {
            d[2 * i] = ott_ldg(ur + i);
            d[2 * i + 1] = ott_ldg(vr + i);
        }
    }
}

// ------------------------------------------------------------------------------------------------ item dispatch
// Resolves one work item.  Returns 0: nothing to do, 1: scale tile (ctx filled), 2: copy item.
OTT_DEV int item_prepare(const Item &it, const Binding &b, unsigned char *smem, TileCtx &c)
{
    if (!b.active || ((b.skip_mask >> it.slot) & 1u)) return 0;
    const PlanDev &plan = *b.plan;
    if (it.slot == plan.thumb_slot && !b.thumb_active) return 0;
    if (it.kind == ITEM_COPY_LUMA || it.kind == ITEM_COPY_CHROMA) return 2;
    const RungDev &R = plan.rung[it.slot];
    const bool luma = it.kind == ITEM_SCALE_LUMA;
    const PlaneTables &T = luma ? R.luma : R.chroma;
    const int src_il = plan.src_fmt == 2 /*P010*/, dst_il = R.fmt == 1 /*NV12*/ || R.fmt == 2 /*P010*/;
    c.T = &T;
    c.bps = plan.bps;
    c.nplanes = luma ? 1 : 2;
    c.x0 = it.tx * T.tw;
    c.y0 = it.ty * T.th;
    c.tw = ott_min(T.tw, T.dstW - c.x0);
    c.th = ott_min(T.th, T.dstH - c.y0);
    c.sy0 = ott_ldg(T.vpos + c.y0);
    c.sr = ott_ldg(T.vpos + c.y0 + c.th - 1) + T.vtaps - c.sy0;
    const int spw = plan.bps == 1 ? 4 : 2;
    const int first = ott_ldg(T.hpos + c.x0), last = ott_ldg(T.hpos + c.x0 + c.tw - 1) + 4 * T.hq;
    c.sx0 = first & ~(spw - 1);
    c.nwords = (last - c.sx0 + spw - 1) / spw + 1;
    c.src_pitch_words = T.sw_words;
    c.mid_pitch = (T.tw + 1) & ~1;
    c.tw_log2 = T.tw == 64 ? 6 : 5;
    c.approx_v = R.approx_v;
    c.src_s = reinterpret_cast<uint32_t *>(smem);
    c.mid_s = reinterpret_cast<int16_t *>(c.src_s + (size_t)c.nplanes * T.sr_max * T.sw_words);
    c.vc_s = c.mid_s + (size_t)c.nplanes * T.sr_max * c.mid_pitch;
    c.vpos_s = reinterpret_cast<int32_t *>((reinterpret_cast<uintptr_t>(c.vc_s + (size_t)T.th * T.vtaps) + 3) & ~(uintptr_t)3);
    c.src_shift = plan.src_fmt == 2 ? 6 : 0;
    c.dst_shift = R.fmt == 2 ? 6 : 0;
    if (luma) {
        c.src[0] = b.src[0]; c.src_pitch[0] = b.src_pitch[0];
        c.src[1] = nullptr;  c.src_pitch[1] = 0;
        c.src_interleaved = 0;
        c.dst[0] = b.dst[it.slot][0]; c.dst_pitch[0] = b.dst_pitch[it.slot][0];
        c.dst[1] = nullptr; c.dst_pitch[1] = 0;
        c.dst_interleaved = 0;
    } else {
        c.src[0] = b.src[1]; c.src_pitch[0] = b.src_pitch[1];
        c.src[1] = src_il ? b.src[1] : b.src[2]; c.src_pitch[1] = src_il ? b.src_pitch[1] : b.src_pitch[2];
        c.src_interleaved = src_il;
        c.dst[0] = b.dst[it.slot][1]; c.dst_pitch[0] = b.dst_pitch[it.slot][1];
        c.dst[1] = dst_il ? nullptr : b.dst[it.slot][2]; c.dst_pitch[1] = dst_il ? 0 : b.dst_pitch[it.slot][2];
        c.dst_interleaved = dst_il;
    }
    return 1;
}

OTT_DEV void item_copy(int tid, const Item &it, const Binding &b)
{
    const PlanDev &plan = *b.plan;
    const RungDev &R = plan.rung[it.slot];
    const int src_il = plan.src_fmt == 2, dst_il = R.fmt == 1 || R.fmt == 2;
    if (it.kind == ITEM_COPY_LUMA) {
        const PlaneTables &T = R.luma;
        const int y0 = it.ty * T.th, rows = ott_min(T.th, T.dstH - y0);
        copy_rows(tid, b.src[0], b.src_pitch[0], b.dst[it.slot][0], b.dst_pitch[it.slot][0], T.dstW * plan.bps, y0, rows);
        return;
    }
    const PlaneTables &T = R.chroma;
    const int y0 = it.ty * T.th, rows = ott_min(T.th, T.dstH - y0);
    if (src_il == dst_il) {                 // planar->planar (two planes) or interleaved->interleaved (one plane)
        const int row_bytes = T.dstW * plan.bps * (src_il ? 2 : 1);
        copy_rows(tid, b.src[1], b.src_pitch[1], b.dst[it.slot][1], b.dst_pitch[it.slot][1], row_bytes, y0, rows);
        if (!src_il)
            copy_rows(tid, b.src[2], b.src_pitch[2], b.dst[it.slot][2], b.dst_pitch[it.slot][2], row_bytes, y0, rows);
    } else {                                // I420 -> NV12
        interleave_rows8(tid, b.src[1], b.src[2], b.src_pitch[1], b.dst[it.slot][1], b.dst_pitch[it.slot][1], T.dstW, y0, rows);
    }
}

}  // namespace ottgpu
