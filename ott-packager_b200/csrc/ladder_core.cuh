// The ladder scaler's per-CTA work, written as barrier-separated PHASES: phase(tid, geom, smem) is called by every thread
// of the CTA (kernel: threadIdx.x, with __syncthreads() between phases).  The same text compiles as plain C++ when
// OTTGPU_EMULATE is defined, which tests/emu uses to run the CTA logic thread-by-thread on the CPU *as a test* of the
// kernel source before GPU time is spent; that build is never part of libottgpu.
//
// Arithmetic (bit-exact to libswscale's legacy scaler):
//   H  8-bit : t = min((sum src*c14) >> 7, 32767)      16-bit(10 significant): t = min((sum v10*c14) >> 9, 32767)
//   V  8-bit exact : clip_u8((64<<12 + sum t*c12) >> 19)
//   V  8-bit x86   : clip_u8(int16((64+8(vs-1))>>4 + sum (t*c12)>>16) >> 3)      (rows < exact_from, target A)
//   V  10-bit      : clip10((1<<16 + sum t*c12) >> 17)
//   one vertical tap (axis unscaled): 8-bit clip_u8((t+64)>>7), 10-bit clip10((t+16)>>5)
//
// Shared-memory layout of one tile (all sizes from PlaneTables, so planner and kernel agree):
//   [ staged source : nplanes * sr_max rows * sw_words 32-bit words ]   4 (8-bit) or 2 (16-bit) samples per word
//   [ H-pass result : nplanes * (sr_max+1) rows * mid_pitch mid_t   ]   mid_pitch = tw (32|64), rows 16-byte aligned
//   [ V coefficients: th * vtaps int16 ][ vpos: th int32 ]
#pragma once
#include "device_types.h"

#if defined(OTTGPU_EMULATE)
#include <cstring>
#define OTT_DEV inline
#define OTT_UNROLL
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
inline int ott_dp2a_us(uint32_t a, uint32_t b, int c, int hi)      // two u16 of a x two s8 of b (bytes 0,1 or 2,3)
{
    const uint32_t bb = hi ? b >> 16 : b;
    return c + (int)(a & 0xffffu) * (int)(int8_t)(bb & 255) + (int)(a >> 16) * (int)(int8_t)((bb >> 8) & 255);
}
inline uint32_t ott_dp2a_uu(uint32_t a, uint32_t b, uint32_t c, int hi)
{
    const uint32_t bb = hi ? b >> 16 : b;
    return c + (a & 0xffffu) * (bb & 255) + (a >> 16) * ((bb >> 8) & 255);
}
inline int ott_dp2a_su(uint32_t a, uint32_t b, int c, int hi)      // two s16 of a x two u8 of b (bytes 0,1 or 2,3)
{
    const uint32_t bb = hi ? b >> 16 : b;
    return c + (int)(int16_t)(a & 0xffffu) * (int)(bb & 255) + (int)(int16_t)(a >> 16) * (int)((bb >> 8) & 255);
}
inline uint32_t ott_prmt(uint32_t a, uint32_t b, uint32_t sel)     // byte permute, selector nibbles 0-7 (no sign replication)
{
    const uint64_t v = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) r |= (uint32_t)((v >> (8 * ((sel >> (4 * k)) & 7))) & 0xffu) << (8 * k);
    return r;
}
inline uint32_t ott_vabsdiff4(uint32_t a, uint32_t b)              // per-byte |a-b|
{
    uint32_t r = 0;
    for (int k = 0; k < 4; ++k) {
        const int x = (int)((a >> (8 * k)) & 255) - (int)((b >> (8 * k)) & 255);
        r |= (uint32_t)(x < 0 ? -x : x) << (8 * k);
    }
    return r;
}
inline uint32_t ott_lanes16(int lo, int hi) { return ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16); }
inline int ott_lane_lo(uint32_t v) { return (int)(int16_t)(v & 0xffffu); }
inline int ott_lane_hi(uint32_t v) { return (int)(int16_t)(v >> 16); }
inline uint32_t ott_vmax_s16x2(uint32_t a, uint32_t b)
{
    return ott_lanes16(ott_lane_lo(a) > ott_lane_lo(b) ? ott_lane_lo(a) : ott_lane_lo(b), ott_lane_hi(a) > ott_lane_hi(b) ? ott_lane_hi(a) : ott_lane_hi(b));
}
inline uint32_t ott_vmin_s16x2(uint32_t a, uint32_t b)
{
    return ott_lanes16(ott_lane_lo(a) < ott_lane_lo(b) ? ott_lane_lo(a) : ott_lane_lo(b), ott_lane_hi(a) < ott_lane_hi(b) ? ott_lane_hi(a) : ott_lane_hi(b));
}
inline uint32_t ott_vmax3_s16x2(uint32_t a, uint32_t b, uint32_t c) { return ott_vmax_s16x2(ott_vmax_s16x2(a, b), c); }
inline uint32_t ott_vmin3_s16x2(uint32_t a, uint32_t b, uint32_t c) { return ott_vmin_s16x2(ott_vmin_s16x2(a, b), c); }
template <typename T> inline T ott_ldg(const T *p) { return *p; }
inline int ott_sad(int a, int b, int c) { return (a > b ? a - b : b - a) + c; }
inline int ott_byte(uint32_t w, int i) { return (int)((w >> (8 * i)) & 0xffu); }
inline uint32_t ott_lo16x2(uint32_t a, uint32_t b) { return (a & 0xffffu) | (b << 16); }          // (a.lo, b.lo)
inline uint32_t ott_hi16x2(uint32_t a, uint32_t b) { return (a >> 16) | (b & 0xffff0000u); }     // (a.hi, b.hi)
}  // namespace ottgpu
#else
#define OTT_DEV __device__ __forceinline__
#define OTT_UNROLL _Pragma("unroll")
namespace ottgpu {
OTT_DEV uint32_t ott_funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
OTT_DEV int ott_dp4a_us(uint32_t a, uint32_t b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
OTT_DEV uint32_t ott_dp4a_uu(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
OTT_DEV int ott_dp2a_us(uint32_t a, uint32_t b, int c, int hi)
{
    int d;
    if (hi) asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
OTT_DEV int ott_dp2a_su(uint32_t a, uint32_t b, int c, int hi)
{
    int d;
    if (hi) asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
OTT_DEV uint32_t ott_dp2a_uu(uint32_t a, uint32_t b, uint32_t c, int hi)
{
    uint32_t d;
    if (hi) asm("dp2a.hi.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
OTT_DEV uint32_t ott_prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
OTT_DEV uint32_t ott_vabsdiff4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }                 // VABSDIFF4.U8
OTT_DEV uint32_t ott_vmax_s16x2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }                    // VIMNMX.S16x2
OTT_DEV uint32_t ott_vmin_s16x2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
OTT_DEV uint32_t ott_vmax3_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }   // VIMNMX3.S16x2
OTT_DEV uint32_t ott_vmin3_s16x2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_s16x2(a, b, c); }
template <typename T> OTT_DEV T ott_ldg(const T *p) { return __ldg(p); }
OTT_DEV int ott_sad(int a, int b, int c) { return (int)__sad(a, b, (unsigned)c); }   // |a-b|+c in one VABSDIFF
OTT_DEV int ott_byte(uint32_t w, int i) { return (int)__byte_perm(w, 0, 0x4440 | i); }   // one PRMT instead of shift+mask
OTT_DEV uint32_t ott_lo16x2(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x5410); }
OTT_DEV uint32_t ott_hi16x2(uint32_t a, uint32_t b) { return __byte_perm(a, b, 0x7632); }
}  // namespace ottgpu
#endif

namespace ottgpu {

OTT_DEV int ott_min(int a, int b) { return a < b ? a : b; }
OTT_DEV int ott_max(int a, int b) { return a > b ? a : b; }
OTT_DEV int ott_clip(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// Everything one tile needs, resolved once per CTA (thread-uniform scalars; no shared-memory pointers are kept in
// here so that the compiler keeps every shared access in the shared address space).
struct TileGeom {
    const PlaneTables *T;
    int bps;                 // source/destination bytes per sample
    int nplanes;             // 1 luma, 2 chroma pair
    int x0, y0, tw, th;      // output tile
    int sy0, sr;             // source rows [sy0, sy0+sr)
    int sx0;                 // first staged source sample (multiple of 16 bytes)
    int nvec;                // 16-byte vectors staged per row
    int spw;                 // smem pitch of the staged source, 32-bit words (multiple of 4)
    int mid_pitch;           // smem pitch (mid_t elements) of the H-pass result: 32 or 64
    int tw_log2;             // T->tw is 32 or 64
    int approx_v;
    int off_src, off_mid, off_vc, off_vpos;   // byte offsets of the regions inside the CTA's shared memory
    int plane_stride;        // bytes between the two staged chroma planes inside a source buffer (128-byte multiple)
    int bulk_ok;             // the tile's source window is fetched by the TMA engine (one 2-D box per plane)
    int box_bytes;           // bytes one TMA box delivers
    const unsigned char *tmap0, *tmap1;   // tensor maps of the plane(s)
    const uint8_t *src0, *src1;      // plane bases (chroma: U,V or the interleaved UV plane twice)
    int src_pitch0, src_pitch1;      // bytes
    int src_interleaved;     // chroma samples interleaved in the source (P010): staged as ONE plane of (U,V) words
    int staged_planes;       // planes in the staged window (1 for luma and for interleaved chroma, else 2)
    int px_bytes;            // bytes per staged sample position: 1 (8-bit), 2 (16-bit), 4 (interleaved 16-bit pair)
    int tma_x;               // TMA box x coordinate in 32-bit elements
    int src_shift;           // 6 for P010 (MSB aligned), else 0
    uint8_t *dst0, *dst1;
    int dst_pitch0, dst_pitch1;
    int dst_interleaved;     // NV12 / P010 chroma writer
    int dst_shift;           // 6 for P010
};

// ------------------------------------------------------------------------------------------------ phase 0: tables
// Copies the tile's vertical coefficients and row positions into its table slot; run by `nthreads` cooperating threads
// (the producer warp in the kernel: the tables are in place before the compute warps are released).
OTT_DEV void phase_tables(int tid, int nthreads, const TileGeom &g, unsigned char *smem)
{
    const PlaneTables &T = *g.T;
    int16_t *vc_s = reinterpret_cast<int16_t *>(smem + g.off_vc);
    int32_t *vpos_s = reinterpret_cast<int32_t *>(smem + g.off_vpos);
    const int n = g.th * T.vtaps, th = g.th;
    const int16_t *vc = T.vc + g.y0 * T.vtaps;
    const int32_t *vp = T.vpos + g.y0;
    if (((n | (g.y0 * T.vtaps)) & 1) == 0) {               // pairs of coefficients per load
        const uint32_t *s2 = reinterpret_cast<const uint32_t *>(vc);
        uint32_t *d2 = reinterpret_cast<uint32_t *>(vc_s);
        for (int i = tid; i < (n >> 1); i += nthreads) d2[i] = ott_ldg(s2 + i);
    } else {
        for (int i = tid; i < n; i += nthreads) vc_s[i] = ott_ldg(vc + i);
    }
    for (int i = tid; i < th; i += nthreads) vpos_s[i] = ott_ldg(vp + i);
}

// ------------------------------------------------------------------------------------------------ phase 1: staging
// Copies the tile's source window into shared memory in 16-byte vectors (LDG.128 -> STS.128) when the plane is 16-byte
// aligned, else sample by sample.  Samples right of the plane edge read as zero (they only meet zero padding taps).
OTT_DEV uint4 load_vec_slow(const uint8_t *row, int byte_x, int row_bytes)
{
    uint32_t w[4] = {0, 0, 0, 0};
    OTT_UNROLL
    for (int i = 0; i < 4; ++i) {
        OTT_UNROLL
        for (int k = 0; k < 4; ++k)
            if (byte_x + 4 * i + k < row_bytes) w[i] |= (uint32_t)ott_ldg(row + byte_x + 4 * i + k) << (8 * k);
    }
    uint4 v;
    v.x = w[0]; v.y = w[1]; v.z = w[2]; v.w = w[3];
    return v;
}

OTT_DEV uint32_t shift_mask16(uint32_t v, int shift) { return (v >> shift) & ((0xffffu >> shift) * 0x10001u); }

OTT_DEV void phase_stage(int tid, const TileGeom &g, unsigned char *smem)
{
    // Register path (planes that are not 16-byte aligned): RAW bytes, same layout the TMA box produces; P010's >>6 is
    // applied afterwards by phase_fixup16 on both paths.
    const PlaneTables &T = *g.T;
    uint32_t *src_s = reinterpret_cast<uint32_t *>(smem + g.off_src);
    // lanes per staged row: smallest power of two >= nvec (capped at 32)
    int lpr_log2 = 0;
    while ((1 << lpr_log2) < g.nvec && lpr_log2 < 5) ++lpr_log2;
    const int lpr = 1 << lpr_log2, v0 = tid & (lpr - 1), r0 = tid >> lpr_log2, rstep = kThreads >> lpr_log2;
    const int sr = g.sr, sy0 = g.sy0, spw = g.spw, nvec = g.nvec, planes = g.staged_planes;
    const int row_bytes = T.srcW * g.px_bytes;
    const int bx0 = g.sx0 * g.px_bytes;
    for (int p = 0; p < planes; ++p) {
        const uint8_t *base = p ? g.src1 : g.src0;
        const int pitch = p ? g.src_pitch1 : g.src_pitch0;
        const bool aligned = (((uintptr_t)base | (uintptr_t)pitch) & 15) == 0;
        for (int r = r0; r < sr; r += rstep) {
            const uint8_t *row = base + (size_t)(sy0 + r) * pitch;
            uint4 *srow = reinterpret_cast<uint4 *>(src_s + p * (g.plane_stride >> 2) + r * spw);
            for (int v = v0; v < nvec; v += lpr) {
                const int bx = bx0 + 16 * v;
                srow[v] = (aligned && bx + 16 <= row_bytes) ? ott_ldg(reinterpret_cast<const uint4 *>(row + bx))
                                                            : load_vec_slow(row, bx, row_bytes);
            }
        }
    }
}

// P010 keeps its 10 significant bits in the upper part of each 16-bit sample: right-align them in place (both halves of
// every staged word, i.e. two luma samples or one (U,V) pair) before the H pass reads the window.
OTT_DEV void phase_fixup16(int tid, const TileGeom &g, unsigned char *smem)
{
    const int shift = g.src_shift;
    if (!shift) return;
    uint32_t *src_s = reinterpret_cast<uint32_t *>(smem + g.off_src);
    int lpr_log2 = 0;
    while ((1 << lpr_log2) < g.nvec && lpr_log2 < 5) ++lpr_log2;
    const int lpr = 1 << lpr_log2, v0 = tid & (lpr - 1), r0 = tid >> lpr_log2, rstep = kThreads >> lpr_log2;
    const int sr = g.sr, spw = g.spw, nvec = g.nvec;
    for (int p = 0; p < g.staged_planes; ++p)
        for (int r = r0; r < sr; r += rstep) {
            uint4 *srow = reinterpret_cast<uint4 *>(src_s + p * (g.plane_stride >> 2) + r * spw);
            for (int v = v0; v < nvec; v += lpr) {
                uint4 q = srow[v];
                q.x = shift_mask16(q.x, shift); q.y = shift_mask16(q.y, shift);
                q.z = shift_mask16(q.z, shift); q.w = shift_mask16(q.w, shift);
                srow[v] = q;
            }
        }
}

// ------------------------------------------------------------------------------------------------ phase 2: H pass
// Thread = one output column of the tile (its coefficients live in registers), striding over the staged rows.
// 8-bit: four taps per dp4a, 14-bit coefficient split into signed high byte and unsigned low byte:
//        sum p*c = 256*sum p*(c>>8) + sum p*(c&255)   (exact)
// Fixed-length filters: the 14-bit coefficients are used whole, as s16 pairs against two source bytes each (dp2a),
// positioned at table-build time against the aligned source words the window starts in - no shift, no hi/lo recombination.
template <int HQ>
OTT_DEV int hdot8(const uint32_t *w, const uint32_t (&c01)[HQ + 1], const uint32_t (&c23)[HQ + 1])
{
    int a = 0;
    OTT_UNROLL
    for (int q = 0; q <= HQ; ++q) {
        const uint32_t s = w[q];
        a = ott_dp2a_su(c01[q], s, a, 0);
        a = ott_dp2a_su(c23[q], s, a, 1);
    }
    return ott_min(a >> 7, 32767);
}

template <int HQ>
OTT_DEV void hpass8(int tid, const TileGeom &g, unsigned char *smem)
{
    const PlaneTables &T = *g.T;
    const int col = tid & (T.tw - 1), grp = tid >> g.tw_log2, ngroups = kThreads >> g.tw_log2;
    if (col >= g.tw) return;
    const int gx = g.x0 + col;
    const int off = ott_ldg(T.hpos + gx) - g.sx0;
    uint32_t c01[HQ + 1], c23[HQ + 1];
    OTT_UNROLL
    for (int q = 0; q <= HQ; ++q) {
        const uint2 c = ott_ldg(reinterpret_cast<const uint2 *>(T.hc_pair) + gx * (HQ + 1) + q);
        c01[q] = c.x;
        c23[q] = c.y;
    }
    const int sr = g.sr, nplanes = g.nplanes, pstride = g.plane_stride >> 2;
    const int wstep = ngroups * g.spw, mstep = ngroups * g.mid_pitch;
    for (int p = 0; p < nplanes; ++p) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(smem + g.off_src) + p * pstride + (off >> 2) + grp * g.spw;
        mid_t *m = reinterpret_cast<mid_t *>(smem + g.off_mid) + col + (p * sr + grp) * g.mid_pitch;
        int left = (sr - grp + ngroups - 1) / ngroups;         // rows this thread owns in this plane
        for (; left >= 4; left -= 4) {                         // four rows in flight per thread
            const int v0 = hdot8<HQ>(w, c01, c23);
            const int v1 = hdot8<HQ>(w + wstep, c01, c23);
            const int v2 = hdot8<HQ>(w + 2 * wstep, c01, c23);
            const int v3 = hdot8<HQ>(w + 3 * wstep, c01, c23);
            m[0] = (mid_t)v0;
            m[mstep] = (mid_t)v1;
            m[2 * mstep] = (mid_t)v2;
            m[3 * mstep] = (mid_t)v3;
            w += 4 * wstep;
            m += 4 * mstep;
        }
        for (; left > 0; --left) {
            m[0] = (mid_t)hdot8<HQ>(w, c01, c23);
            w += wstep;
            m += mstep;
        }
    }
}

// any tap count: coefficients re-read per row (L1-resident); used for very long filters (thumbnail, 4K->360p lanczos)
OTT_DEV void hpass8_any(int tid, const TileGeom &g, unsigned char *smem)
{
    const PlaneTables &T = *g.T;
    const int col = tid & (T.tw - 1), grp = tid >> g.tw_log2, ngroups = kThreads >> g.tw_log2;
    if (col >= g.tw) return;
    const int gx = g.x0 + col, hq = T.hq;
    const int off = ott_ldg(T.hpos + gx) - g.sx0;
    const uint32_t sh = (uint32_t)(off & 3) * 8;
    const uint32_t *phi = T.hc_hi + gx * hq, *plo = T.hc_lo + gx * hq;
    const int nrows = g.nplanes * g.sr, sr = g.sr, pstride = g.plane_stride >> 2;
    const uint32_t *src_s = reinterpret_cast<const uint32_t *>(smem + g.off_src) + (off >> 2);
    mid_t *mid_s = reinterpret_cast<mid_t *>(smem + g.off_mid) + col;
    for (int row = grp; row < nrows; row += ngroups) {
        const int p = row >= sr;
        const uint32_t *w = src_s + p * pstride + (row - p * sr) * g.spw;
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
        mid_s[row * g.mid_pitch] = (mid_t)ott_min((ahi * 256 + (int)alo) >> 7, 32767);
    }
}

// 16-bit samples (10 significant bits, already right-aligned by the staging): two samples per word.  Same coefficient
// split as the 8-bit path, two taps per dp2a: sum v*c = 256*sum v*(c>>8) + sum v*(c&255)   (exact).
template <int HQ>
OTT_DEV int hdot16(const uint32_t *w, uint32_t sh, const uint32_t (&hi)[HQ], const uint32_t (&lo)[HQ])
{
    uint32_t cur = w[0];
    int ahi = 0;
    uint32_t alo = 0;
    OTT_UNROLL
    for (int q = 0; q < HQ; ++q) {
        const uint32_t w1 = w[2 * q + 1], w2 = w[2 * q + 2];
        const uint32_t a0 = ott_funnel_r(cur, w1, sh), a1 = ott_funnel_r(w1, w2, sh);
        ahi = ott_dp2a_us(a0, hi[q], ahi, 0);
        ahi = ott_dp2a_us(a1, hi[q], ahi, 1);
        alo = ott_dp2a_uu(a0, lo[q], alo, 0);
        alo = ott_dp2a_uu(a1, lo[q], alo, 1);
        cur = w2;
    }
    return ott_min((ahi * 256 + (int)alo) >> 9, 32767);
}

template <int HQ>
OTT_DEV void hpass16(int tid, const TileGeom &g, unsigned char *smem)
{
    const PlaneTables &T = *g.T;
    const int col = tid & (T.tw - 1), grp = tid >> g.tw_log2, ngroups = kThreads >> g.tw_log2;
    if (col >= g.tw) return;
    const int gx = g.x0 + col;
    const int off = ott_ldg(T.hpos + gx) - g.sx0;
    const uint32_t sh = (uint32_t)(off & 1) * 16;
    uint32_t hi[HQ], lo[HQ];
    OTT_UNROLL
    for (int q = 0; q < HQ; ++q) {
        hi[q] = ott_ldg(T.hc_hi + gx * HQ + q);
        lo[q] = ott_ldg(T.hc_lo + gx * HQ + q);
    }
    const int sr = g.sr, nplanes = g.nplanes, pstride = g.plane_stride >> 2;
    const int wstep = ngroups * g.spw, mstep = ngroups * g.mid_pitch;
    for (int p = 0; p < nplanes; ++p) {
        const uint32_t *w = reinterpret_cast<const uint32_t *>(smem + g.off_src) + p * pstride + (off >> 1) + grp * g.spw;
        mid_t *m = reinterpret_cast<mid_t *>(smem + g.off_mid) + col + (p * sr + grp) * g.mid_pitch;
        int left = (sr - grp + ngroups - 1) / ngroups;
        for (; left >= 2; left -= 2) {                         // two rows in flight per thread
            const int v0 = hdot16<HQ>(w, sh, hi, lo);
            const int v1 = hdot16<HQ>(w + wstep, sh, hi, lo);
            m[0] = (mid_t)v0;
            m[mstep] = (mid_t)v1;
            w += 2 * wstep;
            m += 2 * mstep;
        }
        if (left > 0) m[0] = (mid_t)hdot16<HQ>(w, sh, hi, lo);
    }
}

// Interleaved 16-bit chroma (P010): every staged word is one (U,V) pair, so both planes read the same rows and the pair
// of samples a dp2a needs is one PRMT of two neighbouring words (low halves = U, high halves = V); no alignment shift.
template <int HQ>
OTT_DEV int hdot16il(const uint32_t *w, int plane, const uint32_t (&hi)[HQ], const uint32_t (&lo)[HQ])
{
    int ahi = 0;
    uint32_t alo = 0;
    OTT_UNROLL
    for (int q = 0; q < HQ; ++q) {
        const uint32_t w0 = w[4 * q], w1 = w[4 * q + 1], w2 = w[4 * q + 2], w3 = w[4 * q + 3];
        const uint32_t a0 = plane ? ott_hi16x2(w0, w1) : ott_lo16x2(w0, w1);
        const uint32_t a1 = plane ? ott_hi16x2(w2, w3) : ott_lo16x2(w2, w3);
        ahi = ott_dp2a_us(a0, hi[q], ahi, 0);
        ahi = ott_dp2a_us(a1, hi[q], ahi, 1);
        alo = ott_dp2a_uu(a0, lo[q], alo, 0);
        alo = ott_dp2a_uu(a1, lo[q], alo, 1);
    }
    return ott_min((ahi * 256 + (int)alo) >> 9, 32767);
}

template <int HQ>
OTT_DEV void hpass16il(int tid, const TileGeom &g, unsigned char *smem)
{
    const PlaneTables &T = *g.T;
    const int col = tid & (T.tw - 1), grp = tid >> g.tw_log2, ngroups = kThreads >> g.tw_log2;
    if (col >= g.tw) return;
    const int gx = g.x0 + col;
    const int off = ott_ldg(T.hpos + gx) - g.sx0;            // in samples = staged words
    uint32_t hi[HQ], lo[HQ];
    OTT_UNROLL
    for (int q = 0; q < HQ; ++q) {
        hi[q] = ott_ldg(T.hc_hi + gx * HQ + q);
        lo[q] = ott_ldg(T.hc_lo + gx * HQ + q);
    }
    const int sr = g.sr, spw = g.spw, mp = g.mid_pitch;
    const uint32_t *src_s = reinterpret_cast<const uint32_t *>(smem + g.off_src) + off;
    mid_t *mid_s = reinterpret_cast<mid_t *>(smem + g.off_mid) + col;
    for (int row = grp; row < 2 * sr; row += ngroups) {      // rows 0..sr-1 -> U, sr..2sr-1 -> V
        const int p = row >= sr, r = row - p * sr;
        mid_s[row * mp] = (mid_t)hdot16il<HQ>(src_s + r * spw, p, hi, lo);
    }
}

OTT_DEV void hpass16il_any(int tid, const TileGeom &g, unsigned char *smem)
{
    const PlaneTables &T = *g.T;
    const int col = tid & (T.tw - 1), grp = tid >> g.tw_log2, ngroups = kThreads >> g.tw_log2;
    if (col >= g.tw) return;
    const int gx = g.x0 + col, taps = 4 * T.hq;
    const int off = ott_ldg(T.hpos + gx) - g.sx0;
    const int16_t *pc = T.hc16 + gx * taps;
    const int sr = g.sr, spw = g.spw, mp = g.mid_pitch;
    const uint32_t *src_s = reinterpret_cast<const uint32_t *>(smem + g.off_src) + off;
    mid_t *mid_s = reinterpret_cast<mid_t *>(smem + g.off_mid) + col;
    for (int row = grp; row < 2 * sr; row += ngroups) {
        const int p = row >= sr, r = row - p * sr;
        const uint32_t *w = src_s + r * spw;
        int acc = 0;
        for (int k = 0; k < taps; ++k) {
            const uint32_t v = w[k];
            acc += (int)(p ? v >> 16 : v & 0xffffu) * ott_ldg(pc + k);
        }
        mid_s[row * mp] = (mid_t)ott_min(acc >> 9, 32767);
    }
}

OTT_DEV void hpass16_any(int tid, const TileGeom &g, unsigned char *smem)
{
    const PlaneTables &T = *g.T;
    const int col = tid & (T.tw - 1), grp = tid >> g.tw_log2, ngroups = kThreads >> g.tw_log2;
    if (col >= g.tw) return;
    const int gx = g.x0 + col, taps = 4 * T.hq;
    const int off = ott_ldg(T.hpos + gx) - g.sx0;
    const uint32_t sh = (uint32_t)(off & 1) * 16;
    const int16_t *pc = T.hc16 + gx * taps;
    const int nrows = g.nplanes * g.sr, sr = g.sr, pstride = g.plane_stride >> 2;
    const uint32_t *src_s = reinterpret_cast<const uint32_t *>(smem + g.off_src) + (off >> 1);
    mid_t *mid_s = reinterpret_cast<mid_t *>(smem + g.off_mid) + col;
    for (int row = grp; row < nrows; row += ngroups) {
        const int p = row >= sr;
        const uint32_t *w = src_s + p * pstride + (row - p * sr) * g.spw;
        uint32_t cur = w[0];
        int acc = 0;
        for (int k = 0; 2 * k < taps; ++k) {
            const uint32_t nxt = w[k + 1];
            const uint32_t a = ott_funnel_r(cur, nxt, sh);
            acc += (int)(a & 0xffffu) * ott_ldg(pc + 2 * k) + (int)(a >> 16) * ott_ldg(pc + 2 * k + 1);
            cur = nxt;
        }
        mid_s[row * g.mid_pitch] = (mid_t)ott_min(acc >> 9, 32767);
    }
}

OTT_DEV void phase_hpass(int tid, const TileGeom &g, unsigned char *smem)
{
    const int hq = g.T->hq;
    if (g.bps == 1) {
        switch (hq) {
        case 1: hpass8<1>(tid, g, smem); break;
        case 2: hpass8<2>(tid, g, smem); break;
        case 3: hpass8<3>(tid, g, smem); break;
        case 4: hpass8<4>(tid, g, smem); break;
        case 5: hpass8<5>(tid, g, smem); break;
        case 6: hpass8<6>(tid, g, smem); break;
        default: hpass8_any(tid, g, smem); break;
        }
    } else if (g.src_interleaved) {
        switch (hq) {
        case 1: hpass16il<1>(tid, g, smem); break;
        case 2: hpass16il<2>(tid, g, smem); break;
        case 3: hpass16il<3>(tid, g, smem); break;
        case 4: hpass16il<4>(tid, g, smem); break;
        case 5: hpass16il<5>(tid, g, smem); break;
        case 6: hpass16il<6>(tid, g, smem); break;
        case 7: hpass16il<7>(tid, g, smem); break;
        case 8: hpass16il<8>(tid, g, smem); break;
        case 9: hpass16il<9>(tid, g, smem); break;
        default: hpass16il_any(tid, g, smem); break;
        }
    } else {
        switch (hq) {
        case 1: hpass16<1>(tid, g, smem); break;
        case 2: hpass16<2>(tid, g, smem); break;
        case 3: hpass16<3>(tid, g, smem); break;
        case 4: hpass16<4>(tid, g, smem); break;
        case 5: hpass16<5>(tid, g, smem); break;
        case 6: hpass16<6>(tid, g, smem); break;
        case 7: hpass16<7>(tid, g, smem); break;
        case 8: hpass16<8>(tid, g, smem); break;
        case 9: hpass16<9>(tid, g, smem); break;
        default: hpass16_any(tid, g, smem); break;
        }
    }
}

// ------------------------------------------------------------------------------------------------ phase 3: V pass + store
// Thread = CPT adjacent output columns (CPT/2 32-bit shared words per tap) of one output row; a CTA sweep covers
// 256/(tw/CPT) rows.  VMODE: 0 exact 8-bit, 1 x86-truncating 8-bit, 2 10-bit, 3 single tap 8-bit, 4 single tap 10-bit.
template <int BPS> struct StoreT { typedef uint8_t type; };
template <> struct StoreT<2> { typedef uint16_t type; };

enum { VM_EXACT8 = 0, VM_APPROX8 = 1, VM_EXACT10 = 2, VM_ONE8 = 3, VM_ONE10 = 4 };

OTT_DEV int sx_lo(uint32_t w) { return (int)(int16_t)(w & 0xffffu); }
OTT_DEV int sx_hi(uint32_t w) { return (int)w >> 16; }

// NW (2 or 4) consecutive 32-bit words with one 8/16-byte shared load
template <int NW>
OTT_DEV void load_words(const uint32_t *p, uint32_t (&w)[NW])
{
    if (NW == 4) {
        const uint4 q = *reinterpret_cast<const uint4 *>(p);
        w[0] = q.x; w[1] = q.y; w[NW - 2] = q.z; w[NW - 1] = q.w;
    } else {
        const uint2 q = *reinterpret_cast<const uint2 *>(p);
        w[0] = q.x; w[1] = q.y;
    }
}

// NV consecutive H-pass results (sign-extended) starting at p (16-byte aligned for NV = 8 with int16, NV = 4 with int32)
template <int NV>
OTT_DEV void load_vals(const mid_t *p, int (&t)[NV])
{
    if (sizeof(mid_t) == 2) {
        uint32_t ww[NV / 2];
        load_words<NV / 2>(reinterpret_cast<const uint32_t *>(p), ww);
        OTT_UNROLL
        for (int k = 0; k < NV / 2; ++k) { t[2 * k] = sx_lo(ww[k]); t[2 * k + 1] = sx_hi(ww[k]); }
    } else {
        OTT_UNROLL
        for (int v = 0; v < NV / 4; ++v) {
            uint32_t ww[4];
            load_words<4>(reinterpret_cast<const uint32_t *>(p) + 4 * v, ww);
            OTT_UNROLL
            for (int k = 0; k < 4; ++k) t[4 * v + k] = (int)ww[k];
        }
    }
}

// Packs results as they are produced so that only a few registers stay live: 8-bit -> 4 samples per word,
// 16-bit -> 2 samples per word (already shifted for P010).
template <int BPS> OTT_DEV uint32_t pack2(int a, int b, int shift)   // 16-bit: two samples
{
    return ((uint32_t)a << shift) | ((uint32_t)b << (16 + shift));
}
OTT_DEV uint32_t pack4(int a, int b, int c, int d) { return (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | ((uint32_t)d << 24); }

// One row of NV adjacent columns, `vtaps` taps (run-time count: the tap loop is unrolled by two, which keeps the kernel
// small enough for the instruction cache -- fully unrolled per-tap-count variants measured slower).
// Output: 8-bit -> NV/4 packed words, 16-bit -> NV/2 packed words.
template <int NV, int BPS>
OTT_DEV void vcolumns(const mid_t *m, int step, const int16_t *cf, int vtaps, int mode, int shift,
                      uint32_t (&out)[BPS == 1 ? NV / 4 : NV / 2])
{
    int a[NV];
    if (mode == VM_ONE8 || mode == VM_ONE10) {
        int t[NV];
        load_vals<NV>(m, t);
        OTT_UNROLL
        for (int k = 0; k < NV; ++k) a[k] = mode == VM_ONE8 ? ott_clip((t[k] + 64) >> 7, 255) : ott_clip((t[k] + 16) >> 5, 1023);
    } else if (mode == VM_APPROX8) {
        const int init = (64 + 8 * (vtaps - 1)) >> 4;
        OTT_UNROLL
        for (int k = 0; k < NV; ++k) a[k] = init;
        int j = 0;
        for (; j + 2 <= vtaps; j += 2) {                   // x86 tap counts are even (filterAlign 2): no tail in practice
            const uint32_t c2 = *reinterpret_cast<const uint32_t *>(cf + j);
            OTT_UNROLL
            for (int u = 0; u < 2; ++u) {
                const int cj = u ? sx_hi(c2) : sx_lo(c2);
                int t[NV];
                load_vals<NV>(m + (j + u) * step, t);
                OTT_UNROLL
                for (int k = 0; k < NV; ++k) a[k] += (t[k] * cj) >> 16;
            }
        }
        for (; j < vtaps; ++j) {
            const int cj = cf[j];
            int t[NV];
            load_vals<NV>(m + j * step, t);
            OTT_UNROLL
            for (int k = 0; k < NV; ++k) a[k] += (t[k] * cj) >> 16;
        }
        OTT_UNROLL
        for (int k = 0; k < NV; ++k) a[k] = ott_clip((int)(int16_t)a[k] >> 3, 255);   // paddw wraparound, psraw 3, packuswb
    } else {
        const int init = mode == VM_EXACT8 ? 64 << 12 : 1 << 16;
        OTT_UNROLL
        for (int k = 0; k < NV; ++k) a[k] = init;
        int j = 0;
        for (; j + 2 <= vtaps; j += 2) {
            OTT_UNROLL
            for (int u = 0; u < 2; ++u) {
                const int cj = cf[j + u];
                int t[NV];
                load_vals<NV>(m + (j + u) * step, t);
                OTT_UNROLL
                for (int k = 0; k < NV; ++k) a[k] += t[k] * cj;
            }
        }
        for (; j < vtaps; ++j) {
            const int cj = cf[j];
            int t[NV];
            load_vals<NV>(m + j * step, t);
            OTT_UNROLL
            for (int k = 0; k < NV; ++k) a[k] += t[k] * cj;
        }
        OTT_UNROLL
        for (int k = 0; k < NV; ++k) a[k] = mode == VM_EXACT8 ? ott_clip(a[k] >> 19, 255) : ott_clip(a[k] >> 17, 1023);
    }
    if (BPS == 1) {
        OTT_UNROLL
        for (int k = 0; k < NV / 4; ++k) out[k] = pack4(a[4 * k], a[4 * k + 1], a[4 * k + 2], a[4 * k + 3]);
    } else {
        OTT_UNROLL
        for (int k = 0; k < NV / 2; ++k) out[k] = pack2<2>(a[2 * k], a[2 * k + 1], shift);
    }
}

// Fast path of the common case (8-bit, x86-truncating rows, even tap count).  The accumulator is kept shifted left by
// 16: A = (A & 0xffff0000) + t*c adds floor(t*c / 65536) to the upper half (its low half is the product's remainder,
// masked away before the next tap), the 16-bit wraparound of paddw is the upper half's natural overflow, and the final
// (int16(acc) >> 3) is A >> 19.  First tap pair peeled so no accumulator initialisation moves are needed.
template <int NV>
OTT_DEV void vrow_approx8(const mid_t *m, int step, const int16_t *cf, int vtaps, uint32_t (&out)[NV / 4])
{
    const int init = ((64 + 8 * (vtaps - 1)) >> 4) << 16;
    int a[NV];
    {
        const uint32_t c2 = *reinterpret_cast<const uint32_t *>(cf);
        const int c0 = sx_lo(c2), c1 = sx_hi(c2);
        int t0[NV], t1[NV];
        load_vals<NV>(m, t0);
        load_vals<NV>(m + step, t1);
        OTT_UNROLL
        for (int k = 0; k < NV; ++k) a[k] = t0[k] * c0 + init;
        OTT_UNROLL
        for (int k = 0; k < NV; ++k) a[k] = (int)((uint32_t)a[k] & 0xffff0000u) + t1[k] * c1;
    }
    for (int j = 2; j < vtaps; j += 2) {
        const uint32_t c2 = *reinterpret_cast<const uint32_t *>(cf + j);
        const int c0 = sx_lo(c2), c1 = sx_hi(c2);
        int t0[NV], t1[NV];
        load_vals<NV>(m + j * step, t0);
        load_vals<NV>(m + (j + 1) * step, t1);
        OTT_UNROLL
        for (int k = 0; k < NV; ++k) a[k] = (int)((uint32_t)a[k] & 0xffff0000u) + t0[k] * c0;
        OTT_UNROLL
        for (int k = 0; k < NV; ++k) a[k] = (int)((uint32_t)a[k] & 0xffff0000u) + t1[k] * c1;
    }
    OTT_UNROLL
    for (int k = 0; k < NV / 4; ++k)
        out[k] = pack4(ott_clip(a[4 * k] >> 19, 255), ott_clip(a[4 * k + 1] >> 19, 255), ott_clip(a[4 * k + 2] >> 19, 255),
                       ott_clip(a[4 * k + 3] >> 19, 255));
}

// NWORDS packed 32-bit words = NWORDS*4/sizeof(S) samples starting at element `elem` of a row of `total` samples
template <typename S, int NWORDS>
OTT_DEV void store_words(uint8_t *row, int elem, const uint32_t (&w)[NWORDS], int total)
{
    constexpr int N = NWORDS * 4 / (int)sizeof(S);
    constexpr int kBytes = NWORDS * 4;                         // 4, 8, 16 or 32
    S *p = reinterpret_cast<S *>(row) + elem;
    const bool vec = elem + N <= total && (((uintptr_t)p) & (uintptr_t)((kBytes > 16 ? 16 : kBytes) - 1)) == 0;
    if (vec) {
        if (kBytes == 4) {
            *reinterpret_cast<uint32_t *>(p) = w[0];
        } else if (kBytes == 8) {
            uint2 q; q.x = w[0]; q.y = w[NWORDS > 1 ? 1 : 0];
            *reinterpret_cast<uint2 *>(p) = q;
        } else {
            OTT_UNROLL
            for (int i = 0; i < kBytes / 16; ++i) {
                uint4 q;
                q.x = w[(4 * i) % NWORDS]; q.y = w[(4 * i + 1) % NWORDS]; q.z = w[(4 * i + 2) % NWORDS]; q.w = w[(4 * i + 3) % NWORDS];
                reinterpret_cast<uint4 *>(p)[i] = q;
            }
        }
        return;
    }
    OTT_UNROLL
    for (int k = 0; k < N; ++k) {
        const uint32_t word = w[(k * (int)sizeof(S)) / 4];
        const uint32_t v = sizeof(S) == 1 ? (word >> (8 * (k & 3))) & 0xffu : (word >> (16 * (k & 1))) & 0xffffu;
        if (elem + k < total) p[k] = (S)v;
    }
}

// byte / halfword interleave of two packed streams (NV12 / P010 chroma writers)
OTT_DEV void zip8(uint32_t u, uint32_t v, uint32_t &lo, uint32_t &hi)
{
    lo = (u & 0xffu) | ((v & 0xffu) << 8) | ((u & 0xff00u) << 8) | ((v & 0xff00u) << 16);
    hi = ((u >> 16) & 0xffu) | (((v >> 16) & 0xffu) << 8) | ((u >> 24) << 16) | ((v >> 24) << 24);
}
OTT_DEV void zip16(uint32_t u, uint32_t v, uint32_t &lo, uint32_t &hi)
{
    lo = (u & 0xffffu) | (v << 16);
    hi = (u >> 16) | (v & 0xffff0000u);
}

template <int CPT, int BPS> struct VRowFast {
    static OTT_DEV void run(const mid_t *, int, const int16_t *, int, uint32_t (&)[BPS == 1 ? CPT / 4 : CPT / 2]) {}
};
template <int CPT> struct VRowFast<CPT, 1> {
    static OTT_DEV void run(const mid_t *m, int step, const int16_t *cf, int vtaps, uint32_t (&out)[CPT / 4]) { vrow_approx8<CPT>(m, step, cf, vtaps, out); }
};
template <int CPT, int BPS>
OTT_DEV void vrow_fast(const mid_t *m, int step, const int16_t *cf, int vtaps, uint32_t (&out)[BPS == 1 ? CPT / 4 : CPT / 2])
{
    VRowFast<CPT, BPS>::run(m, step, cf, vtaps, out);
}

template <int CPT, int BPS>
OTT_DEV void vpass(int tid, const TileGeom &g, unsigned char *smem)
{
    constexpr int NO = BPS == 1 ? CPT / 4 : CPT / 2;            // packed output words per plane
    typedef typename StoreT<BPS>::type S;
    // the geometry lives in shared memory: copy what the loop needs into registers once
    const PlaneTables &T = *g.T;
    const int tw_log2 = g.tw_log2, tile_w = g.tw, tile_h = g.th, x0 = g.x0, y0 = g.y0, sy0 = g.sy0, sr = g.sr;
    const int mid_pitch = g.mid_pitch, approx = g.approx_v, shift = g.dst_shift, nplanes = g.nplanes, il = g.dst_interleaved;
    const int vtaps = T.vtaps, exact_from = T.exact_from, dstW = T.dstW;
    uint8_t *const dst0 = g.dst0, *const dst1 = g.dst1;
    const int dpitch0 = g.dst_pitch0, dpitch1 = g.dst_pitch1;
    const int lpr_log2 = tw_log2 - (CPT == 8 ? 3 : 2);          // lanes per output row: tw / CPT (tw 64|32, CPT 8|4)
    const int lx = tid & ((1 << lpr_log2) - 1), ry = tid >> lpr_log2, rows_per_pass = kThreads >> lpr_log2;
    const int x = CPT * lx;
    if (x >= tile_w) return;
    const mid_t *mid_s = reinterpret_cast<const mid_t *>(smem + g.off_mid);
    const int16_t *vc_s = reinterpret_cast<const int16_t *>(smem + g.off_vc);
    const int32_t *vpos_s = reinterpret_cast<const int32_t *>(smem + g.off_vpos);
    const int gx = x0 + x;
    for (int y = ry; y < tile_h; y += rows_per_pass) {
        const int gy = y0 + y;
        const int r0 = vpos_s[y] - sy0;
        const int16_t *cf = vc_s + y * vtaps;
        int mode;
        if (BPS == 1) mode = vtaps == 1 ? VM_ONE8 : ((approx && gy < exact_from) ? VM_APPROX8 : VM_EXACT8);
        else mode = vtaps == 1 ? VM_ONE10 : VM_EXACT10;
        const bool fast = BPS == 1 && mode == VM_APPROX8 && !(vtaps & 1);
        uint32_t o0[NO];
        if (fast) vrow_fast<CPT, BPS>(mid_s + r0 * mid_pitch + x, mid_pitch, cf, vtaps, o0);
        else vcolumns<CPT, BPS>(mid_s + r0 * mid_pitch + x, mid_pitch, cf, vtaps, mode, shift, o0);
        if (nplanes == 1) {
            store_words<S, NO>(dst0 + (size_t)gy * dpitch0, gx, o0, dstW);
            continue;
        }
        uint32_t o1[NO];
        if (fast) vrow_fast<CPT, BPS>(mid_s + (sr + r0) * mid_pitch + x, mid_pitch, cf, vtaps, o1);
        else vcolumns<CPT, BPS>(mid_s + (sr + r0) * mid_pitch + x, mid_pitch, cf, vtaps, mode, shift, o1);
        if (il) {
            uint32_t z[2 * NO];
            OTT_UNROLL
            for (int k = 0; k < NO; ++k) {
                if (BPS == 1) zip8(o0[k], o1[k], z[2 * k], z[2 * k + 1]);
                else zip16(o0[k], o1[k], z[2 * k], z[2 * k + 1]);
            }
            store_words<S, 2 * NO>(dst0 + (size_t)gy * dpitch0, 2 * gx, z, 2 * dstW);
        } else {
            store_words<S, NO>(dst0 + (size_t)gy * dpitch0, gx, o0, dstW);
            store_words<S, NO>(dst1 + (size_t)gy * dpitch1, gx, o1, dstW);
        }
    }
}

// Columns per thread (8 or 4) come from the plan: the planner picks the width whose (row, column-group) unit count
// divides most evenly among the compute threads.
OTT_DEV void phase_vpass(int tid, const TileGeom &g, unsigned char *smem)
{
    const bool wide = g.T->cpt == 8;
    if (g.bps == 1) {
        if (wide) vpass<8, 1>(tid, g, smem);
        else vpass<4, 1>(tid, g, smem);
    } else {
        if (wide) vpass<8, 2>(tid, g, smem);
        else vpass<4, 2>(tid, g, smem);
    }
}

// ------------------------------------------------------------------------------------------------ copy items
// Same-size rung: rows [y0, y0+rows) of one plane (luma) or of the chroma planes.
OTT_DEV void copy_rows(int tid, const uint8_t *src, int spitch, uint8_t *dst, int dpitch, int row_bytes, int y0, int rows)
{
    const bool vec = ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)spitch | (uintptr_t)dpitch) & 15) == 0);
    if (vec) {
        const int nvec = row_bytes >> 4;
        const int total = rows * nvec;
        // flat index over (row, vector): consecutive threads copy consecutive 16-byte vectors of a row
        for (int i = tid; i < total; i += kThreads) {
            const int r = i / nvec, v = i - r * nvec;
            const uint4 *s = reinterpret_cast<const uint4 *>(src + (size_t)(y0 + r) * spitch);
            uint4 *d = reinterpret_cast<uint4 *>(dst + (size_t)(y0 + r) * dpitch);
            d[v] = ott_ldg(s + v);
        }
        const int tail = row_bytes - (nvec << 4);
        for (int i = tid; i < rows * tail; i += kThreads) {
            const int r = i / tail, b = (nvec << 4) + (i - r * tail);
            dst[(size_t)(y0 + r) * dpitch + b] = ott_ldg(src + (size_t)(y0 + r) * spitch + b);
        }
    } else {
        for (int r = 0; r < rows; ++r)
            for (int i = tid; i < row_bytes; i += kThreads)
                dst[(size_t)(y0 + r) * dpitch + i] = ott_ldg(src + (size_t)(y0 + r) * spitch + i);
    }
}

// I420 chroma -> NV12 interleaved (byte-wise U,V interleave of transvideo.c:543-557), 8 chroma samples per thread step
OTT_DEV void interleave_rows8(int tid, const uint8_t *u, const uint8_t *v, int spitch, uint8_t *dst, int dpitch, int cw,
                              int y0, int rows)
{
    const bool vec = ((((uintptr_t)u | (uintptr_t)v | (uintptr_t)spitch) & 7) == 0) &&
                     ((((uintptr_t)dst | (uintptr_t)dpitch) & 15) == 0);
    const int n8 = vec ? cw >> 3 : 0;
    if (n8) {
        const int total = rows * n8;
        for (int i = tid; i < total; i += kThreads) {
            const int r = i / n8, k = i - r * n8;
            const uint2 a = ott_ldg(reinterpret_cast<const uint2 *>(u + (size_t)(y0 + r) * spitch) + k);
            const uint2 b = ott_ldg(reinterpret_cast<const uint2 *>(v + (size_t)(y0 + r) * spitch) + k);
            uint4 q;
            q.x = (a.x & 0xffu) | ((b.x & 0xffu) << 8) | ((a.x & 0xff00u) << 8) | ((b.x & 0xff00u) << 16);
            q.y = ((a.x >> 16) & 0xffu) | (((b.x >> 16) & 0xffu) << 8) | ((a.x >> 24) << 16) | ((b.x >> 24) << 24);
            q.z = (a.y & 0xffu) | ((b.y & 0xffu) << 8) | ((a.y & 0xff00u) << 8) | ((b.y & 0xff00u) << 16);
            q.w = ((a.y >> 16) & 0xffu) | (((b.y >> 16) & 0xffu) << 8) | ((a.y >> 24) << 16) | ((b.y >> 24) << 24);
            reinterpret_cast<uint4 *>(dst + (size_t)(y0 + r) * dpitch)[k] = q;
        }
    }
    const int done = n8 << 3, tail = cw - done;
    for (int i = tid; i < rows * tail; i += kThreads) {
        const int r = i / tail, x = done + (i - r * tail);
        uint8_t *d = dst + (size_t)(y0 + r) * dpitch;
        d[2 * x] = ott_ldg(u + (size_t)(y0 + r) * spitch + x);
        d[2 * x + 1] = ott_ldg(v + (size_t)(y0 + r) * spitch + x);
    }
}

// ------------------------------------------------------------------------------------------------ item dispatch
// Resolves one work item.  Returns 0: nothing to do, 1: scale tile (geometry filled), 2: copy item.
// off_src: byte offset of the source buffer this tile stages into; the intermediate region(s) and the per-slot tables
// follow the source buffer(s).
OTT_DEV int item_prepare(const Item &it, const Binding &b, const LadderLaunch &L, int buf, int mid_buf, int table_slot, TileGeom &g)
{
    if (!b.active || ((b.skip_mask >> it.slot) & 1u)) return 0;
    const PlanDev &plan = *b.plan;
    if (it.slot == plan.thumb_slot && !b.thumb_active) return 0;
    if (it.kind == ITEM_COPY_LUMA || it.kind == ITEM_COPY_CHROMA) return 2;
    const RungDev &R = plan.rung[it.slot];
    const bool luma = it.kind == ITEM_SCALE_LUMA;
    const PlaneTables &T = luma ? R.luma : R.chroma;
    const int src_il = plan.src_fmt == 2 /*P010*/, dst_il = R.fmt == 1 /*NV12*/ || R.fmt == 2 /*P010*/;
    g.T = &T;
    g.bps = plan.bps;
    g.nplanes = luma ? 1 : 2;
    g.x0 = it.x0; g.y0 = it.y0; g.tw = it.tw; g.th = it.th;
    g.sy0 = it.sy0; g.sr = it.sr; g.sx0 = it.sx0; g.nvec = it.nvec;
    g.spw = T.sw_words;
    g.mid_pitch = T.tw;
    g.tw_log2 = T.tw == 64 ? 6 : 5;
    g.approx_v = R.approx_v;
    g.off_src = buf * L.src_max;
    const int mid_base = (L.double_buffer ? 2 : 1) * L.src_max;
    g.off_mid = mid_base + mid_buf * L.mid_max;
    g.off_vc = mid_base + L.mid_buffers * L.mid_max + table_slot * L.vc_max;
    g.off_vpos = g.off_vc + ((T.th * T.vtaps * 2 + 3) & ~3);
    g.src_shift = plan.src_fmt == 2 ? 6 : 0;
    g.dst_shift = R.fmt == 2 ? 6 : 0;
    if (luma) {
        g.src0 = b.src[0]; g.src_pitch0 = b.src_pitch[0];
        g.src1 = nullptr;  g.src_pitch1 = 0;
        g.src_interleaved = 0;
        g.dst0 = b.dst[it.slot][0]; g.dst_pitch0 = b.dst_pitch[it.slot][0];
        g.dst1 = nullptr; g.dst_pitch1 = 0;
        g.dst_interleaved = 0;
    } else {
        g.src0 = b.src[1]; g.src_pitch0 = b.src_pitch[1];
        g.src1 = src_il ? b.src[1] : b.src[2]; g.src_pitch1 = src_il ? b.src_pitch[1] : b.src_pitch[2];
        g.src_interleaved = src_il;
        g.dst0 = b.dst[it.slot][1]; g.dst_pitch0 = b.dst_pitch[it.slot][1];
        g.dst1 = dst_il ? nullptr : b.dst[it.slot][2]; g.dst_pitch1 = dst_il ? 0 : b.dst_pitch[it.slot][2];
        g.dst_interleaved = dst_il;
    }
    g.staged_planes = (luma || src_il) ? 1 : 2;
    g.px_bytes = (!luma && src_il) ? 4 : plan.bps;
    g.tma_x = g.sx0 * g.px_bytes / 4;
    // one staged plane occupies sr_max rows of spw words, padded to the 128 bytes a TMA destination must be aligned to
    g.plane_stride = (T.sr_max * T.sw_words * 4 + 127) & ~127;
    // TMA staging: the host supplied tensor maps for this frame (16-byte aligned planes) and the tile window fits one
    // box; otherwise the compute warps stage the window through registers
    g.bulk_ok = T.tma_ok && b.tmaps != nullptr;
    g.box_bytes = T.sr_max * T.sw_words * 4;
    g.tmap0 = b.tmaps + (size_t)(it.slot * 3 + (luma ? 0 : 1)) * kTensorMapBytes;
    g.tmap1 = b.tmaps + (size_t)(it.slot * 3 + 2) * kTensorMapBytes;
    return 1;
}

OTT_DEV void item_copy(int tid, const Item &it, const Binding &b)
{
    const PlanDev &plan = *b.plan;
    const RungDev &R = plan.rung[it.slot];
    const int src_il = plan.src_fmt == 2, dst_il = R.fmt == 1 || R.fmt == 2;
    if (it.kind == ITEM_COPY_LUMA) {
        const PlaneTables &T = R.luma;
        const int y0 = it.y0, rows = it.th;
        copy_rows(tid, b.src[0], b.src_pitch[0], b.dst[it.slot][0], b.dst_pitch[it.slot][0], T.dstW * plan.bps, y0, rows);
        return;
    }
    const PlaneTables &T = R.chroma;
    const int y0 = it.y0, rows = it.th;
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
