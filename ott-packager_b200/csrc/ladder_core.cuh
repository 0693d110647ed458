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
#include <cmath>
#define OTT_DEV inline
#define OTT_COLD inline
#define OTT_NOUNROLL
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
inline int ott_sat(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
inline uint32_t ott_pack_s16(int hi, int lo) { return ((uint32_t)ott_sat(hi, -32768, 32767) << 16) | ((uint32_t)ott_sat(lo, -32768, 32767) & 0xffffu); }
inline uint32_t ott_pack_u8(int b1, int b0, uint32_t upper) { return (upper << 16) | ((uint32_t)ott_sat(b1, 0, 255) << 8) | (uint32_t)ott_sat(b0, 0, 255); }
inline float ott_fma_rd(float a, float b, float c)                 // a*b + c rounded ONCE, toward minus infinity (FFMA.RM)
{
    const double r = (double)a * (double)b + (double)c;            // exact for the operand ranges of the V pass (< 53 bits)
    float f = (float)r;
    if ((double)f > r) f = std::nextafterf(f, -INFINITY);
    return f;
}
inline float ott_as_float(uint32_t v) { float f; std::memcpy(&f, &v, 4); return f; }
inline uint32_t ott_as_u32(float f) { uint32_t v; std::memcpy(&v, &f, 4); return v; }
inline float ott_i2f(int v) { return (float)v; }
inline long long ott_emu_vrows[2] = {0, 0};                       // test counters: fast V rows taken by the FMA path / the bias path
}  // namespace ottgpu
#else
#define OTT_DEV __device__ __forceinline__
#define OTT_COLD __device__ __forceinline__        /* rarely taken paths (out-of-line copies measured wrong on sm_100a: kept inline, behind a branch) */
#define OTT_NOUNROLL _Pragma("unroll 1")
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
// two s32 -> saturated s16 pair (hi: upper half) in one I2IP
OTT_DEV uint32_t ott_pack_s16(int hi, int lo)
{
    uint32_t d;
    asm("cvt.pack.sat.s16.s32 %0, %1, %2;" : "=r"(d) : "r"(hi), "r"(lo));
    return d;
}
// two s32 -> saturated u8 pair in bytes 1,0; the low half of `upper` moves to bytes 3,2 (one I2IP)
OTT_DEV uint32_t ott_pack_u8(int b1, int b0, uint32_t upper)
{
    uint32_t d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(b1), "r"(b0), "r"(upper));
    return d;
}
// FFMA.RM, spelled in PTX so that a build with -ftz=true / -use_fast_math cannot turn it into the flush-to-zero form: the
// vertical pass reads a DENORMAL result of this instruction as an integer (vrow_fma8)
OTT_DEV float ott_fma_rd(float a, float b, float c)
{
    float d;
    asm("fma.rm.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
OTT_DEV float ott_as_float(uint32_t v) { return __uint_as_float(v); }
OTT_DEV uint32_t ott_as_u32(float f) { return __float_as_uint(f); }
OTT_DEV float ott_i2f(int v) { return __int2float_rn(v); }
// m16n8k32 int8 MMA, A = u8 source bytes (rows), B = coefficient bytes (columns); accumulates into d
OTT_DEV void ott_mma_u8s8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
OTT_DEV void ott_mma_u8u8(int (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    asm("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
}  // namespace ottgpu
#endif

namespace ottgpu {

OTT_DEV int ott_min(int a, int b) { return a < b ? a : b; }
OTT_DEV int ott_select(bool c, int a, int b) { return c ? a : b; }
OTT_DEV int ott_max(int a, int b) { return a > b ? a : b; }
OTT_DEV int ott_clip(int v, int hi) { return v < 0 ? 0 : (v > hi ? hi : v); }

// byte / halfword interleave of two packed streams (NV12 / P010 chroma writers)
OTT_DEV void zip8(uint32_t u, uint32_t v, uint32_t &lo, uint32_t &hi)
{
    lo = ott_prmt(u, v, 0x5140);      // u0 v0 u1 v1
    hi = ott_prmt(u, v, 0x7362);      // u2 v2 u3 v3
}
OTT_DEV void zip16(uint32_t u, uint32_t v, uint32_t &lo, uint32_t &hi)
{
    lo = (u & 0xffffu) | (v << 16);
    hi = (u >> 16) | (v & 0xffff0000u);
}

// Everything one tile needs: the item's static block plus what depends on the step (thread-uniform scalars; no
// shared-memory pointers are kept in here so that the compiler keeps every shared access in the shared address space).
struct alignas(16) TileGeom : Item {
    int off_src, off_mid, off_vc, off_vpos;   // byte offsets of the regions inside the CTA's shared memory
    int bulk_ok;             // the tile's source window is fetched by the TMA engine (one box per plane)
    int src_z;               // the source frame's buffer index in its arena (third tensor-map coordinate)
    const unsigned char *tmap0, *tmap1;   // tensor maps of the plane(s)
    const uint8_t *src0, *src1;      // plane bases (chroma: U,V or the interleaved UV plane twice)
    int src_pitch0, src_pitch1;      // bytes
    uint8_t *dst0, *dst1;
    int dst_pitch0, dst_pitch1;
    mutable int mid_neg;     // set by the H pass when any intermediate of the tile is negative (the V pass's FMA path needs
                             //   them all >= 0 and falls back to the integer path for such a tile); cleared by item_prepare
};

// ------------------------------------------------------------------------------------------------ phase 0: tables
// Copies the tile's vertical coefficients and row positions into its table slot; run by `nthreads` cooperating threads
// (the producer warp in the kernel: the tables are in place before the compute warps are released).
OTT_DEV void phase_tables(int tid, int nthreads, const TileGeom &g, unsigned char *smem)
{
    int32_t *vc_s = reinterpret_cast<int32_t *>(smem + g.off_vc);
    int32_t *vpos_s = reinterpret_cast<int32_t *>(smem + g.off_vpos);
    const int n = g.th * g.vstride, th = g.th;
    for (int i = tid; i < n; i += nthreads) vc_s[i] = ott_ldg(g.vc_src + i);
    for (int i = tid; i < th; i += nthreads) vpos_s[i] = ott_ldg(g.vpos_src + i);
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

OTT_COLD void phase_stage(int tid, const TileGeom &g, unsigned char *smem)
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
    if (!shift || g.hsplit == 16) return;                 // the MMA path reads the raw bytes: only the dp2a fallback gets here
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
// 8-bit sources: the H pass is an exact banded contraction  t[row][x] = min((sum_k src[row][k] * c[x][k]) >> 7, 32767)  of
// u8 samples with s15 coefficients.  It runs on the tensor cores as int8 MMAs (IMMA.16832): M = 16 staged source rows,
// N = 8 adjacent output columns, K = 32 source bytes per step starting at the group's 4-byte aligned window base, with
// the coefficient split into two bytes, c = (hi << split) + lo, hi signed, lo unsigned:
//     sum src*c = (H << split) + L,   H = sum src*hi  (u8 x s8 MMA),   L = sum src*lo >= 0  (u8 x u8 MMA)
//     (sum >> 7) = (H << (split-7)) + (L >> 7)      exactly, because (H << split) is a multiple of 128 and L >= 0.
// The B fragments come ready-made from the plan (PlaneTables::hmma); a warp keeps them in registers while it walks down
// the rows of its column group.  A fragments are four 32-bit shared loads per K step (rows g / g+8, bytes 4t.. and
// 16+4t..; the staged row pitch is an odd multiple of 16 bytes, so a warp's 8 rows x 16 bytes hit 32 distinct banks).
// Results leave as saturated s16 pairs (the lower clamp is never reached: the plan rejects filters that could get there).
#if defined(OTTGPU_EMULATE)
// the MMA of one lane, evaluated from the shared rows and the whole fragment table (emulation runs lane by lane)
inline void hmma_lane(const uint32_t *row_g, const uint32_t *row_g8, const uint32_t *frag32x4, int lane, int (&H)[4], int (&L)[4])
{
    const int t = lane & 3;
    for (int i = 0; i < 4; ++i) {
        const uint32_t *row = (i & 2) ? row_g8 : row_g;          // rows are passed already offset by +t words
        const int n = 2 * t + (i & 1);
        for (int k = 0; k < 32; ++k) {
            const int src = (int)((row[(k >> 2) - t] >> (8 * (k & 3))) & 255u);
            const uint32_t *f = frag32x4 + (size_t)(4 * n + ((k & 15) >> 2)) * 4;
            const int r = k >> 4, b = k & 3;
            H[i] += src * (int)(int8_t)((f[r] >> (8 * b)) & 255u);
            L[i] += src * (int)((f[2 + r] >> (8 * b)) & 255u);
        }
    }
}
#endif

template <int KS, int SPLIT>
OTT_DEV void hpass8_mma(int tid, const TileGeom &g, unsigned char *smem)
{
    const int warp = tid >> 5, lane = tid & 31;
    const int ks_n = KS ? KS : g.hks;
    const int mt_plane = (g.sr + 15) >> 4, mt_total = g.nplanes * mt_plane;
    // (column group, 16-row group) pairs, column-group major, dealt out in contiguous runs so that a warp re-reads its
    // B fragments at most twice; the runs were cut by the producer (item_hruns)
    const uint32_t hr = g.hrun[warp];
    int left = (int)(hr >> 20);
    if (left == 0) return;
    const int spw = g.spw, pitch = g.mid_pitch;
    const int lane_src = (lane >> 2) * spw + (lane & 3), lane_mid = (lane >> 2) * pitch + 2 * (lane & 3);
    const uint32_t *src_s = reinterpret_cast<const uint32_t *>(smem + g.off_src) + lane_src;
    mid_t *mid_s = reinterpret_cast<mid_t *>(smem + g.off_mid) + lane_mid;
    int nt = (int)(hr & 255u), mt = (int)((hr >> 8) & 4095u);
    // OR of every packed result pair: bits 15 / 31 = "one was negative".  Tiles with fewer than 16 staged rows compute rows
    // beyond the window (stale bytes): they take the integer V path unconditionally.
    const int sr_ = g.sr;
    uint32_t signs = ((sr_ & 15) && sr_ < 16) ? 0x8000u : 0u;
    const int last_mt = ((sr_ & 15) && sr_ >= 16) ? mt_plane - 1 : mt_plane, last_row0 = sr_ - 16;
    while (left > 0) {
        const int kw = g.kw[nt];                                  // window base as a word offset into the staged rows
        const uint32_t *frag = g.hmma + (size_t)nt * ks_n * 128;
#if !defined(OTTGPU_EMULATE)
        uint4 B[KS ? KS : 1];
        if (KS) {
            OTT_UNROLL
#if defined(OTTGPU_EXPERIMENT_NO_BLOAD)      /* timing experiment only (wrong results): what do the fragment loads from L2 cost? */
            for (int ks = 0; ks < KS; ++ks) { B[ks].x = lane + nt; B[ks].y = lane ^ nt; B[ks].z = lane * 3 + nt; B[ks].w = lane + ks; }
#else
            for (int ks = 0; ks < KS; ++ks) B[ks] = ott_ldg(reinterpret_cast<const uint4 *>(frag) + ks * 32 + lane);
#endif
        }
#endif
        int run = ott_min(left, mt_total - mt);                   // row groups of this column group (both planes)
        left -= run;
        while (run > 0) {
            const int p = mt >= mt_plane, m0 = mt - p * mt_plane;
            // a plane's last row group, when partial, is moved up to end at the last staged row (it recomputes rows of the
            // group above, same values): no result is ever derived from bytes beyond the staged window, which keeps the
            // sign flag below exact
            int cnt = ott_min(run, (m0 < last_mt ? last_mt : mt_plane) - m0);   // ... inside one plane
            run -= cnt;
            mt += cnt;
            const int row0 = m0 < last_mt ? m0 << 4 : last_row0;
            const uint32_t *rg = src_s + p * (g.plane_stride >> 2) + row0 * spw + kw;
            const uint32_t *rg8 = rg + 8 * spw;
            uint32_t *o = reinterpret_cast<uint32_t *>(mid_s + (p * g.mid_rows + row0) * pitch + 8 * nt);
            uint32_t *o8 = o + 4 * pitch;                          // 8 rows down: 8 * pitch elements = 4 * pitch words
            OTT_NOUNROLL
            for (; cnt > 0; --cnt, rg += 16 * spw, rg8 += 16 * spw, o += 8 * pitch, o8 += 8 * pitch) {
                int H[4] = {0, 0, 0, 0}, L[4] = {0, 0, 0, 0};
                if (KS) {
                    OTT_UNROLL
                    for (int ks = 0; ks < (KS ? KS : 1); ++ks) {
#if defined(OTTGPU_EMULATE)
                        hmma_lane(rg + 8 * ks, rg8 + 8 * ks, frag + (size_t)ks * 128, lane, H, L);
#else
                        const uint32_t a[4] = {rg[8 * ks], rg8[8 * ks], rg[8 * ks + 4], rg8[8 * ks + 4]};
                        ott_mma_u8s8(H, a, B[ks].x, B[ks].y);
                        ott_mma_u8u8(L, a, B[ks].z, B[ks].w);
#endif
                    }
                } else {
                    for (int ks = 0; ks < ks_n; ++ks) {         // long filters: fragments re-read per step (L1-resident)
#if defined(OTTGPU_EMULATE)
                        hmma_lane(rg + 8 * ks, rg8 + 8 * ks, frag + (size_t)ks * 128, lane, H, L);
#else
                        const uint4 b = ott_ldg(reinterpret_cast<const uint4 *>(frag) + ks * 32 + lane);
                        const uint32_t a[4] = {rg[8 * ks], rg8[8 * ks], rg[8 * ks + 4], rg8[8 * ks + 4]};
                        ott_mma_u8s8(H, a, b.x, b.y);
                        ott_mma_u8u8(L, a, b.z, b.w);
#endif
                    }
                }
                // D fragment: lane (g,t) holds rows g and g+8, columns 2t and 2t+1; (sum >> 7) = (H << (split-7)) + (L >> 7)
                int v[4];
                OTT_UNROLL
                for (int i = 0; i < 4; ++i) v[i] = (SPLIT == 7 ? H[i] : 2 * H[i]) + (int)((uint32_t)L[i] >> 7);
                const uint32_t w0 = ott_pack_s16(v[1], v[0]), w8 = ott_pack_s16(v[3], v[2]);
                o[0] = w0;
                o8[0] = w8;
                signs |= w0 | w8;
            }
        }
        mt = 0;
        ++nt;
    }
    if (signs & 0x80008000u) g.mid_neg = 1;                       // needs an overshoot below level 0: rare in video, see vrow_fma8
}

// STRIP mode (device_types.h): compute warp w owns column group w of the tile -- its B fragments stay in registers while it
// walks down ALL staged rows of the plane(s) in 16-row groups; nothing is dealt out across the CTA and the results land in
// the 8 columns of the intermediate region that only this warp will read in the V pass.
template <int KS, int SPLIT>
OTT_DEV void hpass8_strip(int tid, const TileGeom &g, unsigned char *smem)
{
    const int warp = tid >> 5, lane = tid & 31;
    if (8 * warp >= g.tw) return;
    const int ks_n = KS ? KS : g.hks;
    const int spw = g.spw, pitch = g.mid_pitch;
    const uint32_t *frag = g.hmma + (size_t)warp * ks_n * 128;
#if !defined(OTTGPU_EMULATE)
    uint4 B[KS ? KS : 1];
    if (KS) {
        OTT_UNROLL
        for (int ks = 0; ks < KS; ++ks) B[ks] = ott_ldg(reinterpret_cast<const uint4 *>(frag) + ks * 32 + lane);
    }
#endif
    const int mt_plane = (g.sr + 15) >> 4;
    const uint32_t *src_s = reinterpret_cast<const uint32_t *>(smem + g.off_src) + (lane >> 2) * spw + (lane & 3) + g.kw[warp];
    uint32_t *mid_s = reinterpret_cast<uint32_t *>(reinterpret_cast<mid_t *>(smem + g.off_mid) + (lane >> 2) * pitch + 2 * (lane & 3) + 8 * warp);
    for (int p = 0; p < g.nplanes; ++p) {
        const uint32_t *rg = src_s + p * (g.plane_stride >> 2), *rg8 = rg + 8 * spw;
        uint32_t *o = mid_s + p * ((g.mid_rows * pitch) >> 1), *o8 = o + 4 * pitch;      // pitch is even: element offsets halve to words
        OTT_NOUNROLL
        for (int cnt = mt_plane; cnt > 0; --cnt, rg += 16 * spw, rg8 += 16 * spw, o += 8 * pitch, o8 += 8 * pitch) {
            int H[4] = {0, 0, 0, 0}, L[4] = {0, 0, 0, 0};
            if (KS) {
                OTT_UNROLL
                for (int ks = 0; ks < (KS ? KS : 1); ++ks) {
#if defined(OTTGPU_EMULATE)
                    hmma_lane(rg + 8 * ks, rg8 + 8 * ks, frag + (size_t)ks * 128, lane, H, L);
#else
                    const uint32_t a[4] = {rg[8 * ks], rg8[8 * ks], rg[8 * ks + 4], rg8[8 * ks + 4]};
                    ott_mma_u8s8(H, a, B[ks].x, B[ks].y);
                    ott_mma_u8u8(L, a, B[ks].z, B[ks].w);
#endif
                }
            } else {
                for (int ks = 0; ks < ks_n; ++ks) {             // long filters: fragments re-read per step (L1-resident)
#if defined(OTTGPU_EMULATE)
                    hmma_lane(rg + 8 * ks, rg8 + 8 * ks, frag + (size_t)ks * 128, lane, H, L);
#else
                    const uint4 b = ott_ldg(reinterpret_cast<const uint4 *>(frag) + ks * 32 + lane);
                    const uint32_t a[4] = {rg[8 * ks], rg8[8 * ks], rg[8 * ks + 4], rg8[8 * ks + 4]};
                    ott_mma_u8s8(H, a, b.x, b.y);
                    ott_mma_u8u8(L, a, b.z, b.w);
#endif
                }
            }
            int v[4];
            OTT_UNROLL
            for (int i = 0; i < 4; ++i) v[i] = (SPLIT == 7 ? H[i] : 2 * H[i]) + (int)((uint32_t)L[i] >> 7);
            o[0] = ott_pack_s16(v[1], v[0]);
            o8[0] = ott_pack_s16(v[3], v[2]);
        }
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
        mid_t *m = reinterpret_cast<mid_t *>(smem + g.off_mid) + col + (p * g.mid_rows + grp) * g.mid_pitch;
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
        mid_s[(p * g.mid_rows + r) * mp] = (mid_t)hdot16il<HQ>(src_s + r * spw, p, hi, lo);
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
        mid_s[(p * g.mid_rows + r) * mp] = (mid_t)ott_min(acc >> 9, 32767);
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
        mid_s[(p * g.mid_rows + row - p * sr) * g.mid_pitch] = (mid_t)ott_min(acc >> 9, 32767);
    }
}

// 16-bit sources on the tensor cores: four int8 MMAs per K step over the RAW staged bytes.  A little-endian 16-bit sample
// is already the byte pair (b0, b1), raw = 256*b1 + b0, so nothing has to be rewritten in shared memory (round 2 first
// split every staged sample at bit 7 in a fix-up pass: 27 % of the 16-bit kernel's instructions); P010's "<< 6" is folded
// into the final shift (raw = v << 6: sum raw*c = 64 * sum v*c exactly).  With c = 128*c1 + c0 (c0 < 128, c1 signed):
//     Q0 = sum b0*c0   Q1 = sum b0*c1   Q2 = sum b1*c0   Q3 = sum b1*c1      sum raw*c = 2^15 Q3 + 2^8 Q2 + 2^7 Q1 + Q0
//     (sum >> sh) = (Q3 << (15 - sh)) + ((2 Q2 + Q1 + (Q0 >> 7)) >> (sh - 7))     exactly (Q0 >= 0), sh = 9 + src_shift
// (P010 must carry zeros in its six low bits, as every decoder writes it: libswscale shifts them out, here they would count.)
// Same work distribution as the 8-bit pass; interleaved P010 chroma reads the same staged rows for both planes with a
// fragment table per plane (the byte positions of its samples differ).
#if defined(OTTGPU_EMULATE)
inline void hmma16_lane(const uint32_t *row_g, const uint32_t *row_g8, const uint32_t *frag32x8, int lane, int (&Q0)[4], int (&Q1)[4], int (&Q2)[4], int (&Q3)[4])
{
    const int t = lane & 3;
    for (int i = 0; i < 4; ++i) {
        const uint32_t *row = (i & 2) ? row_g8 : row_g;          // rows are passed already offset by +t words
        const int n = 2 * t + (i & 1);
        for (int k = 0; k < 32; ++k) {
            const int src = (int)((row[(k >> 2) - t] >> (8 * (k & 3))) & 255u);
            const uint32_t *f = frag32x8 + (size_t)(4 * n + ((k & 15) >> 2)) * 8;
            const int r = k >> 4, b = k & 3;
            Q0[i] += src * (int)(int8_t)((f[r] >> (8 * b)) & 255u);
            Q1[i] += src * (int)(int8_t)((f[2 + r] >> (8 * b)) & 255u);
            Q2[i] += src * (int)(int8_t)((f[4 + r] >> (8 * b)) & 255u);
            Q3[i] += src * (int)(int8_t)((f[6 + r] >> (8 * b)) & 255u);
        }
    }
}
#endif

template <int KS>
OTT_DEV void hpass16_mma(int tid, const TileGeom &g, unsigned char *smem)
{
    const int warp = tid >> 5, lane = tid & 31;
    const int ks_n = KS ? KS : g.hks;
    const int mt_plane = (g.sr + 15) >> 4, mt_total = g.nplanes * mt_plane;
    const uint32_t hr = g.hrun[warp];
    int left = (int)(hr >> 20);
    if (left == 0) return;
    const int spw = g.spw, pitch = g.mid_pitch;
    const int shl = 6 - g.src_shift, shr = 2 + g.src_shift;     // sh = 9 (right-aligned samples) or 15 (P010)
    const int lane_src = (lane >> 2) * spw + (lane & 3), lane_mid = (lane >> 2) * pitch + 2 * (lane & 3);
    const uint32_t *src_s = reinterpret_cast<const uint32_t *>(smem + g.off_src) + lane_src;
    mid_t *mid_s = reinterpret_cast<mid_t *>(smem + g.off_mid) + lane_mid;
    int nt = (int)(hr & 255u), mt = (int)((hr >> 8) & 4095u);
    // interleaved chroma: both planes read the one staged plane of (U,V) words
    const int src_plane_words = g.src_interleaved ? 0 : (g.plane_stride >> 2);
    while (left > 0) {
        const int kw = g.kw[nt];
        int run = ott_min(left, mt_total - mt);
        left -= run;
        while (run > 0) {
            const int p = mt >= mt_plane, m0 = mt - p * mt_plane;
            int cnt = ott_min(run, mt_plane - m0);
            run -= cnt;
            mt += cnt;
            const uint32_t *frag = (p ? g.hmma2 : g.hmma) + (size_t)nt * ks_n * 256;
#if !defined(OTTGPU_EMULATE)
            uint4 Ba[KS ? KS : 1], Bb[KS ? KS : 1];
            if (KS) {
                OTT_UNROLL
                for (int ks = 0; ks < KS; ++ks) {
                    Ba[ks] = ott_ldg(reinterpret_cast<const uint4 *>(frag) + (ks * 32 + lane) * 2);
                    Bb[ks] = ott_ldg(reinterpret_cast<const uint4 *>(frag) + (ks * 32 + lane) * 2 + 1);
                }
            }
#endif
            const uint32_t *rg = src_s + p * src_plane_words + (m0 << 4) * spw + kw;
            const uint32_t *rg8 = rg + 8 * spw;
            uint32_t *o = reinterpret_cast<uint32_t *>(mid_s + (p * g.mid_rows + (m0 << 4)) * pitch + 8 * nt);
            uint32_t *o8 = o + 4 * pitch;
#if !defined(OTTGPU_EMULATE)
            if (!KS) {
                // Long filters (the fragments of all K steps do not fit the register file): two row groups share every
                // fragment load, so the loads (L2 latency: the CTAs' shared memory leaves almost no L1) are halved and each
                // one is followed by eight independent MMAs instead of four.
                OTT_NOUNROLL
                for (; cnt > 1; cnt -= 2, rg += 32 * spw, rg8 += 32 * spw, o += 16 * pitch, o8 += 16 * pitch) {
                    int Q[2][4][4];
                    OTT_UNROLL
                    for (int u = 0; u < 2; ++u)
                        OTT_UNROLL
                        for (int q = 0; q < 4; ++q)
                            OTT_UNROLL
                            for (int i = 0; i < 4; ++i) Q[u][q][i] = 0;
                    const uint32_t *rh = rg + 16 * spw, *rh8 = rg8 + 16 * spw;
                    OTT_NOUNROLL
                    for (int ks = 0; ks < ks_n; ++ks) {
                        const uint4 ba = ott_ldg(reinterpret_cast<const uint4 *>(frag) + (ks * 32 + lane) * 2);
                        const uint4 bb = ott_ldg(reinterpret_cast<const uint4 *>(frag) + (ks * 32 + lane) * 2 + 1);
                        const uint32_t a0[4] = {rg[8 * ks], rg8[8 * ks], rg[8 * ks + 4], rg8[8 * ks + 4]};
                        const uint32_t a1[4] = {rh[8 * ks], rh8[8 * ks], rh[8 * ks + 4], rh8[8 * ks + 4]};
                        ott_mma_u8s8(Q[0][0], a0, ba.x, ba.y);
                        ott_mma_u8s8(Q[1][0], a1, ba.x, ba.y);
                        ott_mma_u8s8(Q[0][1], a0, ba.z, ba.w);
                        ott_mma_u8s8(Q[1][1], a1, ba.z, ba.w);
                        ott_mma_u8s8(Q[0][2], a0, bb.x, bb.y);
                        ott_mma_u8s8(Q[1][2], a1, bb.x, bb.y);
                        ott_mma_u8s8(Q[0][3], a0, bb.z, bb.w);
                        ott_mma_u8s8(Q[1][3], a1, bb.z, bb.w);
                    }
                    OTT_UNROLL
                    for (int u = 0; u < 2; ++u) {
                        int v[4];
                        OTT_UNROLL
                        for (int i = 0; i < 4; ++i)
                            v[i] = (Q[u][3][i] << shl) + ((2 * Q[u][2][i] + Q[u][1][i] + (int)((uint32_t)Q[u][0][i] >> 7)) >> shr);
                        o[u * 8 * pitch] = ott_pack_s16(v[1], v[0]);
                        o8[u * 8 * pitch] = ott_pack_s16(v[3], v[2]);
                    }
                }
            }
#endif
            OTT_NOUNROLL
            for (; cnt > 0; --cnt, rg += 16 * spw, rg8 += 16 * spw, o += 8 * pitch, o8 += 8 * pitch) {
                int Q0[4] = {0, 0, 0, 0}, Q1[4] = {0, 0, 0, 0}, Q2[4] = {0, 0, 0, 0}, Q3[4] = {0, 0, 0, 0};
                if (KS) {
                    OTT_UNROLL
                    for (int ks = 0; ks < (KS ? KS : 1); ++ks) {
#if defined(OTTGPU_EMULATE)
                        hmma16_lane(rg + 8 * ks, rg8 + 8 * ks, frag + (size_t)ks * 256, lane, Q0, Q1, Q2, Q3);
#else
                        const uint32_t a[4] = {rg[8 * ks], rg8[8 * ks], rg[8 * ks + 4], rg8[8 * ks + 4]};
                        ott_mma_u8s8(Q0, a, Ba[ks].x, Ba[ks].y);
                        ott_mma_u8s8(Q1, a, Ba[ks].z, Ba[ks].w);
                        ott_mma_u8s8(Q2, a, Bb[ks].x, Bb[ks].y);
                        ott_mma_u8s8(Q3, a, Bb[ks].z, Bb[ks].w);
#endif
                    }
                } else {
                    for (int ks = 0; ks < ks_n; ++ks) {
#if defined(OTTGPU_EMULATE)
                        hmma16_lane(rg + 8 * ks, rg8 + 8 * ks, frag + (size_t)ks * 256, lane, Q0, Q1, Q2, Q3);
#else
                        const uint4 ba = ott_ldg(reinterpret_cast<const uint4 *>(frag) + (ks * 32 + lane) * 2);
                        const uint4 bb = ott_ldg(reinterpret_cast<const uint4 *>(frag) + (ks * 32 + lane) * 2 + 1);
                        const uint32_t a[4] = {rg[8 * ks], rg8[8 * ks], rg[8 * ks + 4], rg8[8 * ks + 4]};
                        ott_mma_u8s8(Q0, a, ba.x, ba.y);
                        ott_mma_u8s8(Q1, a, ba.z, ba.w);
                        ott_mma_u8s8(Q2, a, bb.x, bb.y);
                        ott_mma_u8s8(Q3, a, bb.z, bb.w);
#endif
                    }
                }
                int v[4];
                OTT_UNROLL
                for (int i = 0; i < 4; ++i) v[i] = (Q3[i] << shl) + ((2 * Q2[i] + Q1[i] + (int)((uint32_t)Q0[i] >> 7)) >> shr);
                o[0] = ott_pack_s16(v[1], v[0]);
                o8[0] = ott_pack_s16(v[3], v[2]);
            }
        }
        mt = 0;
        ++nt;
    }
}

// BPS = bytes per sample of the launch (1 | 2): the kernel is instantiated once per sample size so that each image only
// carries its own paths (the hot loops are sensitive to instruction-cache pressure).
template <int BPS>
OTT_DEV void phase_hpass(int tid, const TileGeom &g, unsigned char *smem)
{
    const int hq = g.hq;
    if (BPS == 1) {
        const int hks = g.hks;
#if OTTGPU_STRIP
        if (g.hsplit != 7) hpass8_strip<0, 8>(tid, g, smem);     // a coefficient beyond 127*128+255 (border-folded taps): rare
        else if (hks == 1) hpass8_strip<1, 7>(tid, g, smem);
        else if (hks == 2) hpass8_strip<2, 7>(tid, g, smem);
        else hpass8_strip<0, 7>(tid, g, smem);
#else
        if (g.hsplit != 7) hpass8_mma<0, 8>(tid, g, smem);       // a coefficient beyond 127*128+255 (border-folded taps): rare
        else if (hks == 1) hpass8_mma<1, 7>(tid, g, smem);
        else if (hks == 2) hpass8_mma<2, 7>(tid, g, smem);
        else hpass8_mma<0, 7>(tid, g, smem);
#endif
    } else if (g.hsplit == 16) {
        const int hks = g.hks;
        if (hks == 1) hpass16_mma<1>(tid, g, smem);
        else if (hks == 2) hpass16_mma<2>(tid, g, smem);
        else if (hks == 3) hpass16_mma<3>(tid, g, smem);
        else hpass16_mma<0>(tid, g, smem);
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
OTT_DEV void vcolumns(const mid_t *m, int step, const int32_t *cf, int vtaps, int mode, int shift,
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
            OTT_UNROLL
            for (int u = 0; u < 2; ++u) {
                const int cj = cf[j + u];
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

// Fast path of the common case (8-bit, x86-truncating rows, even tap count):
//     out = clip_u8(int16(init + sum_j floor(t_j*c_j / 65536)) >> 3)            (pmulhw / paddw / psraw / packuswb)
// The per-tap floor cannot be contracted, so every tap-sample costs one multiply; what can be saved is the bookkeeping.
// Each product gets a bias that makes its upper half non-negative:  p = t*c + (b << 16),  b = (max|c| >> 1) + 1, so
// upper16(p) = floor(t*c / 65536) + b  lies in [0, max|c| + 1].  The upper halves of two neighbouring columns' products are
// packed into one register (one PRMT) and summed over the taps as independent 16-bit lanes of an ordinary 32-bit add -
// no lane can overflow because taps * (max|c| + 1) < 65536 (checked by the plan, PlaneTables::vbias).  The multiplies no
// longer depend on the accumulator (no serial chain), there is no masking, and the adds are three-input.  At the end the
// lanes hold sum + taps*b;  (lane - K) with K = taps*b - init, taken modulo 2^16 and shifted, is the paddw/psraw result.
template <int NV>
OTT_DEV void vrow_bias8(const mid_t *m, int step, const int32_t *cf, int vtaps, int bias16, int k16, uint32_t (&out)[NV / 4])
{
    constexpr int NW = NV / 2;
    uint32_t acc[NW];
    OTT_UNROLL
    for (int k = 0; k < NW; ++k) acc[k] = 0;
    // two taps per trip (the tap count is even on this path), plain pointer bumps: the loop stays small
    const uint32_t *m0 = reinterpret_cast<const uint32_t *>(m), *m1 = reinterpret_cast<const uint32_t *>(m + step);
    const uint2 *c2 = reinterpret_cast<const uint2 *>(cf);       // rows of an even number of 32-bit coefficients: 8-byte aligned
    OTT_NOUNROLL
    for (int j = vtaps; j > 0; j -= 2, m0 += step, m1 += step, ++c2) {     // step elements = two rows of step/2 words
        const uint2 cc = *c2;
        const int c0 = (int)cc.x, c1 = (int)cc.y;
        uint32_t w0[NW], w1[NW];
        load_words<NW>(m0, w0);
        load_words<NW>(m1, w1);
        OTT_UNROLL
        for (int k = 0; k < NW; ++k) {
            const uint32_t q0 = ott_hi16x2((uint32_t)(sx_lo(w0[k]) * c0 + bias16), (uint32_t)(sx_hi(w0[k]) * c0 + bias16));
            const uint32_t q1 = ott_hi16x2((uint32_t)(sx_lo(w1[k]) * c1 + bias16), (uint32_t)(sx_hi(w1[k]) * c1 + bias16));
            acc[k] = acc[k] + q0 + q1;
        }
    }
    int lo[NW], hi[NW];
    OTT_UNROLL
    for (int k = 0; k < NW; ++k) {
        lo[k] = (int)(acc[k] * 65536u - (uint32_t)k16) >> 19;      // low lane moved up, corrected, int16 >> 3
        hi[k] = (int)(acc[k] - (uint32_t)k16) >> 19;               // the low lane's bits fall off the shift
    }
    OTT_UNROLL
    for (int k = 0; k < NV / 4; ++k) {
        const uint32_t upper = ott_pack_u8(hi[2 * k + 1], lo[2 * k + 1], 0u);     // samples 4k+3, 4k+2
        out[k] = ott_pack_u8(hi[2 * k], lo[2 * k], upper);                        // samples 4k+1, 4k
    }
}

// Round 3: the same rows as round-down FP32 FMAs.  pmulhw is floor(t*c / 65536), and an FMA rounds ONCE: with an
// accumulator that is a multiple of U = 65536 u inside one binade whose ulp is U, and round-toward-minus-infinity,
//     fma.rm(g, c, acc) = floor_U(acc + g*c) = acc + U * floor(g*c / U)        exactly
// (g*c is exact inside the FMA: 24 x 13 bits).  The operand g = (2^23 + t) u is built from the packed int16 intermediate
// with ONE byte permute (float bits kOne23 | t, valid for t >= 0: the H pass raises TileGeom::mid_neg otherwise and the tile
// takes vrow_bias8); its 2^23 adds 128*c*U per tap, a multiple of the ulp, so the floor is untouched, and over a row that is
// 128 * vsum, taken out of the initial accumulator.  Per tap-sample: one PRMT (ALU pipe) + one FFMA (FMA pipe) - the
// multiply, the floor and the accumulation in one instruction, against IMAD + extraction + 1/2 PRMT + 1/4 IADD3 of the
// bias-lane path.  The plan checks that the accumulator cannot leave the binade (PlaneTables::vsum).
// The unit u is 2^-64 (it rides in the exponent byte of the PRMT constant), which puts the accumulator at 1.5 * 2^-25 and
// lets ONE more round-down FMA finish the row: (acc - 1.5 * 2^-25) * 2^-104 is (sum / 8) * 2^-149, a DENORMAL, whose
// round-down IS psraw 3 and whose bit pattern is the integer itself (negative sums become negative floats = negative
// integers as far as the saturating pack is concerned): no subtract, no shift.
template <int NV>
OTT_DEV void vrow_fma8(const mid_t *m, int step, const int32_t *cf, int vtaps, float acc0, uint32_t (&out)[NV / 4])
{
    constexpr int NW = NV / 2;
    constexpr uint32_t kOne23 = 0x2B000000u;                      // 2^23 * 2^-64: a float whose low mantissa bits are an integer
    float acc[NV];
    OTT_UNROLL
    for (int k = 0; k < NV; ++k) acc[k] = acc0;
    const uint32_t *m0 = reinterpret_cast<const uint32_t *>(m), *m1 = reinterpret_cast<const uint32_t *>(m + step);
    const uint2 *c2 = reinterpret_cast<const uint2 *>(cf);
    int j = vtaps;                                 // even and >= 2 on this path (PlaneTables::vbias): a do-while, so the
    OTT_NOUNROLL                                   // compiler does not build a second copy of the set-up for a zero-trip case
    do {
        const uint2 cc = *c2;
        const float c0 = ott_i2f((int)cc.x), c1 = ott_i2f((int)cc.y);
        uint32_t w0[NW], w1[NW];
        load_words<NW>(m0, w0);
        load_words<NW>(m1, w1);
        OTT_UNROLL
        for (int k = 0; k < NW; ++k) {
            acc[2 * k] = ott_fma_rd(ott_as_float(ott_prmt(w0[k], kOne23, 0x7610)), c0, acc[2 * k]);
            acc[2 * k + 1] = ott_fma_rd(ott_as_float(ott_prmt(w0[k], kOne23, 0x7632)), c0, acc[2 * k + 1]);
        }
        OTT_UNROLL
        for (int k = 0; k < NW; ++k) {
            acc[2 * k] = ott_fma_rd(ott_as_float(ott_prmt(w1[k], kOne23, 0x7610)), c1, acc[2 * k]);
            acc[2 * k + 1] = ott_fma_rd(ott_as_float(ott_prmt(w1[k], kOne23, 0x7632)), c1, acc[2 * k + 1]);
        }
        m0 += step; m1 += step; ++c2;
    } while ((j -= 2) > 0);
    // acc = 1.5 * 2^-25 + 2^-48 * (init + sum of floors)  ->  floor(sum / 8) as a denormal's bits (no paddw wrap: |sum| < 2^15)
    int r[NV];
    OTT_UNROLL
    for (int k = 0; k < NV; ++k) r[k] = (int)ott_as_u32(ott_fma_rd(acc[k], 4.930380657631324e-32f, -2.204051907791789e-39f));
    OTT_UNROLL
    for (int k = 0; k < NV / 4; ++k) {
        const uint32_t upper = ott_pack_u8(r[4 * k + 3], r[4 * k + 2], 0u);
        out[k] = ott_pack_u8(r[4 * k + 1], r[4 * k], upper);
    }
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

// Thread = 8 adjacent output columns of one output row (one 16-byte shared load per tap); a CTA sweep covers
// kThreads / (tw / 8) rows.  Chroma tiles run the row function once per plane (a real loop: one copy of the code).
// The rows of the fast path (8-bit target A, all but the bottom rows of a plane) come first in a loop of their own, so
// that loop carries no per-row mode decision; what is left (exact rows, target B, 16-bit) goes through vcolumns.
template <int BPS, int NO>
OTT_DEV void vpass_store(uint8_t *dst0, uint8_t *dst1, int dpitch0, int dpitch1, int gy, int gx, int dstW, int nplanes, int il,
                         bool vec_ok, const uint32_t (&o0)[NO], const uint32_t (&o1)[NO])
{
    typedef typename StoreT<BPS>::type S;
    if (nplanes == 1) {
        if (BPS == 1 && vec_ok) {
            if (NO == 2) {
                uint2 q; q.x = o0[0]; q.y = o0[NO - 1];
                *reinterpret_cast<uint2 *>(dst0 + (size_t)gy * dpitch0 + gx) = q;
            } else {
                *reinterpret_cast<uint32_t *>(dst0 + (size_t)gy * dpitch0 + gx) = o0[0];
            }
        } else {
            store_words<S, NO>(dst0 + (size_t)gy * dpitch0, gx, o0, dstW);
        }
    } else if (il) {
        uint32_t z[2 * NO];
        OTT_UNROLL
        for (int k = 0; k < NO; ++k) {
            if (BPS == 1) zip8(o0[k], o1[k], z[2 * k], z[2 * k + 1]);
            else zip16(o0[k], o1[k], z[2 * k], z[2 * k + 1]);
        }
        if (BPS == 1 && vec_ok) {
            if (NO == 2) {
                uint4 q; q.x = z[0]; q.y = z[1]; q.z = z[2 * NO - 2]; q.w = z[2 * NO - 1];
                *reinterpret_cast<uint4 *>(dst0 + (size_t)gy * dpitch0 + 2 * gx) = q;
            } else {
                uint2 q; q.x = z[0]; q.y = z[1];
                *reinterpret_cast<uint2 *>(dst0 + (size_t)gy * dpitch0 + 2 * gx) = q;
            }
        } else {
            store_words<S, 2 * NO>(dst0 + (size_t)gy * dpitch0, 2 * gx, z, 2 * dstW);
        }
    } else {
        store_words<S, NO>(dst0 + (size_t)gy * dpitch0, gx, o0, dstW);
        store_words<S, NO>(dst1 + (size_t)gy * dpitch1, gx, o1, dstW);
    }
}

// CPT = adjacent columns per thread.  Round-1 layout: CPT = 8, tw / 8 lanes per output row, a CTA sweep covers
// kThreads / (tw / 8) rows.  Strip mode (8-bit): compute warp w owns columns [8w, 8w + 8) of the tile -- the ones its H
// pass produced; CPT = 8: lane = output row (32 rows per sweep of the warp), CPT = 4 (short tiles of steep filters): two
// lanes per row, 16 rows per sweep.
template <int BPS, int CPT>
OTT_DEV void vpass(int tid, const TileGeom &g, unsigned char *smem)
{
    constexpr int NO = BPS == 1 ? CPT / 4 : CPT / 2;            // packed output words per plane
    // the geometry lives in shared memory: copy what the loop needs into registers once
    const int tile_w = g.tw, tile_h = g.th, y0 = g.y0, sy0 = g.sy0;
    const int mid_pitch = g.mid_pitch, plane_step = g.mid_rows * g.mid_pitch, approx = g.approx_v, shift = g.dst_shift, nplanes = g.nplanes, il = g.dst_interleaved;
#if defined(OTTGPU_EXPERIMENT_V_ONETRIP)   /* timing experiment only (wrong results): the per-row cost without the tap loop */
    const int vtaps = 2, vstride = g.vstride, exact_from = g.exact_from, dstW = g.dstW;
#else
    const int vtaps = g.vtaps, vstride = g.vstride, exact_from = g.exact_from, dstW = g.dstW;
#endif
    uint8_t *const dst0 = g.dst0, *const dst1 = g.dst1;
    const int dpitch0 = g.dst_pitch0, dpitch1 = g.dst_pitch1;
    int x, ry, rows_per_pass;
    if (BPS == 1 && OTTGPU_STRIP) {
        const int warp = tid >> 5, lane = tid & 31;
        x = 8 * warp + (CPT == 8 ? 0 : 4 * (lane & 1));
        ry = CPT == 8 ? lane : lane >> 1;
        rows_per_pass = CPT == 8 ? 32 : 16;
    } else {
        const int lpr_log2 = g.tw_log2 - 3;                      // lanes per output row: tw / 8
        x = CPT * (tid & ((1 << lpr_log2) - 1));
        ry = tid >> lpr_log2;
        rows_per_pass = kThreads >> lpr_log2;
    }
    if (x >= tile_w) return;
    const mid_t *mid_s = reinterpret_cast<const mid_t *>(smem + g.off_mid) + x;
    const int32_t *vc_s = reinterpret_cast<const int32_t *>(smem + g.off_vc);
    const int32_t *vpos_s = reinterpret_cast<const int32_t *>(smem + g.off_vpos);
    const int gx = g.x0 + x;
    int y = ry;
    if (BPS == 1 && approx && g.vbias > 0) {
        const int bias16 = g.vbias << 16, k16 = (vtaps * g.vbias - ((64 + 8 * (vtaps - 1)) >> 4)) << 16;
        const int fast_rows = ott_min(tile_h, exact_from - y0);
        // whole-unit vector stores when the unit lies inside the plane and the surface rows are aligned to the unit
        const bool vec_ok = gx + CPT <= dstW && (il ? ((((uintptr_t)dst0 | (uintptr_t)dpitch0) & (uintptr_t)(2 * CPT - 1)) == 0)
                                                    : (nplanes == 1 && (((uintptr_t)dst0 | (uintptr_t)dpitch0) & (uintptr_t)(CPT - 1)) == 0));
        // FMA path: every intermediate of the tile is >= 0 (flag raised by the H pass before the barrier)
#if defined(OTTGPU_NO_VFMA)                    /* A/B experiments: the integer bias-lane path only */
        const bool fma = false;
#else
        const bool fma = !OTTGPU_STRIP && sizeof(mid_t) == 2 && g.vsum != 0 && g.mid_neg == 0;
#endif
        // 1.5 * 2^-25 + 2^-48 * (init - 128 * vsum): exact (a multiple of the binade's ulp 2^-48)
        const float acc0 = 4.470348358154297e-08f + 3.552713678800501e-15f * (float)(((64 + 8 * (vtaps - 1)) >> 4) - 128 * (int)g.vsum);
        for (; y < fast_rows; y += rows_per_pass) {
            const int32_t *cf = vc_s + y * vstride;
            const mid_t *m = mid_s + (vpos_s[y] - sy0) * mid_pitch;
            uint32_t o0[NO], o1[NO];
            OTT_NOUNROLL
            for (int p = 0; p < nplanes; ++p, m += plane_step) {
                uint32_t t[CPT / 4];
#if defined(OTTGPU_EMULATE)
                ++ott_emu_vrows[fma ? 0 : 1];
#endif
                if (fma) vrow_fma8<CPT>(m, mid_pitch, cf, vtaps, acc0, t);
                else vrow_bias8<CPT>(m, mid_pitch, cf, vtaps, bias16, k16, t);
                OTT_UNROLL
                for (int k = 0; k < CPT / 4; ++k) {
                    if (p == 0) o0[k] = t[k];
                    else o1[k] = t[k];
                }
            }
#if defined(OTTGPU_EXPERIMENT_V_NOSTORE)   /* timing experiment only: rows computed, nothing written (the stores stay reachable) */
            if (o0[0] == 0x12345678u && o1[0] == 0x9abcdef0u)
#endif
            vpass_store<BPS, NO>(dst0, dst1, dpitch0, dpitch1, y0 + y, gx, dstW, nplanes, il, vec_ok, o0, o1);
        }
    }
    for (; y < tile_h; y += rows_per_pass) {
        const int gy = y0 + y;
        const int32_t *cf = vc_s + y * vstride;
        int mode;
        if (BPS == 1) mode = vtaps == 1 ? VM_ONE8 : ((approx && gy < exact_from) ? VM_APPROX8 : VM_EXACT8);
        else mode = vtaps == 1 ? VM_ONE10 : VM_EXACT10;
        const mid_t *m = mid_s + (vpos_s[y] - sy0) * mid_pitch;
        uint32_t o0[NO], o1[NO];
        OTT_NOUNROLL
        for (int p = 0; p < nplanes; ++p, m += plane_step) {
            uint32_t t[NO];
            vcolumns<CPT, BPS>(m, mid_pitch, cf, vtaps, mode, shift, t);
            OTT_UNROLL
            for (int k = 0; k < NO; ++k) {
                if (p == 0) o0[k] = t[k];
                else o1[k] = t[k];
            }
        }
        vpass_store<BPS, NO>(dst0, dst1, dpitch0, dpitch1, gy, gx, dstW, nplanes, il, false, o0, o1);
    }
}

template <int BPS>
OTT_DEV void phase_vpass(int tid, const TileGeom &g, unsigned char *smem)
{
    if (BPS == 1 && OTTGPU_STRIP && g.strip == 2) vpass<BPS, 4>(tid, g, smem);
    else vpass<BPS, 8>(tid, g, smem);
}

// ------------------------------------------------------------------------------------------------ copy items
// Same-size rung: rows [y0, y0+rows) of one plane (luma) or of the chroma planes.
OTT_DEV void copy_rows(int tid, const uint8_t *src, int spitch, uint8_t *dst, int dpitch, int row_bytes, int y0, int rows)
{
    const bool vec = ((((uintptr_t)src | (uintptr_t)dst | (uintptr_t)spitch | (uintptr_t)dpitch) & 15) == 0);
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int kWarps = kThreads / 32;
    const int nvec = vec ? row_bytes >> 4 : 0;
    // a warp takes a row, its lanes stride over the row's 16-byte vectors
    for (int r = warp; r < rows; r += kWarps) {
        const uint8_t *s = src + (size_t)(y0 + r) * spitch;
        uint8_t *d = dst + (size_t)(y0 + r) * dpitch;
        int v = lane;
        for (; v < nvec; v += 128) {                             // up to four vectors in flight per lane (predicated, no
            const uint4 *sv = reinterpret_cast<const uint4 *>(s) + v;   // divergence): one round trip per 2 KB of row
            uint4 *dv = reinterpret_cast<uint4 *>(d) + v;
            const bool p1 = v + 32 < nvec, p2 = v + 64 < nvec, p3 = v + 96 < nvec;
            uint4 a = ott_ldg(sv), b = a, c = a, e = a;
            if (p1) b = ott_ldg(sv + 32);
            if (p2) c = ott_ldg(sv + 64);
            if (p3) e = ott_ldg(sv + 96);
            dv[0] = a;
            if (p1) dv[32] = b;
            if (p2) dv[64] = c;
            if (p3) dv[96] = e;
        }
        for (int i = (nvec << 4) + lane; i < row_bytes; i += 32) d[i] = ott_ldg(s + i);
    }
}

// I420 chroma -> NV12 interleaved (byte-wise U,V interleave of transvideo.c:543-557), 8 chroma samples per thread step
OTT_DEV void interleave_rows8(int tid, const uint8_t *u, const uint8_t *v, int spitch, uint8_t *dst, int dpitch, int cw,
                              int y0, int rows)
{
    const bool vec = ((((uintptr_t)u | (uintptr_t)v | (uintptr_t)spitch) & 7) == 0) &&
                     ((((uintptr_t)dst | (uintptr_t)dpitch) & 15) == 0);
    const int n8 = vec ? cw >> 3 : 0;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int kWarps = kThreads / 32;
    for (int r = warp; r < rows; r += kWarps) {
        const uint8_t *su = u + (size_t)(y0 + r) * spitch, *sv = v + (size_t)(y0 + r) * spitch;
        uint8_t *d = dst + (size_t)(y0 + r) * dpitch;
        for (int k = lane; k < n8; k += 32) {
            const uint2 a = ott_ldg(reinterpret_cast<const uint2 *>(su) + k);
            const uint2 b = ott_ldg(reinterpret_cast<const uint2 *>(sv) + k);
            uint4 q;
            zip8(a.x, b.x, q.x, q.y);
            zip8(a.y, b.y, q.z, q.w);
            reinterpret_cast<uint4 *>(d)[k] = q;
        }
        for (int x = (n8 << 3) + lane; x < cw; x += 32) {
            d[2 * x] = ott_ldg(su + x);
            d[2 * x + 1] = ott_ldg(sv + x);
        }
    }
}

// ------------------------------------------------------------------------------------------------ item dispatch
// Resolves one work item.  Returns 0: nothing to do, 1: scale tile (geometry filled), 2: copy item.
// off_src: byte offset of the source buffer this tile stages into; the intermediate region(s) and the per-slot tables
// follow the source buffer(s).
// g's static block (the Item) is already in place; this adds what depends on the step.
OTT_DEV int item_prepare(const Binding &b, const LadderLaunch &L, int buf, int mid_buf, int table_slot, TileGeom &g)
{
    const int slot = g.slot;
    if (!b.active || ((b.skip_mask >> slot) & 1u)) return 0;
    if (g.kind == ITEM_COPY_LUMA || g.kind == ITEM_COPY_CHROMA) return (g.thumb && !b.thumb_active) ? 0 : 2;
    if (g.thumb && !b.thumb_active) return 0;
    const bool luma = g.kind == ITEM_SCALE_LUMA;
    g.off_src = buf * L.src_max;
    const int mid_base = (L.double_buffer ? 2 : 1) * L.src_max;
    g.off_mid = mid_base + mid_buf * L.mid_max;
    g.off_vc = mid_base + L.mid_buffers * L.mid_max + table_slot * L.vc_max;
    g.off_vpos = g.off_vc + g.vpos_off;
    if (luma) {
        g.src0 = b.src[0]; g.src_pitch0 = b.src_pitch[0];
        g.src1 = nullptr;  g.src_pitch1 = 0;
        g.dst0 = b.dst[slot][0]; g.dst_pitch0 = b.dst_pitch[slot][0];
        g.dst1 = nullptr; g.dst_pitch1 = 0;
    } else {
        const int src_il = g.src_interleaved, dst_il = g.dst_interleaved;
        g.src0 = b.src[1]; g.src_pitch0 = b.src_pitch[1];
        g.src1 = src_il ? b.src[1] : b.src[2]; g.src_pitch1 = src_il ? b.src_pitch[1] : b.src_pitch[2];
        g.dst0 = b.dst[slot][1]; g.dst_pitch0 = b.dst_pitch[slot][1];
        g.dst1 = dst_il ? nullptr : b.dst[slot][2]; g.dst_pitch1 = dst_il ? 0 : b.dst_pitch[slot][2];
    }
    // TMA staging: the host supplied tensor maps for this frame (16-byte aligned planes) and the tile window fits one
    // box; otherwise the compute warps stage the window through registers
    g.bulk_ok = g.tma_ok && b.tmaps != nullptr;
    g.src_z = b.src_z;
    g.mid_neg = 0;
    g.tmap0 = b.tmaps + (size_t)(slot * 3 + (luma ? 0 : 1)) * kTensorMapBytes;
    g.tmap1 = b.tmaps + (size_t)(slot * 3 + 2) * kTensorMapBytes;
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
