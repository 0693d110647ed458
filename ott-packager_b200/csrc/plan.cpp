#include "plan.h"
#include "kernels.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <cuda_runtime.h>

namespace ottgpu {

int fmt_bps(int fmt)
{
    switch (fmt) {
    case OTTGPU_FMT_I420: case OTTGPU_FMT_NV12: return 1;
    case OTTGPU_FMT_P010: case OTTGPU_FMT_I420P10: return 2;
    default: return 0;
    }
}

size_t packed_frame_bytes(int fmt, int w, int h)
{
    const size_t cw = (w + 1) / 2, ch = (h + 1) / 2;
    return ((size_t)w * h + 2 * cw * ch) * (size_t)fmt_bps(fmt);
}

bool rung_layout(const ottgpu_rung_cfg &r, RungLayout &out)
{
    const int bps = fmt_bps(r.fmt);
    if (!bps || r.width <= 0 || r.height <= 0) return false;
    const int cw = (r.width + 1) / 2, ch = (r.height + 1) / 2;
    const int al = r.pitch_align > 1 ? r.pitch_align : 1;
    auto up = [al](int v) { return (v + al - 1) / al * al; };
    const bool semi = r.fmt == OTTGPU_FMT_NV12 || r.fmt == OTTGPU_FMT_P010;
    out.planes = semi ? 2 : 3;
    out.pitch[0] = up(r.width * bps);
    out.offset[0] = 0;
    size_t at = (size_t)out.pitch[0] * r.height;
    if (semi) {
        out.pitch[1] = up(2 * cw * bps);
        out.offset[1] = at;
        at += (size_t)out.pitch[1] * ch;
        out.pitch[2] = 0;
        out.offset[2] = 0;
    } else {
        out.pitch[1] = out.pitch[2] = up(cw * bps);
        out.offset[1] = at;
        at += (size_t)out.pitch[1] * ch;
        out.offset[2] = at;
        at += (size_t)out.pitch[2] * ch;
    }
    out.bytes = at;
    return true;
}

PolyphaseTable plan_table(int src_n, int dst_n, int filter, int target, bool vertical)
{
    const bool exact = target == OTTGPU_TARGET_B;
    const int align = exact ? 1 : (vertical ? 2 : 4);            // x86 filterAlign, SURVEY.md A.1
    return make_polyphase(src_n, dst_n, filter == OTTGPU_FILTER_LANCZOS ? Kernel::Lanczos : Kernel::Bicubic, align,
                          vertical ? 1 << 12 : 1 << 14, exact);
}

Plan::~Plan()
{
    cudaSetDevice(gpu);
    for (void *p : allocations) cudaFree(p);
}

namespace {

// Two CTAs per SM: 1 source buffer + 2 intermediate regions + tables must stay under ~112 KB per CTA.
size_t env_kb(const char *name, size_t dflt_kb)
{
    const char *v = std::getenv(name);               // tuning knobs for profiling sessions, not part of the API
    return (v && std::atoi(v) > 0 ? (size_t)std::atoi(v) : dflt_kb) * 1024;
}
// The caps are tried from the most generous pair downwards until the launch really gets two resident CTAs per SM (the
// table regions grow with the filter length, so a fixed pair can fall off that cliff: 2160p lanczos did).  Set by get_plan()
// under its lock.  Orders measured on the B200: 8-bit sources like a 48 KB window, 16-bit ones (twice the
// bytes per staged sample) a bigger window and smaller intermediate regions.
size_t kSrcCap = 48 * 1024;                                 // the staged-source buffer
size_t kMidCap = 27 * 1024;                                 // one H-pass result region (there are two)
struct CapPair { int src_kb, mid_kb; };
const CapPair kCaps8[] = {{48, 27}, {44, 27}, {40, 26}, {36, 26}, {32, 24}, {28, 22}, {24, 20}, {20, 16}};
// strip mode (8-bit): ONE intermediate region, so it may be about twice as large
const CapPair kCapsStrip[] = {{48, 50}, {44, 50}, {40, 48}, {36, 44}, {32, 40}, {28, 36}, {24, 32}, {20, 24}};
const CapPair kCaps16[] = {{68, 18}, {64, 20}, {60, 20}, {56, 20}, {48, 20}, {40, 18}, {32, 16}, {24, 14}};
const int kMaxTileRows = (int)(env_kb("OTTGPU_MAX_TILE_ROWS", 128) / 1024);
const int kTileRounds = [] { const char *v = std::getenv("OTTGPU_TILE_ROUNDS"); return v ? std::atoi(v) : 1; }();   // 1: also try tile heights that fill whole V rounds          // keeps enough tiles per frame for load balance
constexpr size_t kSmemHardLimit = 224 * 1024; // == kMaxTileSmem (kernels.h); sm_100 allows 227 KB per CTA

template <typename T>
const T *upload(Plan &plan, const std::vector<T> &v, std::string &err)
{
    void *d = nullptr;
    if (cudaMalloc(&d, v.size() * sizeof(T) + 64) != cudaSuccess) { err = "cudaMalloc(tables) failed"; return nullptr; }   // + slack: table slices are bulk-copied in 16-byte units
    plan.allocations.push_back(d);
    if (!v.empty() && cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
        err = "cudaMemcpy(tables) failed";
        return nullptr;
    }
    return static_cast<const T *>(d);
}

struct TileRegions { size_t src, mid, vc; };


// ---- H-pass work of one scale tile dealt to the compute warps (Item::hrun): contiguous runs of (column group, 16-row group)
// pairs, column-group major.  There is no barrier between a tile's V pass and the next tile's H pass, so a warp reaches the
// H->V barrier after (its V rows of the PREVIOUS tile) + (its H pairs of this one), and the V rows are dealt by thread index:
// the low warps own the rows of the last, partial sweep.  Equal H shares therefore left the high warps waiting at the barrier
// (17 % of the warp-time samples).  The pairs go preferentially to the warps with the lighter V share; `vload` is the V share
// to assume per warp: null = this tile's own shape (the default: a CTA's consecutive tiles are mostly neighbours of one rung's
// tile row), else the plan-wide average of the main launch's tiles (OTT_HBAL_MODE=1, measured worse: 0.559 vs 0.550 ms on cfg2).
// Measured on the B200 (ms per 64 frames, cfg2 / cfg4): equal shares 0.570 / 0.444, OTT_HBAL 0.25 0.558 / 0.438, 0.5 0.554 / 0.435,
// 0.6 0.551 / 0.433, 0.75 0.554 / 0.436, 1.0 0.555 / 0.437, 1.5 0.569 / 0.451.  Costs in instructions per warp (profiles/README.md): a V row = planes * (60 + 22 * taps), an H pair =
// 2 + 24 * K steps (8-bit) / 40 + 28 * K steps (16-bit; cfg3, ms per 64 frames at half that cost: OTT_HBAL 0 6.19, 0.3 5.96,
// 0.6 6.02, 1.0 6.09, 1.5 6.27).  OTT_HBAL scales the V term (default 0.6, measured; 0 = equal shares).
static void tile_vload(const Item &it, double (&load)[kThreads / 32])
{
    constexpr int kWarps = kThreads / 32;
    const int lpr_log2 = it.tw_log2 - 3, rpp = kThreads >> lpr_log2, th = it.th;
    const double vrow = it.nplanes * (60.0 + 22.0 * it.vtaps);
    for (int w = 0; w < kWarps; ++w) {
        const int ry0 = (32 * w) >> lpr_log2;
        load[w] = vrow * (ry0 < th ? (th - ry0 + rpp - 1) / rpp : 0);
    }
}

static void deal_hruns(Item &it, const double *vload)
{
    constexpr int kWarps = kThreads / 32;
    static const double hbal = [] { const char *e = std::getenv("OTT_HBAL"); return e ? std::atof(e) : 0.6; }();
    const int groups = (it.tw + 7) >> 3, mt_total = it.nplanes * ((it.sr + 15) >> 4);
    const int total = groups * mt_total, per = (total + kWarps - 1) / kWarps;
    const double hpair = it.bps == 1 ? 2.0 + 24.0 * it.hks : 40.0 + 28.0 * it.hks;   // 16-bit: four products, fragments from L2 (measured: cfg3 best at twice the count)
    int share[kWarps];
    double load[kWarps];
    if (vload) for (int w = 0; w < kWarps; ++w) load[w] = vload[w];
    else tile_vload(it, load);
    for (int w = 0; w < kWarps; ++w) { load[w] *= hbal; share[w] = 0; }
    if (hbal <= 0.0) {
        for (int w = 0, f = 0; w < kWarps; ++w) { share[w] = std::max(0, std::min(per, total - f)); f += share[w]; }
    } else {
        for (int k = 0; k < total; ++k) {                        // greedy: the next pair to the least loaded warp
            int best = 0;
            for (int w = 1; w < kWarps; ++w)
                if (load[w] < load[best]) best = w;
            load[best] += hpair;
            ++share[best];
        }
    }
    int nt = 0, mt = 0;
    for (int w = 0; w < kWarps; ++w) {
        it.hrun[w] = (uint32_t)nt | ((uint32_t)mt << 8) | ((uint32_t)share[w] << 20);
        mt += share[w];
        while (mt_total > 0 && mt >= mt_total) { mt -= mt_total; ++nt; }
    }
}

// must match item_prepare() in ladder_core.cuh
TileRegions tile_regions(int nplanes, int staged_planes, int sr_max, int sw_words, int mid_pitch, int th, int vstride)
{
    TileRegions r;
    r.src = (size_t)staged_planes * (((size_t)sr_max * sw_words * 4 + 127) & ~(size_t)127);   // planes 128-byte aligned (TMA)
    // rows per plane rounded up to the 16-row groups of the H pass, row pitch tw + 8 (bank-conflict free, 16-byte multiple)
    r.mid = (size_t)nplanes * ((sr_max + 15) & ~15) * mid_pitch * sizeof(mid_t);
    r.vc = (size_t)th * vstride * 4 + (size_t)th * 4;                // 32-bit coefficients + row positions
    r.vc = (r.vc + 15) & ~(size_t)15;
    return r;
}

// Builds tables + tiling for one plane kind of one rung.  Host vectors are kept in `keep` only long enough to upload.
bool build_plane(Plan &plan, PlaneTables &T, int srcW, int srcH, int dstW, int dstH, int bps, int px_bytes, int nplanes, int filter,
                 int target, int exact_from, TileRegions &regions, bool &soft_out, std::vector<int32_t> &hpos_out,
                 std::vector<int32_t> &vpos_out, std::vector<int32_t> &kbase_out, std::string &err)
{
    const PolyphaseTable h = plan_table(srcW, dstW, filter, target, false);
    const PolyphaseTable v = plan_table(srcH, dstH, filter, target, true);
    const int hq = (h.taps + 3) / 4, taps4 = 4 * hq;
    std::vector<uint32_t> hi((size_t)dstW * hq), lo((size_t)dstW * hq);
    std::vector<int16_t> c16((size_t)dstW * taps4, 0);
    for (int x = 0; x < dstW; ++x)
        for (int j = 0; j < h.taps; ++j) c16[(size_t)x * taps4 + j] = h.coef[(size_t)x * h.taps + j];
    for (int x = 0; x < dstW; ++x)
        for (int q = 0; q < hq; ++q) {
            uint32_t a = 0, b = 0;
            for (int k = 0; k < 4; ++k) {
                const int c = c16[(size_t)x * taps4 + 4 * q + k];
                a |= (uint32_t)((c >> 8) & 255) << (8 * k);      // signed high byte
                b |= (uint32_t)(c & 255) << (8 * k);             // unsigned low byte
            }
            hi[(size_t)x * hq + q] = a;
            lo[(size_t)x * hq + q] = b;
        }
    // 8-bit path: the H pass is an exact u8 x s15 banded contraction, run on the tensor cores as two int8 MMAs
    // (m16n8k32, rows = 16 source rows, columns = 8 output columns, K = 32 source bytes per step) with the coefficient
    // split c = (hi << hsplit) + lo.  Per group of 8 output columns: the K window starts at the 4-byte aligned source
    // byte hkbase[group]; the B fragments are stored exactly as the 32 lanes hold them (lane = 4*column + k/4).
    std::vector<uint32_t> frags;
    std::vector<int32_t> kbase;
    int hks = 1, hsplit = 7;
    const bool mma8 = bps == 1 && px_bytes == 1;
    if (mma8) {
        const int groups = (dstW + 7) / 8;
        kbase.resize(groups);
        for (int gi = 0; gi < groups; ++gi) {
            int lo_pos = 1 << 30, hi_pos = 0;
            for (int x = 8 * gi; x < std::min(8 * gi + 8, dstW); ++x) {
                lo_pos = std::min(lo_pos, h.pos[x]);
                hi_pos = std::max(hi_pos, h.pos[x] + h.taps);
            }
            kbase[gi] = lo_pos & ~3;
            hks = std::max(hks, (hi_pos - kbase[gi] + 31) / 32);
        }
        long long worst_low = 0;
        for (int x = 0; x < dstW; ++x) {
            long long neg = 0;
            for (int j = 0; j < h.taps; ++j) {
                const int c = h.coef[(size_t)x * h.taps + j];
                if (c < -16384 || c > 127 * 128 + 255) hsplit = 8;
                if (c < 0) neg += c;
            }
            worst_low = std::min(worst_low, (255 * neg) >> 7);
        }
        if (worst_low < -32768) { err = "horizontal filter overshoots the 16-bit intermediate range"; return false; }
        frags.assign((size_t)groups * hks * 32 * 4, 0);
        for (int gi = 0; gi < groups; ++gi)
            for (int ks = 0; ks < hks; ++ks)
                for (int lane = 0; lane < 32; ++lane) {
                    const int x = 8 * gi + (lane >> 2), t = lane & 3;
                    uint32_t *f = &frags[(((size_t)gi * hks + ks) * 32 + lane) * 4];
                    if (x >= dstW) continue;
                    for (int r = 0; r < 2; ++r)
                        for (int b = 0; b < 4; ++b) {
                            const int j = kbase[gi] + 32 * ks + 16 * r + 4 * t + b - h.pos[x];
                            const int c = (j >= 0 && j < h.taps) ? h.coef[(size_t)x * h.taps + j] : 0;
                            int chi, clo;
                            if (hsplit == 7) { chi = std::min(c >> 7, 127); clo = c - 128 * chi; }
                            else { chi = c >> 8; clo = c & 255; }
                            f[r] |= (uint32_t)(chi & 255) << (8 * b);
                            f[2 + r] |= (uint32_t)(clo & 255) << (8 * b);
                        }
                }
    }
    // 16-bit path (10 significant bits in a 16-bit container): the contraction runs as four int8 MMAs over the RAW staged
    // bytes -- a little-endian sample is the byte pair (b0, b1) -- with the coefficient split at bit 7, c = 128*c1 + c0
    // (c0 < 128, c1 signed):
    //     Q0 = sum b0*c0    Q1 = sum b0*c1    Q2 = sum b1*c0    Q3 = sum b1*c1     sum raw*c = 2^15 Q3 + 2^8 Q2 + 2^7 Q1 + Q0
    // Fragments per (column group, K step, lane): {Q0.r0, Q0.r1, Q1.r0, Q1.r1, Q2.r0, Q2.r1, Q3.r0, Q3.r1}; interleaved chroma
    // (P010: U.b0 U.b1 V.b0 V.b1) gets one table per plane (the non-zero byte positions differ).  hsplit = 16 marks the plane
    // as running this path.
    std::vector<uint32_t> frags_v;
    bool mma16 = false;
    if (bps == 2) {
        mma16 = true;
        for (int16_t c : h.coef)
            if (c < -16384 || c > 16383) mma16 = false;          // c1 must fit a signed byte: else the dp2a path
    }
    if (mma16) {
        const int groups = (dstW + 7) / 8, planes = px_bytes == 4 ? 2 : 1;
        kbase.resize(groups);
        for (int gi = 0; gi < groups; ++gi) {
            int lo_pos = 1 << 30, hi_pos = 0;
            for (int x = 8 * gi; x < std::min(8 * gi + 8, dstW); ++x) {
                lo_pos = std::min(lo_pos, h.pos[x]);
                hi_pos = std::max(hi_pos, h.pos[x] + h.taps);
            }
            kbase[gi] = (lo_pos * px_bytes) & ~3;                // BYTES from the start of the staged row's plane
            hks = std::max(hks, (hi_pos * px_bytes - kbase[gi] + 31) / 32);
        }
        hsplit = 16;
        for (int pl = 0; pl < planes; ++pl) {
            std::vector<uint32_t> &out = pl ? frags_v : frags;
            out.assign((size_t)groups * hks * 32 * 8, 0);
            for (int gi = 0; gi < groups; ++gi)
                for (int ks = 0; ks < hks; ++ks)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int x = 8 * gi + (lane >> 2), t = lane & 3;
                        uint32_t *f = &out[(((size_t)gi * hks + ks) * 32 + lane) * 8];
                        if (x >= dstW) continue;
                        for (int r = 0; r < 2; ++r)
                            for (int b = 0; b < 4; ++b) {
                                const int pos = kbase[gi] + 32 * ks + 16 * r + 4 * t + b;      // byte in the staged row
                                const int xs = pos / px_bytes, w = pos % px_bytes;
                                const int lo_byte = px_bytes == 4 ? 2 * pl : 0;                // where this plane's s0 sits
                                const bool is_b0 = w == lo_byte, is_b1 = w == lo_byte + 1;
                                const int j = xs - h.pos[x];
                                const int c = (j >= 0 && j < h.taps) ? h.coef[(size_t)x * h.taps + j] : 0;
                                const int c0 = c & 127, c1 = c >> 7;
                                f[r] |= (uint32_t)((is_b0 ? c0 : 0) & 255) << (8 * b);
                                f[2 + r] |= (uint32_t)((is_b0 ? c1 : 0) & 255) << (8 * b);
                                f[4 + r] |= (uint32_t)((is_b1 ? c0 : 0) & 255) << (8 * b);
                                f[6 + r] |= (uint32_t)((is_b1 ? c1 : 0) & 255) << (8 * b);
                            }
                    }
        }
    }
    T.srcW = srcW; T.srcH = srcH; T.dstW = dstW; T.dstH = dstH;
    T.hq = hq;
    T.hks = hks;
    T.hsplit = hsplit;
    T.vtaps = v.taps;
    {
        int cmax = 0;
        for (int16_t c : v.coef) cmax = std::max(cmax, std::abs((int)c));
        T.vbias = (v.taps % 2 == 0 && (long long)v.taps * (cmax + 2) < 65536) ? (cmax >> 1) + 1 : 0;
        // FMA path of the same rows (vrow_fma8): every row's coefficients must add up to the same value (libswscale
        // normalises them to 4096) and 129 * sum|c| must stay far inside the accumulator's binade (2^22 steps either way)
        long long sum0 = 0, amax = 0;
        bool same = dstH > 0 && v.taps > 0;
        for (int y = 0; y < dstH && same; ++y) {
            long long sum = 0, asum = 0;
            for (int j = 0; j < v.taps; ++j) { const int c = v.coef[(size_t)y * v.taps + j]; sum += c; asum += std::abs(c); }
            if (y == 0) sum0 = sum;
            same = sum == sum0;
            amax = std::max(amax, asum);
        }
        T.vsum = (T.vbias > 0 && same && sum0 > 0 && sum0 < 65536 && 129 * amax + 4096 < (1ll << 21)) ? (int)sum0 : 0;
    }
    T.exact_from = exact_from;
    const bool strip = OTTGPU_STRIP && mma8;
    // strip mode reads the V coefficients of 32 different rows at once (lane = row), 8 bytes per lane: a row stride of an odd
    // number of 8-byte units keeps those loads free of bank conflicts
    T.vstride = (strip && v.taps % 2 == 0 && (v.taps / 2) % 2 == 0) ? v.taps + 2 : v.taps;
    // px_bytes = bytes per staged sample position: bps for planar samples, 4 for P010's interleaved (U,V) pairs, which are
    // staged as ONE plane of 32-bit words
    const int spv = 16 / px_bytes;                               // sample positions per 16-byte staging vector
    const int slack = 4 / px_bytes > 0 ? 4 / px_bytes : 1;       // one 32-bit word of read-ahead for the funnel shift
    const int staged_planes = px_bytes == 4 ? 1 : nplanes;
    const int row_quant = 8;                                     // tile heights are multiples of 8 (a V sweep covers 32 rows)
    auto rows_needed = [&](int th) {
        int m = 0;
        for (int y0 = 0; y0 < dstH; y0 += th) {
            const int y1 = std::min(y0 + th, dstH) - 1;
            m = std::max(m, v.pos[y1] + v.taps - v.pos[y0]);
        }
        return m;
    };
    // Tile shape.  The H pass is recomputed for the vertical halo rows of every tile (tall tiles are cheap, narrow ones
    // only cost staged bytes); the V pass hands out (row, column-group) units to kThreads threads, so a tile whose unit
    // count is not a multiple of kThreads leaves threads idle at the barrier.  Candidates inside the shared-memory caps
    // are ranked by an instruction-count model of both passes per useful output sample.
    struct Shape { int tw, th, cpt, sw_words, sr_max; double cost; bool fits, soft; };
    auto staged_words = [&](int tw) {
        const int tiles_x = (dstW + tw - 1) / tw;
        int nvec = 0;
        for (int tx = 0; tx < tiles_x; ++tx) {
            const int x0 = tx * tw, x1 = std::min(x0 + tw, dstW) - 1;
            int s0 = h.pos[x0] & ~(spv - 1), s1 = h.pos[x1] + taps4 + slack;
            if (mma8 || mma16) {                                 // every column group reads its whole K window (kbase in bytes)
                int b0 = 1 << 30, b1 = 0;
                for (int gi = x0 / 8; gi <= x1 / 8; ++gi) {
                    b0 = std::min(b0, kbase[gi] & ~15);
                    b1 = std::max(b1, kbase[gi] + 32 * hks);
                }
                s0 = b0 / px_bytes;
                s1 = (b1 + px_bytes - 1) / px_bytes;
            }
            nvec = std::max(nvec, (s1 - s0 + spv - 1) / spv);
        }
        if (mma8 || mma16) nvec |= 1;                                     // row pitch = odd multiple of 16 bytes: the fragment loads
                                                                 // (8 rows x 16 bytes) then hit 32 distinct banks
        return nvec * 4;
    };
    auto model = [&](int tw, int th, int cpt, int sr) {
        const int ngroups = kThreads / tw;
        const double hrows = (double)((nplanes * sr + ngroups - 1) / ngroups) * ngroups;          // rows incl. idle groups
        const double h_instr = hrows * tw * (6.0 + 3.0 * hq);
        const int units = th * tw / cpt;
        const double vunits = (double)((units + kThreads - 1) / kThreads) * kThreads;
        const double v_instr = vunits * nplanes * cpt * (3.3 * v.taps + 12.0) + vunits * 70.0;
        const int tiles_y = (dstH + th - 1) / th;
        const double edge = (double)tiles_y * th / dstH;                                          // partial last tile row
        const double cols = (double)((dstW + tw - 1) / tw) * tw / dstW;
        return (h_instr + v_instr) / ((double)nplanes * th * tw) * edge * cols;
    };
    auto best_for = [&](int tw) {
        Shape sh{tw, row_quant, tw == 64 ? 8 : 4, staged_words(tw), 0, 1e30, false, false};
        auto fits = [&](int th, bool hard) {
            const TileRegions r = tile_regions(nplanes, staged_planes, rows_needed(th), sh.sw_words, tw + 8, th, T.vstride);
            return hard ? r.src + r.mid + r.vc + 16 <= kSmemHardLimit : (r.src <= kSrcCap && r.mid <= kMidCap);
        };
        // tallest tile inside the caps (measured better than minimising the modelled cost over all heights: per-tile
        // overheads the model does not see favour big tiles), rebalanced so the last tile row is not a sliver; then the
        // V thread width that wastes fewer thread slots
        int th = std::min((dstH + row_quant - 1) / row_quant * row_quant, kMaxTileRows);
        while (th > row_quant && !fits(th, false)) th -= row_quant;
        if (fits(th, false)) {
            const int nty = (dstH + th - 1) / th;
            th = ((dstH + nty - 1) / nty + row_quant - 1) / row_quant * row_quant;
            const int sr = rows_needed(th);
            // V thread width: 8 columns per thread (the kernel carries that unit only; tw is 32 or 64, so a row has 4 or 8 lanes)
            for (int cpt = 8; cpt >= 8; cpt -= 4) {               // the kernel carries the 8-column unit only
                if (tw / cpt < 4) continue;                      // at least 4 lanes per row
                const double c = model(tw, th, cpt, sr);
                if (c < sh.cost * (cpt == 4 ? 0.97 : 1.0)) { sh.cost = c; sh.th = th; sh.cpt = cpt; sh.sr_max = sr; sh.fits = sh.soft = true; }
            }
        }
        for (int th = row_quant / 2; !sh.soft && th >= 1; th /= 2) {      // steep filters: shorter than the quantum, still inside the caps
            if (!fits(th, false)) continue;
            sh.th = th; sh.cpt = 8; sh.sr_max = rows_needed(th); sh.cost = model(tw, th, 8, sh.sr_max); sh.fits = sh.soft = true;
        }
        if (!sh.soft) {
            // very long filters (4K -> thumbnail): single-buffered launch, shrink until the tile fits at all
            int th = row_quant;
            while (th > 1 && !fits(th, true)) th /= 2;
            sh.fits = fits(th, true);
            sh.th = th;
            sh.sr_max = rows_needed(th);
            sh.cost = 1e29;
        }
        return sh;
    };
    Shape best;
    if (strip) {
        // Strip tiles: G column groups of 8 output columns wide (G <= 12 = one per compute warp; the plane's column groups are
        // dealt out evenly over ceil(groups / G) tiles), th rows high: multiples of 32 (V pass: lane = row) or 16 / 8 / ...
        // (two lanes per row).  Per G the tallest tile inside the caps; the winner maximises busy warps x useful H rows.
        const int groups = (dstW + 7) / 8;
        auto strip_words = [&](int G) {
            const int tiles_x = (groups + G - 1) / G;
            int nvec = 0;
            for (int tx = 0; tx < tiles_x; ++tx) {
                const int g0 = (int)((long long)tx * groups / tiles_x), g1 = (int)((long long)(tx + 1) * groups / tiles_x);
                int b0 = 1 << 30, b1 = 0;
                for (int gi = g0; gi < g1; ++gi) {
                    b0 = std::min(b0, kbase[gi] & ~15);
                    b1 = std::max(b1, kbase[gi] + 32 * hks);
                }
                nvec = std::max(nvec, (b1 - b0 + 15) / 16);
            }
            return (nvec | 1) * 4;                               // odd multiple of 16 bytes (fragment loads: 32 distinct banks)
        };
        static const int kHeights[] = {128, 96, 64, 32, 16, 8, 4, 2, 1};
        best = Shape{0, 0, 8, 0, 0, 0.0, false, false};
        for (int pass = 0; pass < 2 && !best.fits; ++pass)       // pass 0: inside the two-CTAs-per-SM caps, pass 1: anything that fits one CTA
            for (int G = std::min(kStripGroups, groups); G >= 1; --G) {
                const int sw = strip_words(G);
                if (sw > 256 && pass == 0) continue;             // one TMA box row
                for (int th : kHeights) {
                    if (th > kMaxTileRows && th > 32) continue;
                    const int nty = (dstH + th - 1) / th;
                    const int q = th >= 32 ? 32 : th;
                    const int th2 = std::min(th, ((dstH + nty - 1) / nty + q - 1) / q * q);           // rebalanced: no sliver at the bottom
                    const int sr = rows_needed(th2);
                    const TileRegions r = tile_regions(nplanes, staged_planes, sr, sw, kStripWidth + 8, th2, T.vstride);
                    const bool ok = pass ? r.src + r.mid + (size_t)kGeomSlots8 * r.vc + 16 <= kSmemHardLimit : (r.src <= kSrcCap && r.mid <= kMidCap && sr <= 256);
                    if (!ok) continue;
                    const int tiles_x = (groups + G - 1) / G;
                    const double busy = (double)groups / (tiles_x * kStripGroups);                     // compute warps with a column group
                    const double lanes = th2 >= 32 ? (double)th2 / ((th2 + 31) / 32 * 32) : (th2 >= 16 ? 1.0 : th2 / 16.0);
                    const double useful = (double)th2 * srcH / dstH / sr;   // staged rows an ideal (halo-free) tile would need / rows it stages
                    const double score = busy * lanes * std::min(1.0, useful);
                    if (score > best.cost) best = Shape{8 * G, th2, th2 >= 32 ? 8 : 4, sw, sr, score, true, pass == 0};
                    break;                                       // shorter tiles of the same width only score lower
                }
            }
    } else {
        best = best_for(dstW > 64 ? 64 : 32);
        if (best.tw == 64) {
            const Shape narrow = best_for(32);
            if (narrow.soft && (!best.soft || narrow.cost < 0.95 * best.cost)) best = narrow;
            else if (!best.fits && narrow.fits) best = narrow;
        }
    }
    if (!best.fits) { err = "filter too long for one tile (shared memory)"; return false; }
    soft_out = best.soft;
    T.tw = best.tw;
    T.mid_pitch = strip ? kStripWidth + 8 : best.tw + 8;         // strip: ONE pitch for every tile of a launch (the warps' column ranges
                                                                 // of the shared intermediate region must coincide from tile to tile)
    T.th = best.th;
    T.cpt = best.cpt;
    T.sw_words = best.sw_words;
    T.sr_max = best.sr_max;
    T.tiles_x = strip ? ((dstW + 7) / 8 + T.tw / 8 - 1) / (T.tw / 8) : (dstW + T.tw - 1) / T.tw;
    T.tiles_y = (dstH + T.th - 1) / T.th;
    // one TMA box per staged plane: inner dimension <= 256 32-bit elements, <= 256 rows, rows a whole number of words
    T.tma_ok = T.sw_words <= 256 && T.sr_max <= 256 && (srcW * px_bytes) % 4 == 0;
    regions = tile_regions(nplanes, staged_planes, T.sr_max, T.sw_words, T.mid_pitch, T.th, T.vstride);
    hpos_out = h.pos;
    vpos_out = v.pos;
    kbase_out = kbase;

    if (mma8 || mma16) { hi.clear(); lo.clear(); c16.clear(); }  // the MMA paths read only the fragment tables
    T.hpos = upload(plan, h.pos, err);
    T.hc_hi = upload(plan, hi, err);
    T.hc_lo = upload(plan, lo, err);
    T.hc16 = upload(plan, c16, err);
    T.hmma = upload(plan, frags, err);
    T.hmma2 = upload(plan, frags_v, err);
    T.hkbase = upload(plan, kbase, err);
    T.vpos = upload(plan, v.pos, err);
    {
        std::vector<int32_t> vcp((size_t)dstH * T.vstride, 0);
        for (int y = 0; y < dstH; ++y)
            for (int j = 0; j < v.taps; ++j) vcp[(size_t)y * T.vstride + j] = v.coef[(size_t)y * v.taps + j];
        T.vc = upload(plan, vcp, err);
    }
    return T.hpos && T.hc_hi && T.hc_lo && T.hc16 && T.hmma && T.hmma2 && T.hkbase && T.vpos && T.vc;
}

void copy_tiling(PlaneTables &T, int w, int h)
{
    std::memset(&T, 0, sizeof(T));
    T.srcW = T.dstW = w;
    T.srcH = T.dstH = h;
    T.tw = w;
    T.th = 32;                                                   // rows per copy item
    T.tiles_x = 1;
    T.tiles_y = (h + T.th - 1) / T.th;
}

struct Key {
    ottgpu_cfg c;
    bool operator<(const Key &o) const { return std::memcmp(&c, &o.c, sizeof(c)) < 0; }
};
std::mutex g_mu;
std::map<Key, std::weak_ptr<Plan>> g_cache;

std::shared_ptr<Plan> build_plan(const ottgpu_cfg &cfg, std::string &err);

bool compatible(int src_fmt, int dst_fmt, bool same_size)
{
    if (fmt_bps(src_fmt) != fmt_bps(dst_fmt) || !fmt_bps(src_fmt)) return false;
    if (src_fmt == OTTGPU_FMT_NV12) return false;                // sources are planar I420 / P010 / I420P10
    if (same_size) return src_fmt == dst_fmt || (src_fmt == OTTGPU_FMT_I420 && dst_fmt == OTTGPU_FMT_NV12);
    return true;
}

}  // namespace

std::shared_ptr<Plan> get_plan(const ottgpu_cfg &cfg_in, std::string &err)
{
    // normalise the key: only fields that influence tables/tiling; unused rung entries zeroed
    Key key;
    std::memset(&key, 0, sizeof(key));
    key.c.gpu = cfg_in.gpu;
    key.c.src_width = cfg_in.src_width; key.c.src_height = cfg_in.src_height; key.c.src_fmt = cfg_in.src_fmt;
    key.c.n_rungs = cfg_in.n_rungs;
    for (int i = 0; i < cfg_in.n_rungs && i < OTTGPU_MAX_RUNGS; ++i) key.c.rung[i] = cfg_in.rung[i];
    key.c.filter = cfg_in.filter; key.c.target = cfg_in.target;
    key.c.thumb_width = cfg_in.thumb_width; key.c.thumb_height = cfg_in.thumb_height;
    const ottgpu_cfg &cfg = key.c;

    if (cfg.n_rungs < 0 || cfg.n_rungs > OTTGPU_MAX_RUNGS) { err = "n_rungs out of range"; return nullptr; }
    if (cfg.src_width < 8 || cfg.src_height < 8 || cfg.src_width > 16384 || cfg.src_height > 16384) { err = "source size unsupported"; return nullptr; }
    if (cfg.src_fmt != OTTGPU_FMT_I420 && cfg.src_fmt != OTTGPU_FMT_P010 && cfg.src_fmt != OTTGPU_FMT_I420P10) { err = "source format unsupported"; return nullptr; }
    if (cfg.filter != OTTGPU_FILTER_BICUBIC && cfg.filter != OTTGPU_FILTER_LANCZOS) { err = "filter unsupported"; return nullptr; }
    if (cfg.target != OTTGPU_TARGET_A && cfg.target != OTTGPU_TARGET_B) { err = "parity target unsupported"; return nullptr; }

    std::lock_guard<std::mutex> lock(g_mu);
    auto found = g_cache.find(key);
    if (found != g_cache.end())
        if (auto p = found->second.lock()) return p;

    if (cudaSetDevice(cfg.gpu) != cudaSuccess) { err = "cudaSetDevice failed"; return nullptr; }
    // tile budgets: an explicit pair from the environment (profiling sessions), else the first pair of the list that
    // keeps two CTAs of the main launch resident per SM
    const char *env_src = std::getenv("OTTGPU_SRC_CAP_KB"), *env_mid = std::getenv("OTTGPU_MID_CAP_KB");
    std::vector<CapPair> tries;
    if (env_src && env_mid && std::atoi(env_src) > 0 && std::atoi(env_mid) > 0) tries.push_back({std::atoi(env_src), std::atoi(env_mid)});
    else if (fmt_bps(cfg.src_fmt) == 1 && OTTGPU_STRIP) tries.assign(std::begin(kCapsStrip), std::end(kCapsStrip));
    else if (fmt_bps(cfg.src_fmt) == 1) tries.assign(std::begin(kCaps8), std::end(kCaps8));
    else tries.assign(std::begin(kCaps16), std::end(kCaps16));
    std::shared_ptr<Plan> plan;
    for (size_t k = 0; k < tries.size(); ++k) {
        kSrcCap = (size_t)tries[k].src_kb * 1024;
        kMidCap = (size_t)tries[k].mid_kb * 1024;
        plan = build_plan(cfg, err);
        if (!plan) return nullptr;
        if (plan->items.empty() || k + 1 == tries.size()) break;
        if (ladder_ctas_per_sm(ladder_smem_bytes(fmt_bps(cfg.src_fmt), plan->src_max[0], plan->mid_max[0], plan->vc_max[0], nullptr, nullptr)) >= 2) break;
    }
    g_cache[key] = plan;
    return plan;
}

namespace {

std::shared_ptr<Plan> build_plan(const ottgpu_cfg &cfg, std::string &err)
{
    auto plan = std::make_shared<Plan>();
    plan->gpu = cfg.gpu;
    plan->cfg = cfg;
    const int bps = fmt_bps(cfg.src_fmt);
    const int scw = (cfg.src_width + 1) / 2, sch = (cfg.src_height + 1) / 2;
    const bool thumb = cfg.thumb_width > 0 && cfg.thumb_height > 0;
    plan->n_slots = cfg.n_rungs + (thumb ? 1 : 0);
    plan->thumb_slot = thumb ? cfg.n_rungs : -1;

    PlanDev &P = plan->host_copy;
    std::memset(&P, 0, sizeof(P));
    P.src_fmt = cfg.src_fmt;
    P.bps = bps;
    P.n_slots = plan->n_slots;
    P.thumb_slot = plan->thumb_slot;

    {
        void *d = nullptr;
        if (cudaMalloc(&d, sizeof(PlanDev)) != cudaSuccess) { err = "cudaMalloc(plan) failed"; return nullptr; }
        plan->allocations.push_back(d);
        plan->dev = static_cast<PlanDev *>(d);
    }
    struct Sortable { Item it; int band; };
    std::vector<Sortable> ordered;
    for (int s = 0; s < plan->n_slots; ++s) {
        ottgpu_rung_cfg rc;
        if (s == plan->thumb_slot) { rc.width = cfg.thumb_width; rc.height = cfg.thumb_height; rc.fmt = bps == 1 ? OTTGPU_FMT_I420 : OTTGPU_FMT_I420P10; rc.pitch_align = 0; }
        else rc = cfg.rung[s];
        if (rc.width < 2 || rc.height < 2 || rc.width > 16384 || rc.height > 16384) { err = "rung size unsupported"; return nullptr; }
        const bool same = rc.width == cfg.src_width && rc.height == cfg.src_height;
        if (!compatible(cfg.src_fmt, rc.fmt, same)) { err = "source/rung format combination unsupported"; return nullptr; }
        if (!rung_layout(rc, plan->layout[s])) { err = "bad rung layout"; return nullptr; }
        RungDev &R = P.rung[s];
        R.fmt = rc.fmt;
        R.scaled = !same;
        R.approx_v = bps == 1 && cfg.target == OTTGPU_TARGET_A;
        const int dcw = (rc.width + 1) / 2, dch = (rc.height + 1) / 2;
        const bool is_thumb = s == plan->thumb_slot;
        auto copy_item = [&](uint8_t kind, int y0, int rows) {
            Item it;
            std::memset(&it, 0, sizeof(it));
            it.slot = (uint8_t)s;
            it.kind = kind;
            it.y0 = (uint16_t)y0;
            it.th = (uint16_t)rows;
            it.bps = (uint8_t)bps;
            it.thumb = is_thumb;
            return it;
        };
        if (same) {
            copy_tiling(R.luma, cfg.src_width, cfg.src_height);
            copy_tiling(R.chroma, scw, sch);
            for (int ty = 0; ty < R.luma.tiles_y; ++ty) {
                const int y0 = ty * R.luma.th;
                ordered.push_back({copy_item(ITEM_COPY_LUMA, y0, std::min(R.luma.th, cfg.src_height - y0)), y0});
            }
            for (int ty = 0; ty < R.chroma.tiles_y; ++ty) {
                const int y0 = ty * R.chroma.th;
                ordered.push_back({copy_item(ITEM_COPY_CHROMA, y0, std::min(R.chroma.th, sch - y0)), 2 * y0});
            }
            continue;
        }
        // swscale switches to the C vertical functions once dstY >= dstH-2 (luma rows);
        // chroma row uv is emitted while dstY == 2*uv
        const int luma_exact = R.approx_v ? std::max(rc.height - 2, 0) : 0;
        const int chroma_exact = R.approx_v ? std::max((rc.height - 2 + 1) / 2, 0) : 0;
        // the thumbnail scaler is SWS_BICUBIC whatever the ladder uses (transvideo.c:2155-2157)
        const int filt = s == plan->thumb_slot ? OTTGPU_FILTER_BICUBIC : cfg.filter;
        TileRegions rl, rc2;
        bool soft_l = true, soft_c = true;
        std::vector<int32_t> lh, lv, chh, cv, lk, ck;
        if (!build_plane(*plan, R.luma, cfg.src_width, cfg.src_height, rc.width, rc.height, bps, bps, 1, filt, cfg.target, luma_exact, rl, soft_l, lh, lv, lk, err)) return nullptr;
        if (!build_plane(*plan, R.chroma, scw, sch, dcw, dch, bps, cfg.src_fmt == OTTGPU_FMT_P010 ? 4 : bps, 2, filt, cfg.target, chroma_exact, rc2, soft_c, chh, cv, ck, err)) return nullptr;
        // launch class per plane kind: 0 = inside the two-CTAs-per-SM budget, 1 = big tiles, 2 = thumbnail
        const int class_l = is_thumb ? 2 : (soft_l ? 0 : 1), class_c = is_thumb ? 2 : (soft_c ? 0 : 1);
        auto account = [&](int cls, const TileRegions &r) {
            plan->src_max[cls] = std::max(plan->src_max[cls], (int)r.src);
            plan->mid_max[cls] = std::max(plan->mid_max[cls], (int)r.mid);
            plan->vc_max[cls] = std::max(plan->vc_max[cls], (int)r.vc);
        };
        account(class_l, rl);
        account(class_c, rc2);
        // tile geometry resolved here once (the kernel only loads it); items are ordered by the first source row each
        // tile reads, so tiles of different rungs over the same band run together and share the source through L2
        bool bad_item = false;
        auto scale_items = [&](const PlaneTables &T, const PlaneTables *T_dev, const std::vector<int32_t> &hp, const std::vector<int32_t> &vp,
                               const std::vector<int32_t> &kb, uint8_t kind, int band_scale, int cls) {
            const bool luma = kind == ITEM_SCALE_LUMA, src_il = cfg.src_fmt == OTTGPU_FMT_P010;
            const bool dst_il = R.fmt == OTTGPU_FMT_NV12 || R.fmt == OTTGPU_FMT_P010;
            const int pxb = (!luma && src_il) ? 4 : bps;
            const int spv = 16 / pxb, slack = 4 / pxb > 0 ? 4 / pxb : 1;
            for (int ty = 0; ty < T.tiles_y; ++ty)
                for (int tx = 0; tx < T.tiles_x; ++tx) {
                    Item it;
                    std::memset(&it, 0, sizeof(it));
                    it.slot = (uint8_t)s;
                    it.kind = kind;
                    it.thumb = is_thumb;
                    const bool strip = OTTGPU_STRIP && !kb.empty() && bps == 1;
                    int x0 = tx * T.tw, y0 = ty * T.th;
                    int tw = std::min(T.tw, T.dstW - x0), th = std::min(T.th, T.dstH - y0);
                    if (strip) {                                 // the plane's column groups dealt out evenly over the tiles of a row
                        const int groups = (T.dstW + 7) / 8;
                        const int g0 = (int)((long long)tx * groups / T.tiles_x), g1 = (int)((long long)(tx + 1) * groups / T.tiles_x);
                        x0 = 8 * g0;
                        tw = std::min(8 * g1, T.dstW) - x0;
                    }
                    const int sy0 = vp[y0], sr = std::min(vp[y0 + th - 1] + T.vtaps, T.srcH) - sy0;
                    int sx0 = hp[x0] & ~(spv - 1), last = hp[x0 + tw - 1] + 4 * T.hq + slack;   // + one word of funnel-shift slack
                    if (!kb.empty()) {                           // MMA paths: the K windows (bytes) of the tile's column groups
                        int b0 = 1 << 30, b1 = 0;
                        for (int gi = x0 / 8; gi <= (x0 + tw - 1) / 8; ++gi) {
                            b0 = std::min(b0, kb[gi] & ~15);
                            b1 = std::max(b1, kb[gi] + 32 * T.hks);
                        }
                        sx0 = b0 / pxb;
                        last = (b1 + pxb - 1) / pxb;
                    }
                    it.T = T_dev;
                    it.x0 = (uint16_t)x0; it.y0 = (uint16_t)y0; it.tw = (uint16_t)tw; it.th = (uint16_t)th;
                    it.sy0 = (uint16_t)sy0; it.sr = (uint16_t)sr; it.sx0 = (uint16_t)sx0;
                    it.nvec = (uint16_t)((last - sx0 + spv - 1) / spv);
                    it.spw = (uint16_t)T.sw_words;
                    it.mid_pitch = (uint16_t)T.mid_pitch;
                    it.mid_rows = (uint16_t)((sr + 15) & ~15);
                    it.tma_x = (uint16_t)(sx0 * pxb / 4);
                    // one staged plane occupies sr_max rows of spw words, padded to the 128 bytes a TMA destination must be aligned to
                    it.plane_stride = (T.sr_max * T.sw_words * 4 + 127) & ~127;
                    it.box_bytes = T.sr_max * T.sw_words * 4;
                    it.bps = (uint8_t)bps;
                    it.nplanes = luma ? 1 : 2;
                    it.tw_log2 = T.tw == 64 ? 6 : 5;
                    it.approx_v = (uint8_t)R.approx_v;
                    it.src_interleaved = !luma && src_il;
                    it.dst_interleaved = !luma && dst_il;
                    it.staged_planes = (luma || src_il) ? 1 : 2;
                    it.px_bytes = (uint8_t)pxb;
                    it.src_shift = src_il ? 6 : 0;
                    it.dst_shift = R.fmt == OTTGPU_FMT_P010 ? 6 : 0;
                    it.tma_ok = (uint8_t)T.tma_ok;
                    // the tile's slices of the V tables (device pointers); fetched by two bulk copies when 16-byte aligned
                    it.vc_src = T.vc + (size_t)y0 * T.vstride;
                    it.vpos_src = T.vpos + y0;
                    it.vpos_off = (uint16_t)(T.th * T.vstride * 4);
                    it.vstride = (uint16_t)T.vstride;
                    it.strip = strip ? (T.cpt == 8 ? 1 : 2) : 0;
                    const size_t vcb = ((size_t)th * T.vstride * 4 + 15) & ~(size_t)15, vpb = ((size_t)th * 4 + 15) & ~(size_t)15;
                    const bool bulk = ((size_t)y0 * T.vstride * 4) % 16 == 0 && ((size_t)y0 * 4) % 16 == 0 && it.vpos_off % 16 == 0 &&
                                      vcb <= (size_t)it.vpos_off && vcb < 65536;
                    it.vc_bytes = bulk ? (uint16_t)vcb : 0;
                    it.vpos_bytes = bulk ? (uint16_t)vpb : 0;
                    it.vtaps = (uint16_t)T.vtaps; it.vbias = (uint16_t)T.vbias; it.vsum = (uint16_t)T.vsum; it.exact_from = (uint16_t)T.exact_from; it.dstW = (uint16_t)T.dstW;
                    it.hks = (uint8_t)std::min(T.hks, 255); it.hsplit = (uint8_t)T.hsplit; it.hq = (uint8_t)std::min(T.hq, 255);
                    if (!kb.empty()) {
                        const int frag_words = bps == 1 ? 128 : 256;                       // per column group and K step
                        it.hmma = T.hmma + (size_t)(x0 / 8) * T.hks * frag_words;
                        it.hmma2 = (!luma && src_il) ? T.hmma2 + (size_t)(x0 / 8) * T.hks * frag_words : it.hmma;
                        for (int gi = x0 / 8; gi <= (x0 + tw - 1) / 8; ++gi) {
                            const int kwv = (kb[gi] - sx0 * pxb) >> 2;
                            if (kwv > 65535 || kwv < 0) { err = "source window of one tile too wide"; bad_item = true; }
                            it.kw[gi - x0 / 8] = (uint16_t)kwv;
                        }
                    }
                    if (!kb.empty() && !strip) deal_hruns(it, nullptr);   // H-pass runs of the compute warps (MMA paths, round-1 layout)
                    if (cls == 2) plan->thumb_items.push_back(it);
                    else if (cls == 1) plan->big_items.push_back(it);
                    else ordered.push_back({it, band_scale * sy0});
                }
        };
        const PlaneTables *dev_luma = &plan->dev->rung[s].luma, *dev_chroma = &plan->dev->rung[s].chroma;   // addresses only
        scale_items(R.luma, dev_luma, lh, lv, lk, ITEM_SCALE_LUMA, 1, class_l);
        scale_items(R.chroma, dev_chroma, chh, cv, ck, ITEM_SCALE_CHROMA, 2, class_c);
        if (bad_item) return nullptr;
    }
    std::stable_sort(ordered.begin(), ordered.end(), [](const Sortable &a, const Sortable &b) { return a.band < b.band; });
    if (!OTTGPU_STRIP) {                                          // second pass: H shares against the plan-wide average V share
        static const int mode = [] { const char *e = std::getenv("OTT_HBAL_MODE"); return e ? std::atoi(e) : 0; }();   // measured: 0 wins
        constexpr int kWarps = kThreads / 32;
        double avg[kWarps] = {0};
        int n_tiles = 0;
        for (const Sortable &so : ordered)
            if ((so.it.kind == ITEM_SCALE_LUMA || so.it.kind == ITEM_SCALE_CHROMA) && so.it.hmma) {
                double l[kWarps];
                tile_vload(so.it, l);
                for (int w = 0; w < kWarps; ++w) avg[w] += l[w];
                ++n_tiles;
            }
        if (mode == 1 && n_tiles > 0) {
            for (int w = 0; w < kWarps; ++w) avg[w] /= n_tiles;
            for (Sortable &so : ordered)
                if ((so.it.kind == ITEM_SCALE_LUMA || so.it.kind == ITEM_SCALE_CHROMA) && so.it.hmma) deal_hruns(so.it, avg);
        }
    }
    for (const Sortable &so : ordered) plan->items.push_back(so.it);

    if (cudaMemcpy(plan->dev, &P, sizeof(PlanDev), cudaMemcpyHostToDevice) != cudaSuccess) { err = "cudaMemcpy(plan) failed"; return nullptr; }
    return plan;
}

}  // namespace

}  // namespace ottgpu
