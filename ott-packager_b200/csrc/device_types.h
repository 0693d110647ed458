// Device-visible descriptors shared by the host planner and the sm_100a kernels.
//
// Vocabulary: a *channel* is one transvideo pipeline (one fillet process in the reference); a *rung* is one ABR
// rendition (core->cd->transvideo_info[i]); a *plan* is everything sws_getContext() would have produced for all rungs
// of one channel configuration (polyphase tables + tiling); a *work item* is one output tile of one plane of one rung
// of one channel's current frame and is executed by one CTA.
#pragma once
#include <cstdint>

namespace ottgpu {

constexpr int kMaxRungs = 8;
constexpr int kMaxRungSlots = kMaxRungs + 1;   // slot 8 = thumbnail (transvideo.c:2155-2174)
// H-pass intermediates in shared memory: 32-bit (no extraction work in the V pass) or 16-bit (half the footprint)
#if !defined(OTTGPU_MID_BITS)
#define OTTGPU_MID_BITS 16      /* measured: 32-bit intermediates save V-pass instructions but halve the tile height: slower */
#endif
#if OTTGPU_MID_BITS == 32
typedef int32_t mid_t;
#else
typedef int16_t mid_t;
#endif
// tile descriptors (and V-table regions) in flight per CTA (powers of two).  16-bit ladders run long lanczos filters whose
// V tables are large (tile rows x taps x 4 bytes): two regions instead of four leave the shared memory to the tiles
// (measured on cfg3: 167 -> 146 us per frame); the 8-bit ladders are fastest with four (final build, cfg4, ms per 64 frames:
// two 0.446, four 0.430, eight 0.439; -DOTTGPU_SLOTS8=n).
#if !defined(OTTGPU_SLOTS8)
#define OTTGPU_SLOTS8 4
#endif
constexpr int kGeomSlots8 = OTTGPU_SLOTS8, kGeomSlots16 = 2;
constexpr int geom_slots(int bps) { return bps == 2 ? kGeomSlots16 : kGeomSlots8; }
constexpr int kTensorMapBytes = 128;           // sizeof(CUtensorMap)
#if !defined(OTTGPU_COMPUTE_WARPS)
#define OTTGPU_COMPUTE_WARPS 12                /* experiments: 15 warps at 64 registers (strip mode) */
#endif
constexpr int kThreads = 32 * OTTGPU_COMPUTE_WARPS;   // compute threads of a ladder CTA (12 warps); the kernel adds one producer warp
// 8-bit scale tiles in STRIP mode: a tile is up to 12 column groups of 8 output columns wide and compute warp w owns column
// group w for BOTH passes (H pass: its 8 columns of every staged row; V pass: lane = output row, 8 columns each), so a
// tile needs no CTA-wide barrier at all (a __syncwarp between the passes), one H-pass result region is enough (a warp only
// ever touches its own columns of it) and the warps of a CTA drift apart freely: while one is in its V pass (multiply /
// ALU pipes) another runs the next tile's H pass (shared loads / tensor pipe).  0 = the default layout (64-column tiles,
// work dealt out across the CTA, one named barrier between the passes).
// MEASURED (round 2, B200, profiles/README.md): bit-exact, and no faster -- cfg2 0.584 vs 0.583 ms per 64 frames, cfg1
// 0.342 vs 0.316, cfg4 0.464 vs 0.452.  The barrier stalls (18 % of the samples) go away, but the kernel issues the same
// useful instructions at the same rate (0.61 per scheduler cycle): the wider tiles stage wider source rows, so the low
// rungs get shorter tiles (more halo rows in the H pass), lane = row costs 2-way bank conflicts on 2:1 rungs, and warps
// without a column group in a tile idle.  Kept as a build option (scripts/build_variant.sh strip -DOTTGPU_STRIP=1).
#if !defined(OTTGPU_STRIP)
#define OTTGPU_STRIP 0
#endif
constexpr int kStripGroups = kThreads / 32;    // column groups per strip tile = compute warps
constexpr int kStripWidth = 8 * kStripGroups;  // 96 output columns

enum ItemKind : uint8_t {
    ITEM_SCALE_LUMA = 0,     // H+V polyphase of one luma tile
    ITEM_SCALE_CHROMA = 1,   // both chroma planes of one tile (so NV12/P010 writers can interleave)
    ITEM_COPY_LUMA = 2,      // same-size rung: libswscale's unscaled path = plane copy
    ITEM_COPY_CHROMA = 3,    // same-size rung chroma: copy or I420->NV12 interleave (transvideo.c:543-557)
};

struct PlaneTables {
    const int32_t *hpos;     // [dstW]   first source column of each output column
    const uint32_t *hc_hi;   // [dstW][hq] 8-bit path: per tap quad, high bytes  (coef >> 8, signed)   packed s8x4
    const uint32_t *hc_lo;   // [dstW][hq]                          low bytes   (coef & 255, unsigned) packed u8x4
    const int16_t *hc16;     // [dstW][4*hq] plain 14-bit coefficients (16-bit sample path)
    const uint32_t *hc_pair; // [dstW][hq+1][2] 8-bit path: s16x2 coefficient pairs laid out against the source WORDS the
                             //   window starts in (taps shifted by hpos&3 bytes, zero padded): {bytes 0,1}, {bytes 2,3}
    const uint32_t *hmma;    // 8-bit path: [ceil(dstW/8)][hks][32 lanes][4] IMMA.16832 B fragments of each group of 8 output
                             //   columns: {hi.b0, hi.b1, lo.b0, lo.b1}, coefficient = hi << hsplit | lo, laid out against
                             //   the source bytes hkbase[group] + 32*kstep + k
    const uint32_t *hmma2;   // 16-bit path, interleaved chroma: the V plane's fragments (hmma holds U's)
    const int32_t *hkbase;   // [ceil(dstW/8)] first source byte of the group's K window (multiple of 4)
    const int32_t *vpos;     // [dstH]   first source row of each output row
    const int32_t *vc;       // [dstH][vtaps] 12-bit coefficients, widened to 32 bits (the V pass multiplies them as they are)
    int srcW, srcH, dstW, dstH;
    int hq;                  // horizontal taps / 4 (tables zero-padded to a multiple of four taps)
    int hks;                 // 8-bit path: K steps (32 source bytes each) of the H-pass contraction
    int hsplit;              // 8-bit path: 7 (c = 128*hi + lo, lo < 256) or 8 (c = 256*hi + lo); 16: the 16-bit MMA path
    int mid_pitch;           // H-pass result row pitch in shared memory (mid_t elements): tw + 8, so that the fragment
                             //   stores of the H pass and the 16-byte loads of the V pass are both bank-conflict free
    int vtaps;
    int vstride;             // 32-bit entries per output row in vc (>= vtaps)
    int vbias;               // 8-bit target A fast path: per-product bias (max|c| >> 1) + 1 of the V pass, 0 = not usable
                             //   (odd tap count, or taps * (max|c| + 2) would overflow a 16-bit lane)
    int vsum;                // 8-bit target A FMA path (round 3): the common sum of every row's V coefficients (4096) when the
                             //   vertical pass may run as round-down FP32 FMAs (even taps, sum|c| small enough for the float
                             //   accumulator to stay inside one binade), 0 = not usable
    int exact_from;          // 8-bit target A: first output row computed with the exact vertical formula
    int tw, th;              // tile size in output samples
    int cpt;                 // V pass: adjacent columns per thread (8 | 4), chosen so the tile's work divides evenly
    int tiles_x, tiles_y;
    int sw_words;            // shared-memory source row pitch in 32-bit words (max over tiles, + funnel slack)
    int sr_max;              // max source rows any tile needs
    int tma_ok;              // tiles can be fetched with one 2-D TMA box (sw_words x sr_max u32 elements) per plane
};

struct RungDev {
    PlaneTables luma, chroma;
    int fmt;                 // OTTGPU_FMT_*
    int scaled;              // 0: same size as the source (copy items)
    int approx_v;            // 1: 8-bit target A vertical pass (pmulhw-style truncation) on rows < exact_from
};

struct PlanDev {
    int src_fmt;             // OTTGPU_FMT_*
    int bps;                 // bytes per sample (1 | 2)
    int n_slots;             // rungs (+1 if a thumbnail slot exists)
    int thumb_slot;          // -1: none
    RungDev rung[kMaxRungSlots];
};

// Per channel, per launch: where this step's frame lives.
struct Binding {
    const PlanDev *plan;
    int active;              // 0: channel produces nothing this step (yadif priming) -> its items exit at once
    int thumb_active;        // thumbnail slot wanted this step
    uint32_t skip_mask;      // rung slots already produced elsewhere this step (yadif wrote the full-size rung)
    const unsigned char *tmaps;  // CUtensorMap[kMaxRungSlots][3] (luma, U, V boxes) of the ARENA this step's source frame lives in, or null
    int src_z;               // the frame's buffer index inside that arena (third tensor-map coordinate)
    const uint8_t *src[3];   // Y,U,V  or  Y,UV
    int src_pitch[3];        // bytes
    uint8_t *dst[kMaxRungSlots][3];
    int dst_pitch[kMaxRungSlots][3];
};

// One work item = one CTA-sized piece of work.  Everything that does not depend on where the channel's frame lives is
// resolved on the host when the plan is built, so that the kernel's producer warp only copies this block into shared
// memory and patches the frame / surface pointers of the step in: the producer's per-tile latency bounds the tile rate
// of a CTA (measured: with the geometry derived on the device a tile cost ~3 us whatever its size).
struct Item {
    const PlaneTables *T;    // the tables of this plane kind of this rung (device pointer into the plan)
    const int32_t *vc_src;   // this tile's slice of the V coefficient table ...
    const int32_t *vpos_src; // ... and of the row positions
    uint16_t chan;           // index into the launch's Binding array
    uint8_t slot;            // rung slot
    uint8_t kind;            // ItemKind
    uint16_t x0, y0;         // output tile origin (copy items: y0 = first row)
    uint16_t tw, th;         // output tile size (copy items: th = rows)
    uint16_t sy0, sr;        // source rows [sy0, sy0+sr) the tile reads
    uint16_t sx0, nvec;      // first staged source sample (16-byte aligned) and 16-byte vectors per staged row
    uint16_t spw;            // smem pitch of the staged source, 32-bit words
    uint16_t mid_pitch;      // smem pitch (mid_t elements) of the H-pass result
    uint16_t mid_rows;       // rows per plane in the H-pass result (sr rounded up to the H pass's 16-row groups)
    uint16_t tma_x;          // TMA box x coordinate in 32-bit elements
    int32_t plane_stride;    // bytes between the two staged chroma planes inside a source buffer (128-byte multiple)
    int32_t box_bytes;       // bytes one TMA box delivers
    uint16_t vc_bytes;       // V coefficient slice as a bulk copy (16-byte multiple), 0: copied by the producer's lanes
    uint16_t vpos_bytes;     // row position slice as a bulk copy
    uint16_t vpos_off;       // offset of the row positions inside the table region (= T->th * vtaps * 4)
    uint8_t bps;             // source/destination bytes per sample
    uint8_t nplanes;         // 1 luma, 2 chroma pair
    uint8_t tw_log2;         // T->tw is 32 or 64
    uint8_t approx_v;
    uint8_t src_interleaved; // chroma samples interleaved in the source (P010): staged as ONE plane of (U,V) words
    uint8_t dst_interleaved; // NV12 / P010 chroma writer
    uint8_t staged_planes;   // planes in the staged window (1 for luma and for interleaved chroma, else 2)
    uint8_t px_bytes;        // bytes per staged sample position: 1 (8-bit), 2 (16-bit), 4 (interleaved 16-bit pair)
    uint8_t src_shift;       // 6 for P010 (MSB aligned), else 0
    uint8_t dst_shift;       // 6 for P010
    uint8_t tma_ok;          // the source window fits one 2-D TMA box per plane
    uint8_t thumb;           // item of the thumbnail slot: runs only on the steps the thumbnail is due
    uint16_t vstride;        // V coefficient table: 32-bit entries per output row (>= vtaps; padded so that 32 lanes reading
                             //   32 different rows hit different banks)
    // 8-bit H pass: the (column group, 16-row group) pairs dealt to each compute warp as one contiguous run:
    // bits 0-7 first column group, 8-19 first row group, 20-31 number of pairs
    uint32_t hrun[kThreads / 32 <= 12 ? 12 : 16];
    // copies of the table fields the hot loops need, so that no thread waits on a chain of global loads through T
    // at the start of a tile (measured: ~1.1 us of fixed cost per tile per CTA went there)
    const uint32_t *hmma;    // 8-bit: B fragments of the tile's first column group
    uint16_t kw[16];         // 8-bit: K-window base of each of the tile's column groups, in 32-bit words from sx0
    uint16_t vtaps, vbias, exact_from, dstW;
    uint8_t hks, hsplit, hq;
    uint8_t strip;           // 8-bit strip mode: 1 = lane owns a row x 8 columns in the V pass, 2 = x 4 columns (short tiles)
    uint16_t vsum;           // copy of T->vsum (0: the V pass may not use the FMA path)
    uint16_t pad3;
    const uint32_t *hmma2;   // 16-bit interleaved chroma: fragments of the second (V) plane, else = hmma
};
static_assert(sizeof(Item) % 16 == 0 && (sizeof(Item) == 192 || OTTGPU_COMPUTE_WARPS > 12), "Item is copied into shared memory as twelve 16-byte vectors");

// Launch-wide constants of the persistent ladder kernel.
struct LadderLaunch {
    const Item *items;
    const Binding *binds;
    int n_items;
    int bps;                 // bytes per sample of every item's channel (1 | 2): selects the kernel instance
    int *counter;            // dynamic tile scheduler (zeroed before every launch)
    int claim;               // work items a CTA claims per atomic on the counter (1, 2 or 4; set by the launcher)
    int src_max;             // bytes of ONE staged-source buffer (max over the launch's tiles, multiple of 16)
    int mid_max;             // bytes of the H-pass result region (multiple of 16)
    int vc_max;              // bytes of the V coefficient + vpos region
    int double_buffer;       // 1: two source buffers (kept for experiments; the launcher always uses one: the V pass hides the copy)
    int mid_buffers;         // 2: H-pass results alternate between two regions, so a warp that finishes its share of
                             //    the V pass starts the next tile's H pass without waiting for the others
};

// yadif work description, one per channel per launch
struct YadifJob {
    const uint8_t *prev, *cur, *next;   // packed planar frames (I420 / I420P10)
    uint8_t *dst;
    int width, height, bps;
    int parity, tff;                    // as resolved by yadif: parity = tff ^ 1 for mode 0
    int active;
    int fast;                           // 8-bit, width % 16 == 0, 8-byte aligned frames: eight samples per thread
};

// decode-side converter (transvideo.c:2797-2840): one plane of one picture
enum ConvertMode : int {
    CONV_COPY8 = 0,      // 8-bit plane, rows re-packed
    CONV_REDUCE10 = 1,   // 10-bit plane -> 8 bit, 2x2 ordered dither (== the Bayer row >> 5 of the generic 1-tap path)
    CONV_VFILT8 = 2,     // 4:2:2 chroma, 8-bit: vertical 2:1 polyphase
    CONV_VFILT10 = 3,    // 4:2:2 chroma, 10-bit: vertical 2:1 polyphase with the 8x8 Bayer dither
};

struct ConvertPlane {
    const uint8_t *src;
    int src_pitch;           // bytes
    uint8_t *dst;
    int dst_pitch;           // bytes (packed I420: the plane width)
    int w, h;                // destination plane size in samples
    int mode;
    int dither_offset;       // column offset into the Bayer row (0: Y/U, 3: V)
    int fast;                // rows can be moved as aligned 16-sample vectors
};

struct ConvertJob {
    ConvertPlane plane[3];
    const int32_t *vpos;     // [chroma rows] first source row of each output row
    const int16_t *vc;       // [chroma rows * vtaps] 12-bit coefficients
    int vtaps;
    int vbias;               // 8-bit target A fast path: per-product bias (max|c| >> 1) + 1 of the V pass, 0 = not usable
                             //   (odd tap count, or taps * (max|c| + 2) would overflow a 16-bit lane)
    int exact_from;          // first chroma row that takes the exact vertical path under target A
    int approx;              // 1: target A (x86 default flags), 0: target B
    uint8_t bayer[64];       // ff_dither_8x8_128
};

// thumbnail JPEG (jpeg_core.cuh)
struct JpegHuff {
    uint16_t dc_code[2][12];
    uint8_t dc_len[2][12];
    uint16_t ac_code[2][256];
    uint8_t ac_len[2][256];
};

struct JpegJob {
    const uint8_t *plane[3];
    int pitch[3];
    int width, height;           // luma size; chroma planes are ceil(w/2) x ceil(h/2)
    int mcus_x, mcus_y, n_blocks;
    int16_t *coef;               // [n_blocks][64], zigzag order, scan order of blocks (MCU-major: Y00 Y01 Y10 Y11 Cb Cr)
    uint32_t *bits;              // [n_blocks + 1] code length of each block, then (after the scan) its bit position; [n] = total
    uint32_t *stream;            // entropy-coded bits, big-endian inside each 32-bit word, zeroed before the launch
    uint32_t stream_words;
    uint8_t *out;                // header (header_bytes, written by the host) + stuffed entropy data + EOI
    uint32_t out_capacity;
    int header_bytes;
    uint32_t *result;            // [0] total bytes of the file, [1] 1 if something did not fit
    uint32_t *ff_count;          // [chunks + 1] scratch of the stuffing pass
    uint32_t *hist;              // [4][256] symbol counts of the picture: DC luma, DC chroma, AC luma, AC chroma
    const JpegHuff *huff;        // the picture's tables (device memory, written between the two launches)
    uint16_t q[64];              // quantisation table, natural order (the DQT segment)
    uint16_t qmat16[64], bias16[64];   // libavcodec's 16-bit quantiser constants for it: level = ((|x| + bias16) * qmat16) >> 16
};

}  // namespace ottgpu
