// libottgpu runtime: channels (yadif window + per-rung surfaces), batches (one fused launch per stage over many
// channels), pools, and the C ABI of include/ottgpu.h.
//
// What it stands in for in the reference: the avfilter graph + SwsContexts + packed-buffer copies that
// video_prepare_thread / video_scale_thread own (transvideo.c:1504-1683, 1896-2429).  Frame scheduling follows
// yadif_common.c ff_yadif_filter_frame as summarised in SURVEY.md Appendix C.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#if !defined(OTTGPU_EMULATE)
#include <cuda.h>            // CUtensorMap + cuTensorMapEncodeTiled prototype (resolved through the runtime, no -lcuda)
#endif
#include "../../include/ottgpu.h"
#include "kernels.h"
#include "plan.h"

using namespace ottgpu;

namespace {

thread_local std::string t_err;

int fail(int code, const std::string &what)
{
    t_err = what;
    return code;
}
int fail_cuda(cudaError_t e, const char *where)
{
    t_err = std::string(where) + ": " + cudaGetErrorString(e);
    cudaGetLastError();
    return OTTGPU_ERR_CUDA;
}
#define CK(call)                                                  \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return fail_cuda(e_, #call);       \
    } while (0)

std::mutex g_dev_mu;
bool g_dev_ready[64] = {false};

#if !defined(OTTGPU_EMULATE)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode_tiled = nullptr;
#endif

int prepare_device(int gpu)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
    if (gpu < 0 || gpu >= n || gpu >= 64) return fail(OTTGPU_ERR_ARG, "gpu ordinal out of range");
    CK(cudaSetDevice(gpu));
    std::lock_guard<std::mutex> lock(g_dev_mu);
    if (!g_dev_ready[gpu]) {
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, gpu));
        if (prop.major != 10) return fail(OTTGPU_ERR_CUDA, "libottgpu carries sm_100a code only; device is not compute capability 10.x");
        CK(configure_kernels());
#if !defined(OTTGPU_EMULATE)
        if (!g_encode_tiled) {
            void *fn = nullptr;
            cudaDriverEntryPointQueryResult st;
            CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &st));
            if (st != cudaDriverEntryPointSuccess || !fn) return fail(OTTGPU_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
            g_encode_tiled = (EncodeTiledFn)fn;
        }
#endif
        g_dev_ready[gpu] = true;
    }
    return OTTGPU_OK;
}

struct FrameRef {
    const uint8_t *dev = nullptr;
    int slot = -1;               // >= 0: channel-owned upload slot; -1: caller's device memory
    int interlaced = 0, tff = 1;
    int64_t pts = 0, dts = 0;
    void *opaque = nullptr;
    bool valid = false;
};

void plane_ptrs(int fmt, const uint8_t *base, int w, int h, const uint8_t *p[3], int pitch[3])
{
    const int bps = fmt_bps(fmt), cw = (w + 1) / 2, ch = (h + 1) / 2;
    p[0] = base;
    pitch[0] = w * bps;
    p[1] = base + (size_t)w * h * bps;
    if (fmt == OTTGPU_FMT_P010 || fmt == OTTGPU_FMT_NV12) {
        pitch[1] = 2 * cw * bps;
        p[2] = nullptr;
        pitch[2] = 0;
    } else {
        pitch[1] = pitch[2] = cw * bps;
        p[2] = p[1] + (size_t)cw * ch * bps;
    }
}

}  // namespace

struct ottgpu_channel {
    ottgpu_cfg cfg{};
    std::shared_ptr<Plan> plan;
    int gpu = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    size_t src_bytes = 0;
    // upload slots: the yadif window (3) + the frame being uploaded + one step of release delay, so that the upload
    // of step t+1 (copy stream) never targets a frame the kernels of step t may still read
    static constexpr int kSlots = 6;
    // The slots of all channels of a batch live in ONE allocation: slot s of channel i at arena + (s * arena_n + i) * stride
    // - when every channel of a step takes the same slot (the steady state) and the host frames are consecutive buffers of
    // one pinned pool, the step's uploads are a single cudaMemcpyAsync.  A channel on its own owns an arena with arena_n = 1.
    uint8_t *arena = nullptr;
    size_t arena_stride = 0;
    int arena_n = 1, arena_idx = 0;
    bool own_arena = false;
    bool slot_used[kSlots] = {false};
    uint8_t *slot_ptr(int s) const { return arena + ((size_t)s * arena_n + arena_idx) * arena_stride; }
    int slot_cooling = -1;                             // released last step, reusable from the next one
    FrameRef prev, cur, next;
    uint8_t *deint = nullptr;                          // yadif output when no rung surface can take it directly
    uint8_t *stage[2][kMaxRungSlots] = {{nullptr}, {nullptr}};   // device surfaces for host-destined rungs / thumbnail,
                                                                 // one set per descriptor set (read-back of step t overlaps step t+1)
    int thumb_count = 0;
    int direct_rung = -1;                              // rung whose surface is byte-identical to a deinterlaced frame
    ottgpu_batch *solo = nullptr;
    ottgpu_batch *member_of = nullptr;                 // the multi-channel batch whose upload arena this channel uses
    // Tensor maps: one set per ARENA the channel's source frames come from (its upload arena, a caller's ottgpu_pool, the
    // deinterlacer scratch): the maps are three-dimensional, {row words, rows, buffer index}, so any buffer of the arena
    // is addressed by the z coordinate and no map is ever encoded, uploaded or evicted on the frame path.
    struct MapSet { const uint8_t *base = nullptr; size_t stride = 0; unsigned char *dev = nullptr; };
    std::vector<MapSet> maps;
    unsigned char *maps_pool = nullptr;                // all sets in one allocation made at creation (no cudaMalloc mid-stream)
    bool wants_maps = false;
};

struct ottgpu_batch {
    std::vector<ottgpu_channel *> ch;
    int gpu = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;                // host frames are uploaded here while the previous step computes
    cudaStream_t readback_stream = nullptr;            // results for host destinations are copied back here
    Item *d_items = nullptr, *d_big_items = nullptr;   // static per batch: every channel's plan items, chan filled in;
                                                       // 8-bit channels' items first, then the 16-bit ones (one kernel
                                                       // instance per sample size)
    int n_items = 0, n_big_items = 0, max_thumb_items = 0;
    int n_items_bps[2] = {0, 0}, n_big_items_bps[2] = {0, 0};
    int src_max[3] = {0, 0, 0}, mid_max[3] = {0, 0, 0}, vc_max[3] = {0, 0, 0};   // launch classes: rung, big-tile rung, thumbnail
    // Descriptors are double-buffered so that an ASYNC submit can be prepared while the previous step still runs.
    struct Set {
        Binding *h_binds = nullptr, *d_binds = nullptr;    // pinned / device
        YadifJob *h_jobs = nullptr, *d_jobs = nullptr;
        Item *h_thumb_items = nullptr, *d_thumb_items = nullptr;
        int *d_counters = nullptr;                          // tile schedulers: three launch classes x two sample sizes
        uint8_t *h_thumbs = nullptr;                        // pinned landing area for this step's thumbnails (all channels)
        struct PendingThumb { void *dst; const uint8_t *src; size_t bytes; };
        std::vector<PendingThumb> pending_thumbs;           // pinned -> the caller's malloc'ed buffer, done when the step retires
        cudaEvent_t done = nullptr, uploaded = nullptr, computed = nullptr, read_back = nullptr, k0 = nullptr, k1 = nullptr, y0 = nullptr, y1 = nullptr;
        bool in_flight = false, timed = false, timed_yadif = false, has_readback = false;
    } set[2];
    std::vector<size_t> thumb_offset;                      // per channel, into Set::h_thumbs
    size_t thumb_stage_bytes = 0;
    int turn = 0;
    int last_launches = 0, last_items = 0, last_uploads = 0;
    uint8_t *arena = nullptr;                              // upload slots of all channels (when they share a frame size)
    size_t arena_stride = 0;
    struct Upload { uint8_t *dst; const void *src; size_t bytes; };
    std::vector<Upload> uploads;
    bool timing = false;
    double ms_sum = 0.0, yadif_ms_sum = 0.0;
    int ms_count = 0, yadif_ms_count = 0;
    struct Readback { void *host; const uint8_t *dev; size_t bytes; int rung; };
    std::vector<Readback> readbacks;
    // Device surfaces of host-destined rungs, one allocation per rung and descriptor set: channel i's surface of rung r at
    // stage_arena[set][r] + i * stage_stride[r] (stride = the largest channel's surface, 256-byte aligned = ottgpu_pool's
    // buffer stride).  Consecutive channels landing in consecutive buffers of one pinned pool are read back with ONE copy per
    // rung instead of one per channel per rung.
    uint8_t *stage_arena[2][kMaxRungs] = {{nullptr}, {nullptr}};
    size_t stage_stride[kMaxRungs] = {0};
    int last_readbacks = 0;
};

namespace {

bool referenced(const ottgpu_channel *c, const uint8_t *dev)
{
    return (c->prev.valid && c->prev.dev == dev) || (c->cur.valid && c->cur.dev == dev) || (c->next.valid && c->next.dev == dev);
}

void release_frame(ottgpu_channel *c, const FrameRef &f, ottgpu_frame_out *out)
{
    if (!f.valid || referenced(c, f.dev)) return;
    if (f.slot >= 0) {
        if (c->slot_cooling >= 0) c->slot_used[c->slot_cooling] = false;
        c->slot_cooling = f.slot;
    }
    else if (out && out->n_released < 4) out->released[out->n_released++] = f.dev;
}

size_t arena_stride_for(size_t bytes) { return (bytes + 255) & ~(size_t)255; }      // = ottgpu_pool's buffer stride

int take_slot(ottgpu_channel *c)
{
    if (!c->arena) {                                   // a channel that uploads for the first time outside a shared arena
        c->arena_stride = arena_stride_for(c->src_bytes);
        c->arena_n = 1;
        c->arena_idx = 0;
        if (cudaMalloc((void **)&c->arena, ottgpu_channel::kSlots * c->arena_stride + 64) != cudaSuccess) {
            cudaGetLastError();
            c->arena = nullptr;
            return -1;
        }
        c->own_arena = true;
    }
    for (int i = 0; i < ottgpu_channel::kSlots; ++i)
        if (!c->slot_used[i]) {
            c->slot_used[i] = true;
            return i;
        }
    return -1;
}

// device pools created through ottgpu_pool_create: a frame pointer inside one of them is (pool base, stride, index)
struct PoolSpan { const uint8_t *base; size_t stride; int count; int gpu; };
std::mutex g_pools_mu;
std::vector<PoolSpan> g_pools, g_host_pools;

// both addresses inside one pinned host pool created through ottgpu_pool_create
bool same_host_pool(const void *a, const void *b)
{
    std::lock_guard<std::mutex> lock(g_pools_mu);
    for (const PoolSpan &s : g_host_pools) {
        const uint8_t *lo = s.base, *hi = s.base + s.stride * s.count;
        if ((const uint8_t *)a >= lo && (const uint8_t *)a < hi) return (const uint8_t *)b >= lo && (const uint8_t *)b < hi;
    }
    return false;
}

#if !defined(OTTGPU_EMULATE)
bool find_span(const ottgpu_channel *c, const uint8_t *p, PoolSpan &out)
{
    if (c->arena && p >= c->arena && p < c->arena + (size_t)ottgpu_channel::kSlots * c->arena_n * c->arena_stride) {
        out = {c->arena, c->arena_stride, ottgpu_channel::kSlots * c->arena_n, c->gpu};
        return true;
    }
    std::lock_guard<std::mutex> lock(g_pools_mu);
    for (const PoolSpan &s : g_pools)
        if (s.gpu == c->gpu && p >= s.base && p < s.base + s.stride * s.count) {
            out = s;
            return true;
        }
    return false;
}
#endif

int ensure_dev(uint8_t **p, size_t bytes)
{
    if (*p) return OTTGPU_OK;
    if (cudaMalloc((void **)p, bytes + 64) != cudaSuccess) {
        cudaGetLastError();
        return fail(OTTGPU_ERR_NOMEM, "cudaMalloc(surface) failed");
    }
    return OTTGPU_OK;
}

// Tensor maps for the frame at `frame` (device memory, packed planar): one 3-D box shape per scaled rung plane, shared by
// every buffer of the arena the frame lives in; *z is the frame's index inside that arena.
// Returns null maps when TMA cannot be used for this frame (alignment, or more distinct arenas than the channel keeps
// sets for) -- the kernel then stages through registers; nothing on this path synchronises the device.
int frame_tensor_maps(ottgpu_channel *c, const uint8_t *frame, cudaStream_t stream, const unsigned char **out, int *z)
{
    *out = nullptr;
    *z = 0;
#if !defined(OTTGPU_EMULATE)
    if (!c->wants_maps) return OTTGPU_OK;
    const ottgpu_cfg &cfg = c->cfg;
    PoolSpan span;
    if (!find_span(c, frame, span)) span = {frame, arena_stride_for(c->src_bytes), 1, c->gpu};   // a lone buffer: arena of one
    const size_t idx = (size_t)(frame - span.base) / span.stride;
    if (span.base + idx * span.stride != frame || (span.stride & 15)) return OTTGPU_OK;        // not at a buffer start
    const uint8_t *pl[3];
    int pitch[3];
    plane_ptrs(cfg.src_fmt, span.base, cfg.src_width, cfg.src_height, pl, pitch);
    for (int k = 0; k < (cfg.src_fmt == OTTGPU_FMT_P010 ? 2 : 3); ++k)
        if (((uintptr_t)pl[k] | (uintptr_t)pitch[k]) & 15) return OTTGPU_OK;
    *z = (int)idx;
    for (auto &m : c->maps)
        if (m.base == span.base && m.stride == span.stride) {
            *out = m.dev;
            return OTTGPU_OK;
        }
    constexpr size_t kSetBytes = (size_t)kMaxRungSlots * 3 * kTensorMapBytes;
    constexpr size_t kMaxSets = 8;
    if (c->maps.size() >= kMaxSets) return OTTGPU_OK;   // more source arenas than sets: this frame is staged through registers
    if (!c->maps_pool && cudaMalloc((void **)&c->maps_pool, kMaxSets * kSetBytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(OTTGPU_ERR_NOMEM, "cudaMalloc(tensor maps) failed");
    }
    c->maps.emplace_back();
    ottgpu_channel::MapSet *slot = &c->maps.back();
    slot->dev = c->maps_pool + (c->maps.size() - 1) * kSetBytes;
    alignas(64) static thread_local CUtensorMap host[kMaxRungSlots * 3];
    std::memset(host, 0, sizeof(host));
    const Plan &plan = *c->plan;
    // planes as 3-D arrays of 32-bit elements: {elements per row, rows, buffers}; P010 has Y and the interleaved UV plane
    const int cw = (cfg.src_width + 1) / 2, chh = (cfg.src_height + 1) / 2, bps = fmt_bps(cfg.src_fmt);
    const bool semi = cfg.src_fmt == OTTGPU_FMT_P010;
    const int nmaps = semi ? 2 : 3;
    const int row_bytes[3] = {cfg.src_width * bps, semi ? 2 * cw * bps : cw * bps, cw * bps};
    const int rows[3] = {cfg.src_height, chh, chh};
    for (int s = 0; s < plan.n_slots; ++s) {
        const RungDev &R = plan.host_copy.rung[s];
        if (!R.scaled) continue;
        for (int k = 0; k < nmaps; ++k) {
            const PlaneTables &T = k ? R.chroma : R.luma;
            if (!T.tma_ok) continue;
            const cuuint64_t gdim[3] = {(cuuint64_t)(row_bytes[k] / 4), (cuuint64_t)rows[k], (cuuint64_t)span.count};
            const cuuint64_t gstride[2] = {(cuuint64_t)pitch[k], (cuuint64_t)span.stride};
            const cuuint32_t box[3] = {(cuuint32_t)T.sw_words, (cuuint32_t)T.sr_max, 1};
            const cuuint32_t estride[3] = {1, 1, 1};
            const CUresult r = g_encode_tiled(&host[s * 3 + k], CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, (void *)pl[k], gdim, gstride, box,
                                              estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                c->maps.pop_back();
                return fail(OTTGPU_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
            }
        }
    }
    CK(cudaMemcpyAsync(slot->dev, host, kSetBytes, cudaMemcpyHostToDevice, stream));
    slot->base = span.base;
    slot->stride = span.stride;
    *out = slot->dev;
#else
    (void)c;
    (void)frame;
    (void)stream;
#endif
    return OTTGPU_OK;
}

void drop_window(ottgpu_channel *c)
{
    for (int i = 0; i < ottgpu_channel::kSlots; ++i) c->slot_used[i] = false;
    c->slot_cooling = -1;
    c->prev = c->cur = c->next = FrameRef();
}

}  // namespace

extern "C" {

int ottgpu_abi_version(void) { return OTTGPU_ABI_VERSION; }

const char *ottgpu_build_info(void)
{
#if defined(OTTGPU_EMULATE)
    return "libottgpu EMULATION (tests/emu: kernel source compiled for the CPU; test infrastructure, not a product build)";
#else
    return "libottgpu sm_100a (hand-written CUDA kernels: ladder_kernel<1|2>, yadif_kernel, convert_kernel, jpeg_*_kernel; no CPU fallback)";
#endif
}

int ottgpu_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail_cuda(e, "cudaGetDeviceCount");
    return n;
}

const char *ottgpu_strerror(int code)
{
    switch (code) {
    case OTTGPU_OK: return "ok";
    case OTTGPU_ERR_ARG: return "bad argument";
    case OTTGPU_ERR_CUDA: return "CUDA error";
    case OTTGPU_ERR_NOMEM: return "out of memory / pool exhausted";
    case OTTGPU_ERR_STATE: return "bad call sequence";
    default: return "unknown error";
    }
}

const char *ottgpu_last_error(void) { return t_err.c_str(); }

size_t ottgpu_frame_bytes(int fmt, int width, int height) { return packed_frame_bytes(fmt, width, height); }

size_t ottgpu_rung_bytes(const ottgpu_rung_cfg *r)
{
    RungLayout l;
    return r && rung_layout(*r, l) ? l.bytes : 0;
}

int ottgpu_rung_layout(const ottgpu_rung_cfg *r, size_t offset[3], int pitch[3])
{
    RungLayout l;
    if (!r || !rung_layout(*r, l)) return fail(OTTGPU_ERR_ARG, "bad rung");
    for (int i = 0; i < 3; ++i) { offset[i] = l.offset[i]; pitch[i] = l.pitch[i]; }
    return l.planes;
}

int ottgpu_filter_table(int src_n, int dst_n, int filter, int target, int vertical, int32_t *pos, int16_t *coef)
{
    if (src_n < 4 || dst_n < 1 || !pos || !coef) return fail(OTTGPU_ERR_ARG, "bad table request");
    const PolyphaseTable t = plan_table(src_n, dst_n, filter, target, vertical != 0);
    if (t.taps > 256) return fail(OTTGPU_ERR_ARG, "filter longer than 256 taps: the caller's coef buffer holds dst_n*256 entries");
    std::memcpy(pos, t.pos.data(), sizeof(int32_t) * t.pos.size());
    std::memcpy(coef, t.coef.data(), sizeof(int16_t) * t.coef.size());
    return t.taps;
}

/* -------------------------------------------------------------------------------------------------- channel */
int ottgpu_channel_create(const ottgpu_cfg *cfg, ottgpu_channel **out)
{
    if (!cfg || !out) return fail(OTTGPU_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->deint < OTTGPU_DEINT_ALL || cfg->deint > OTTGPU_DEINT_BYPASS) return fail(OTTGPU_ERR_ARG, "bad deint mode");
    if (cfg->deint == OTTGPU_DEINT_ALL && cfg->src_fmt == OTTGPU_FMT_P010)
        return fail(OTTGPU_ERR_ARG, "yadif needs planar samples (vf_yadif has no P010); use I420P10 or deint=FLAGGED/BYPASS");
    if ((cfg->thumb_width > 0) != (cfg->thumb_height > 0) || (cfg->thumb_width > 0 && cfg->thumb_period < 1))
        return fail(OTTGPU_ERR_ARG, "bad thumbnail settings");
    int rc = prepare_device(cfg->gpu);
    if (rc) return rc;
    std::string err;
    std::shared_ptr<Plan> plan = get_plan(*cfg, err);
    if (!plan)
        return fail(err.find("cudaMalloc") != std::string::npos ? OTTGPU_ERR_NOMEM
                                                                : (err.find("cuda") != std::string::npos ? OTTGPU_ERR_CUDA : OTTGPU_ERR_ARG), err);
    std::unique_ptr<ottgpu_channel> c(new ottgpu_channel);
    c->cfg = *cfg;
    c->plan = plan;
    c->gpu = cfg->gpu;
    c->src_bytes = packed_frame_bytes(cfg->src_fmt, cfg->src_width, cfg->src_height);
    c->thumb_count = cfg->thumb_period > 0 ? ((cfg->thumb_phase % cfg->thumb_period) + cfg->thumb_period) % cfg->thumb_period : 0;
    if (cfg->stream) {
        c->stream = (cudaStream_t)cfg->stream;
    } else {
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    }
    for (int r = 0; r < cfg->n_rungs; ++r) {
        const ottgpu_rung_cfg &g = cfg->rung[r];
        if (g.width == cfg->src_width && g.height == cfg->src_height && g.fmt == cfg->src_fmt && g.pitch_align <= 1) {
            c->direct_rung = r;
            break;
        }
    }
    for (int sidx = 0; sidx < plan->n_slots; ++sidx) {
        const RungDev &R = plan->host_copy.rung[sidx];
        if (R.scaled && (R.luma.tma_ok || R.chroma.tma_ok)) c->wants_maps = true;
    }
    // the thumbnail's device staging (one per descriptor set) is allocated now: a cudaMalloc in the middle of a running
    // stream can stall it for milliseconds (measured: a one-off 3-136 ms gap at the first thumbnail of one channel)
    if (plan->thumb_slot >= 0)
        for (auto &set : c->stage) {
            rc = ensure_dev(&set[plan->thumb_slot], plan->layout[plan->thumb_slot].bytes);
            if (rc) {
                for (auto &other : c->stage) cudaFree(other[plan->thumb_slot]);
                if (c->own_stream) cudaStreamDestroy(c->stream);
                return rc;
            }
        }
    ottgpu_channel *raw = c.release();
    rc = ottgpu_batch_create(&raw, 1, &raw->solo);
    if (rc) {
        ottgpu_channel_destroy(raw);
        return rc;
    }
    *out = raw;
    return OTTGPU_OK;
}

int ottgpu_channel_gpu(const ottgpu_channel *ch) { return ch ? ch->gpu : OTTGPU_ERR_ARG; }

int ottgpu_channel_describe(const ottgpu_channel *ch, char *buf, size_t size)
{
    if (!ch) return fail(OTTGPU_ERR_ARG, "null channel");
    const Plan &p = *ch->plan;
    std::string out;
    char line[512];
    int db = 0, mb = 1;
    const size_t smem = ladder_smem_bytes(p.host_copy.bps, p.src_max[0], p.mid_max[0], p.vc_max[0], &db, &mb);
    snprintf(line, sizeof line, "source %dx%d fmt %d, %d rung slot(s), %zu work items/frame (+%zu big-tile, +%zu thumbnail), shared memory %zu B (%s), regions src %d mid %d vc %d\n",
             p.cfg.src_width, p.cfg.src_height, p.cfg.src_fmt, p.n_slots, p.items.size(), p.big_items.size(), p.thumb_items.size(), smem,
             mb == 2 ? "1 source buffer + 2 intermediate regions" : "1 source buffer + 1 intermediate region", p.src_max[0], p.mid_max[0], p.vc_max[0]);
    out += line;
    for (int s = 0; s < p.n_slots; ++s) {
        const RungDev &R = p.host_copy.rung[s];
        for (int k = 0; k < 2; ++k) {
            const PlaneTables &T = k ? R.chroma : R.luma;
            const char *hpath = p.host_copy.bps == 1 ? (T.hsplit == 7 ? "IMMA u8 (7-bit split)" : "IMMA u8 (8-bit split)")
                                                     : (T.hsplit == 16 ? "IMMA 16-bit raw bytes (4 products)" : "dp2a");
            snprintf(line, sizeof line, "  slot %d %s %s %dx%d -> %dx%d  taps H %d V %d  H pass %s, %d K step(s)  tile %dx%d  tiles %dx%d  src rows/tile %d  staged row %d B  exact_from %d  V bias %d\n",
                     s, s == p.thumb_slot ? "thumb" : "rung", k ? "chroma" : "luma", T.srcW, T.srcH, T.dstW, T.dstH, 4 * T.hq,
                     T.vtaps, hpath, T.hks, T.tw, T.th, T.tiles_x, T.tiles_y, T.sr_max, T.sw_words * 4, T.exact_from, T.vbias);
            out += R.scaled ? line : std::string("  slot ") + std::to_string(s) + (k ? " chroma" : " luma") + " same size: copy items\n";
        }
    }
    if (buf && size) {
        const size_t n = std::min(size - 1, out.size());
        std::memcpy(buf, out.data(), n);
        buf[n] = 0;
    }
    return (int)out.size();
}

int ottgpu_channel_reset(ottgpu_channel *ch)
{
    if (!ch) return fail(OTTGPU_ERR_ARG, "null channel");
    cudaSetDevice(ch->gpu);
    cudaStreamSynchronize(ch->stream);
    drop_window(ch);
    // thumb_count is NOT touched: the reference's thumbnail_count is a local of the thread loop that the discontinuity
    // block (transvideo.c:1959-1970) leaves alone, so the 150-frame cadence runs straight through a signal loss
    return OTTGPU_OK;
}

void *ottgpu_channel_pending_opaque(const ottgpu_channel *ch)
{
    // the frame submitted last has not been produced yet (yadif's one-frame latency); BYPASS keeps nothing
    if (!ch || ch->cfg.deint == OTTGPU_DEINT_BYPASS || !ch->next.valid) return nullptr;
    return ch->next.opaque;
}

void ottgpu_channel_destroy(ottgpu_channel *ch)
{
    if (!ch) return;
    cudaSetDevice(ch->gpu);
    cudaStreamSynchronize(ch->stream);
    if (ch->solo) ottgpu_batch_destroy(ch->solo);
    if (ch->member_of)                                 // destroyed before its batch: the batch must not touch it any more
        for (ottgpu_channel *&p : ch->member_of->ch)
            if (p == ch) p = nullptr;
    if (ch->own_arena) cudaFree(ch->arena);
    cudaFree(ch->deint);
    for (auto &set : ch->stage)
        for (uint8_t *p : set) cudaFree(p);
    cudaFree(ch->maps_pool);
    if (ch->own_stream) cudaStreamDestroy(ch->stream);
    delete ch;
}

int ottgpu_channel_submit(ottgpu_channel *ch, const ottgpu_frame_in *in, ottgpu_frame_out *out)
{
    if (!ch || !in || !out) return fail(OTTGPU_ERR_ARG, "null argument");
    if (!in->data) return fail(OTTGPU_ERR_ARG, "frame without data");
    return ottgpu_batch_submit(ch->solo, in, out, OTTGPU_SUBMIT_SYNC);
}

/* ---------------------------------------------------------------------------------------------------- batch */
static int batch_build(ottgpu_batch *b, ottgpu_channel **channels, int n);

int ottgpu_batch_create(ottgpu_channel **channels, int n, ottgpu_batch **out)
{
    if (!channels || n < 1 || n > 65535 || !out) return fail(OTTGPU_ERR_ARG, "bad batch arguments");
    *out = nullptr;
    for (int i = 0; i < n; ++i) {
        if (!channels[i]) return fail(OTTGPU_ERR_ARG, "null channel in batch");
        if (channels[i]->gpu != channels[0]->gpu) return fail(OTTGPU_ERR_ARG, "channels of one batch must share a GPU");
    }
    ottgpu_batch *b = new ottgpu_batch;
    const int rc = batch_build(b, channels, n);
    if (rc != OTTGPU_OK) {
        const std::string why = t_err;                 // destroy() must not clobber the reason
        ottgpu_batch_destroy(b);
        t_err = why;
        return rc;
    }
    *out = b;
    return OTTGPU_OK;
}

static int batch_build(ottgpu_batch *b, ottgpu_channel **channels, int n)
{
    b->ch.assign(channels, channels + n);
    b->gpu = channels[0]->gpu;
    b->stream = channels[0]->stream;
    CK(cudaSetDevice(b->gpu));
    // One stream per batch.  Channels that created their own stream adopt the batch's (their own is retired); a stream
    // shared this way is deliberately never destroyed, so no channel or batch can be left with a dangling handle.
    for (int i = 1; i < n; ++i) {
        ottgpu_channel *c = channels[i];
        if (c->stream == b->stream) continue;
        if (!c->own_stream) return fail(OTTGPU_ERR_ARG, "channels of one batch must use the same cudaStream_t");
        CK(cudaStreamSynchronize(c->stream));
        cudaStreamDestroy(c->stream);
        c->stream = b->stream;
        c->own_stream = false;
        if (c->solo) c->solo->stream = b->stream;
    }
    if (n > 1) channels[0]->own_stream = false;
    {
        // one upload arena for the whole batch when every channel has the same frame size and none holds frames yet
        bool same = true, fresh = true;
        for (int i = 0; i < n; ++i) {
            same = same && channels[i]->src_bytes == channels[0]->src_bytes;
            for (bool used : channels[i]->slot_used) fresh = fresh && !used;
            fresh = fresh && !channels[i]->prev.valid && !channels[i]->cur.valid && !channels[i]->next.valid;
        }
        if (n > 1 && same && fresh) {
            b->arena_stride = arena_stride_for(channels[0]->src_bytes);
            CK(cudaMalloc((void **)&b->arena, (size_t)ottgpu_channel::kSlots * n * b->arena_stride + 64));
            for (int i = 0; i < n; ++i) {
                ottgpu_channel *c = channels[i];
                if (c->own_arena) cudaFree(c->arena);
                c->own_arena = false;
                c->arena = b->arena;
                c->arena_stride = b->arena_stride;
                c->arena_n = n;
                c->arena_idx = i;
                c->member_of = b;
                c->maps.clear();                           // sets keyed on the old arena are stale
            }
        }
    }
    CK(cudaStreamCreateWithFlags(&b->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&b->readback_stream, cudaStreamNonBlocking));
    std::vector<Item> all, big;
    for (int bps = 1; bps <= 2; ++bps)
        for (int i = 0; i < n; ++i) {
            const Plan &p = *channels[i]->plan;
            if (p.host_copy.bps != bps) continue;
            for (Item it : p.items) { it.chan = (uint16_t)i; all.push_back(it); b->n_items_bps[bps - 1]++; }
            for (Item it : p.big_items) { it.chan = (uint16_t)i; big.push_back(it); b->n_big_items_bps[bps - 1]++; }
        }
    for (int i = 0; i < n; ++i) {
        const Plan &p = *channels[i]->plan;
        for (int k = 0; k < 3; ++k) {
            b->src_max[k] = std::max(b->src_max[k], p.src_max[k]);
            b->mid_max[k] = std::max(b->mid_max[k], p.mid_max[k]);
            b->vc_max[k] = std::max(b->vc_max[k], p.vc_max[k]);
        }
        for (int r = 0; r < channels[i]->cfg.n_rungs && r < kMaxRungs; ++r)
            b->stage_stride[r] = std::max(b->stage_stride[r], arena_stride_for(p.layout[r].bytes + 64));
        b->max_thumb_items += (int)p.thumb_items.size();
        b->thumb_offset.push_back(b->thumb_stage_bytes);
        if (p.thumb_slot >= 0) b->thumb_stage_bytes += (p.layout[p.thumb_slot].bytes + 63) & ~(size_t)63;
    }
    b->n_items = (int)all.size();
    CK(cudaMalloc((void **)&b->d_items, std::max<size_t>(all.size(), 1) * sizeof(Item)));
    if (!all.empty()) CK(cudaMemcpy(b->d_items, all.data(), all.size() * sizeof(Item), cudaMemcpyHostToDevice));
    b->n_big_items = (int)big.size();
    CK(cudaMalloc((void **)&b->d_big_items, std::max<size_t>(big.size(), 1) * sizeof(Item)));
    if (!big.empty()) CK(cudaMemcpy(b->d_big_items, big.data(), big.size() * sizeof(Item), cudaMemcpyHostToDevice));
    for (ottgpu_batch::Set &st : b->set) {
        CK(cudaMalloc((void **)&st.d_thumb_items, std::max(b->max_thumb_items, 1) * sizeof(Item)));
        CK(cudaMallocHost((void **)&st.h_thumb_items, std::max(b->max_thumb_items, 1) * sizeof(Item)));
        CK(cudaMalloc((void **)&st.d_binds, n * sizeof(Binding)));
        CK(cudaMallocHost((void **)&st.h_binds, n * sizeof(Binding)));
        CK(cudaMalloc((void **)&st.d_jobs, n * sizeof(YadifJob)));
        CK(cudaMallocHost((void **)&st.h_jobs, n * sizeof(YadifJob)));
        CK(cudaMalloc((void **)&st.d_counters, 6 * sizeof(int)));
        CK(cudaMallocHost((void **)&st.h_thumbs, std::max<size_t>(b->thumb_stage_bytes, 64)));
        CK(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&st.uploaded, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&st.computed, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&st.read_back, cudaEventDisableTiming));
        CK(cudaEventCreate(&st.k0));
        CK(cudaEventCreate(&st.k1));
        CK(cudaEventCreate(&st.y0));
        CK(cudaEventCreate(&st.y1));
    }
    return OTTGPU_OK;
}

void ottgpu_batch_destroy(ottgpu_batch *b)
{
    if (!b) return;
    cudaSetDevice(b->gpu);
    if (b->stream) cudaStreamSynchronize(b->stream);
    if (b->copy_stream) {
        cudaStreamSynchronize(b->copy_stream);
        cudaStreamDestroy(b->copy_stream);
    }
    if (b->readback_stream) {
        cudaStreamSynchronize(b->readback_stream);
        cudaStreamDestroy(b->readback_stream);
    }
    if (b->arena) {
        for (ottgpu_channel *c : b->ch)
            if (c && c->arena == b->arena) {               // the channel outlives the batch: it starts over with an arena of its own
                c->arena = nullptr;
                c->member_of = nullptr;
                c->maps.clear();
                for (bool &u : c->slot_used) u = false;
                c->slot_cooling = -1;
                c->prev = c->cur = c->next = FrameRef();
            }
        cudaFree(b->arena);
    }
    cudaFree(b->d_items);
    cudaFree(b->d_big_items);
    for (auto &per_set : b->stage_arena)
        for (uint8_t *a : per_set) cudaFree(a);
    for (ottgpu_batch::Set &st : b->set) {
        cudaFree(st.d_thumb_items);
        cudaFree(st.d_binds);
        cudaFree(st.d_jobs);
        cudaFree(st.d_counters);
        cudaFreeHost(st.h_binds);
        cudaFreeHost(st.h_jobs);
        cudaFreeHost(st.h_thumb_items);
        cudaFreeHost(st.h_thumbs);
        if (st.done) cudaEventDestroy(st.done);
        if (st.uploaded) cudaEventDestroy(st.uploaded);
        if (st.computed) cudaEventDestroy(st.computed);
        if (st.read_back) cudaEventDestroy(st.read_back);
        if (st.k0) cudaEventDestroy(st.k0);
        if (st.k1) cudaEventDestroy(st.k1);
        if (st.y0) cudaEventDestroy(st.y0);
        if (st.y1) cudaEventDestroy(st.y1);
    }
    delete b;
}

int ottgpu_batch_last_launches(const ottgpu_batch *b) { return b ? b->last_launches : 0; }
int ottgpu_batch_last_items(const ottgpu_batch *b) { return b ? b->last_items : 0; }
int ottgpu_batch_last_uploads(const ottgpu_batch *b) { return b ? b->last_uploads : 0; }
int ottgpu_batch_last_readbacks(const ottgpu_batch *b) { return b ? b->last_readbacks : 0; }

}  // extern "C"

namespace {
// the step that used this descriptor set has finished: harvest its kernel time
int retire(ottgpu_batch *b, ottgpu_batch::Set &st)
{
    if (!st.in_flight) return OTTGPU_OK;
    CK(cudaEventSynchronize(st.done));
    if (st.has_readback) CK(cudaEventSynchronize(st.read_back));
    st.has_readback = false;
    for (const ottgpu_batch::Set::PendingThumb &t : st.pending_thumbs) std::memcpy(t.dst, t.src, t.bytes);
    st.pending_thumbs.clear();
    if (st.timed) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, st.k0, st.k1));
        b->ms_sum += ms;
        b->ms_count++;
        st.timed = false;
    }
    if (st.timed_yadif) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, st.y0, st.y1));
        b->yadif_ms_sum += ms;
        b->yadif_ms_count++;
        st.timed_yadif = false;
    }
    st.in_flight = false;
    return OTTGPU_OK;
}
}  // namespace

extern "C" {

int ottgpu_batch_wait(ottgpu_batch *b)
{
    if (!b) return fail(OTTGPU_ERR_ARG, "null batch");
    CK(cudaSetDevice(b->gpu));
    CK(cudaStreamSynchronize(b->stream));
    CK(cudaStreamSynchronize(b->readback_stream));
    for (ottgpu_batch::Set &st : b->set) {
        int rc = retire(b, st);
        if (rc) return rc;
    }
    return OTTGPU_OK;
}

int ottgpu_batch_timing(ottgpu_batch *b, int enable)
{
    if (!b) return fail(OTTGPU_ERR_ARG, "null batch");
    int rc = ottgpu_batch_wait(b);
    if (rc) return rc;
    b->timing = enable != 0;
    b->ms_sum = b->yadif_ms_sum = 0.0;
    b->ms_count = b->yadif_ms_count = 0;
    return OTTGPU_OK;
}

int ottgpu_batch_timing_read(ottgpu_batch *b, double *ms_sum, int *launches)
{
    if (!b || !ms_sum || !launches) return fail(OTTGPU_ERR_ARG, "null argument");
    *ms_sum = b->ms_sum;
    *launches = b->ms_count;
    return OTTGPU_OK;
}

int ottgpu_batch_timing_read_yadif(ottgpu_batch *b, double *ms_sum, int *launches)
{
    if (!b || !ms_sum || !launches) return fail(OTTGPU_ERR_ARG, "null argument");
    *ms_sum = b->yadif_ms_sum;
    *launches = b->yadif_ms_count;
    return OTTGPU_OK;
}

}  // extern "C"

namespace {
int batch_submit_body(ottgpu_batch *b, const ottgpu_frame_in *in, ottgpu_frame_out *out, int flags);

// what a failed submit must put back so that the batch and its channels look as if the call had never been made
struct ChannelState {
    FrameRef prev, cur, next;
    bool slot_used[ottgpu_channel::kSlots];
    int slot_cooling, thumb_count;
};
}  // namespace

extern "C" {

int ottgpu_batch_submit(ottgpu_batch *b, const ottgpu_frame_in *in, ottgpu_frame_out *out, int flags)
{
    if (!b || !in || !out) return fail(OTTGPU_ERR_ARG, "null argument");
    CK(cudaSetDevice(b->gpu));
    const int n = (int)b->ch.size(), turn = b->turn;
    for (int i = 0; i < n; ++i)
        if (!b->ch[i]) return fail(OTTGPU_ERR_STATE, "a channel of this batch has been destroyed");
    std::vector<ChannelState> saved(n);
    for (int i = 0; i < n; ++i) {
        const ottgpu_channel *c = b->ch[i];
        saved[i].prev = c->prev; saved[i].cur = c->cur; saved[i].next = c->next;
        std::memcpy(saved[i].slot_used, c->slot_used, sizeof(c->slot_used));
        saved[i].slot_cooling = c->slot_cooling;
        saved[i].thumb_count = c->thumb_count;
        out[i].thumbnail = nullptr;                     // so that a failure half way through frees only what this call allocated
    }
    const int rc = batch_submit_body(b, in, out, flags);
    if (rc == OTTGPU_OK) return rc;
    // Roll back: yadif windows, upload slots and thumbnail counters as before the call, the thumbnails this call
    // malloc()ed released, the descriptor set handed back.  Uploads already queued on the copy stream only touch slots
    // that are free again after the roll-back; a later upload into the same slot is ordered behind them on that stream.
    const std::string why = t_err;
    for (int i = 0; i < n; ++i) {
        ottgpu_channel *c = b->ch[i];
        c->prev = saved[i].prev; c->cur = saved[i].cur; c->next = saved[i].next;
        std::memcpy(c->slot_used, saved[i].slot_used, sizeof(c->slot_used));
        c->slot_cooling = saved[i].slot_cooling;
        c->thumb_count = saved[i].thumb_count;
        if (out[i].thumbnail) std::free(out[i].thumbnail);
        out[i].thumbnail = nullptr;
        out[i].produced = 0;
        out[i].opaque = nullptr;
        out[i].n_released = 0;
    }
    ottgpu_batch::Set &st = b->set[turn];
    cudaStreamSynchronize(b->copy_stream);              // whatever part of the step was enqueued no longer reads the descriptors
    cudaStreamSynchronize(b->stream);
    cudaGetLastError();
    if (!st.in_flight) {                                // failed before the step was enqueued as a whole
        st.pending_thumbs.clear();
        st.has_readback = false;
        b->turn = turn;
    }
    t_err = why;
    return rc;
}

}  // extern "C"

namespace {

int batch_submit_body(ottgpu_batch *b, const ottgpu_frame_in *in, ottgpu_frame_out *out, int flags)
{
    const int turn = b->turn;
    ottgpu_batch::Set &st = b->set[turn];
    b->turn ^= 1;
    {                                                   // at most two steps in flight: this set's previous user must be done
        int rc = retire(b, st);
        if (rc) return rc;
        st.pending_thumbs.clear();                      // (left over only if a previous submit failed half way)
    }
    const int n = (int)b->ch.size();
    cudaStream_t s = b->stream;
    // uploads of this step may overwrite slots last read two steps ago: order the copy stream after that step only
    // (this set's previous use was retired above; the other set = the previous step may still be running)
    int n_jobs = 0, n_thumb_items = 0, n_thumb_items16 = 0, max_w = 0, max_h = 0, live = 0;
    std::vector<ottgpu_batch::Upload> &uploads = b->uploads;
    uploads.clear();
    bool all_fast = true;
    b->readbacks.clear();
    b->last_launches = b->last_items = b->last_uploads = 0;

    for (int i = 0; i < n; ++i) {
        ottgpu_channel *c = b->ch[i];
        const Plan &plan = *c->plan;
        const ottgpu_cfg &cfg = c->cfg;
        Binding &bind = st.h_binds[i];
        YadifJob &job = st.h_jobs[i];
        std::memset(&bind, 0, sizeof(bind));
        std::memset(&job, 0, sizeof(job));
        bind.plan = plan.dev;
        ottgpu_frame_out &o = out[i];
        o.produced = 0;
        o.opaque = nullptr;
        o.pts = o.dts = 0;
        o.interlaced = 0;
        o.tff = 1;
        o.thumbnail = nullptr;
        o.n_released = 0;
        if (!in[i].data) continue;

        // ---- 1. bring the frame to the device
        FrameRef f;
        f.valid = true;
        f.interlaced = in[i].interlaced != 0;
        f.tff = in[i].tff != 0;
        f.pts = in[i].pts;
        f.dts = in[i].dts;
        f.opaque = in[i].opaque;
        if (in[i].mem == OTTGPU_MEM_HOST) {
            f.slot = take_slot(c);
            if (f.slot < 0) return fail(OTTGPU_ERR_NOMEM, "no device slot for the incoming frame");
            f.dev = c->slot_ptr(f.slot);
            uploads.push_back({c->slot_ptr(f.slot), in[i].data, c->src_bytes});
        } else if (in[i].mem == OTTGPU_MEM_DEVICE) {
            f.dev = static_cast<const uint8_t *>(in[i].data);
        } else {
            return fail(OTTGPU_ERR_ARG, "bad input memory kind");
        }

        // ---- 2. yadif frame scheduling (yadif_common.c ff_yadif_filter_frame)
        FrameRef target;
        bool run_yadif = false;
        if (cfg.deint == OTTGPU_DEINT_BYPASS) {
            target = f;
            c->next = f;                               // held until the step completes
        } else {
            FrameRef old_prev = c->prev;
            c->prev = c->cur;
            c->cur = c->next;
            c->next = f;
            if (!c->cur.valid) c->cur = c->next;
            release_frame(c, old_prev, &o);
            if (!c->prev.valid) continue;              // first frame only primes the window
            target = c->cur;
            run_yadif = cfg.deint == OTTGPU_DEINT_ALL || target.interlaced;
            if (run_yadif && cfg.src_fmt == OTTGPU_FMT_P010)
                return fail(OTTGPU_ERR_ARG, "interlaced-flagged P010 frame: yadif needs planar samples");
        }
        live++;
        o.produced = 1;
        o.opaque = target.opaque;
        o.pts = target.pts;
        o.dts = target.dts;

        // ---- 3. where do the rungs go
        uint32_t skip = 0;
        for (int r = 0; r < cfg.n_rungs; ++r) {
            uint8_t *base;
            if (o.dst_mem[r] == OTTGPU_MEM_DEVICE) {
                base = static_cast<uint8_t *>(o.dst[r]);
            } else {
                int rc = ensure_dev(&b->stage_arena[turn][r], b->stage_stride[r] * n);
                if (rc) return rc;
                base = b->stage_arena[turn][r] + (size_t)i * b->stage_stride[r];
                b->readbacks.push_back({o.dst[r], base, plan.layout[r].bytes, r});
            }
            if (!base || (o.dst_mem[r] != OTTGPU_MEM_DEVICE && !o.dst[r])) return fail(OTTGPU_ERR_ARG, "missing destination buffer");
            for (int p = 0; p < plan.layout[r].planes; ++p) {
                bind.dst[r][p] = base + plan.layout[r].offset[p];
                bind.dst_pitch[r][p] = plan.layout[r].pitch[p];
            }
        }
        const uint8_t *src_base = target.dev;
        if (run_yadif) {
            uint8_t *ydst;
            if (c->direct_rung >= 0) {
                ydst = bind.dst[c->direct_rung][0];    // the full-size rung IS the deinterlaced frame
                skip |= 1u << c->direct_rung;
            } else {
                int rc = ensure_dev(&c->deint, c->src_bytes);
                if (rc) return rc;
                ydst = c->deint;
            }
            job.prev = c->prev.dev;
            job.cur = c->cur.dev;
            job.next = c->next.dev;
            job.dst = ydst;
            job.width = cfg.src_width;
            job.height = cfg.src_height;
            job.bps = fmt_bps(cfg.src_fmt);
            job.tff = target.interlaced ? target.tff : 1;
            job.parity = job.tff ^ 1;
            job.active = 1;
            job.fast = job.bps == 1 && cfg.src_width % 16 == 0 && (cfg.src_width * cfg.src_height) % 8 == 0 &&
                       (((cfg.src_width / 2) * ((cfg.src_height + 1) / 2)) % 8) == 0 &&
                       (((uintptr_t)job.prev | (uintptr_t)job.cur | (uintptr_t)job.next | (uintptr_t)job.dst) & 7) == 0;
            all_fast = all_fast && job.fast;
            n_jobs++;
            max_w = std::max(max_w, cfg.src_width);
            max_h = std::max(max_h, cfg.src_height);
            src_base = ydst;
        }
        plane_ptrs(cfg.src_fmt, src_base, cfg.src_width, cfg.src_height, bind.src, bind.src_pitch);
        {
            int rc = frame_tensor_maps(c, src_base, b->copy_stream, &bind.tmaps, &bind.src_z);
            if (rc) return rc;
        }
        bind.active = 1;
        bind.skip_mask = skip;

        // ---- 4. thumbnail cadence (transvideo.c:2159-2186)
        if (plan.thumb_slot >= 0 && ++c->thumb_count == cfg.thumb_period) {
            c->thumb_count = 0;
            const int ts = plan.thumb_slot;
            int rc = ensure_dev(&c->stage[turn][ts], plan.layout[ts].bytes);
            if (rc) return rc;
            for (int p = 0; p < plan.layout[ts].planes; ++p) {
                bind.dst[ts][p] = c->stage[turn][ts] + plan.layout[ts].offset[p];
                bind.dst_pitch[ts][p] = plan.layout[ts].pitch[p];
            }
            bind.thumb_active = 1;
            // the consumer frees the thumbnail with free() (transvideo.c:254-255), so it is malloc'ed; the device-to-host copy
            // lands in pinned memory (a copy into pageable memory would block this thread behind every queued read-back)
            // and is moved into the malloc'ed buffer when the step retires
            void *host = std::malloc(plan.layout[ts].bytes);
            if (!host) return fail(OTTGPU_ERR_NOMEM, "malloc(thumbnail) failed");
            o.thumbnail = host;
            uint8_t *landing = st.h_thumbs + b->thumb_offset[i];
            b->readbacks.push_back({landing, c->stage[turn][ts], plan.layout[ts].bytes, kMaxRungs});
            st.pending_thumbs.push_back({host, landing, plan.layout[ts].bytes});
            // 8-bit channels' items fill the list from the front, 16-bit ones from the back
            for (Item it : plan.thumb_items) {
                it.chan = (uint16_t)i;
                if (plan.host_copy.bps == 2) st.h_thumb_items[b->max_thumb_items - ++n_thumb_items16] = it;
                else st.h_thumb_items[n_thumb_items++] = it;
            }
        }

        if (cfg.deint == OTTGPU_DEINT_BYPASS) {        // nothing is kept across calls
            FrameRef done = c->next;
            c->next = FrameRef();
            release_frame(c, done, &o);
        }
    }

    // ---- 5. enqueue: descriptors, yadif, ladder, thumbnails, read-backs
    // The small descriptor copies and the reset of the tile schedulers always ride on the copy stream: they belong to this
    // step's own descriptor set (its previous user was retired above), so they overlap the previous step's kernels instead
    // of sitting between two launches on the compute stream; with host frames in flight they also stay behind this step's
    // uploads instead of queueing behind the NEXT step's bulk uploads in the H2D copy engine.
    cudaStream_t ds = b->copy_stream;
    // Frame uploads: consecutive host buffers going to consecutive arena slots are merged into one copy (the steady state
    // of a batch fed from one pinned pool: ONE cudaMemcpyAsync per step instead of one per channel)
    for (size_t u = 0; u < uploads.size();) {
        size_t v = u + 1;
        size_t bytes = uploads[u].bytes;
        while (v < uploads.size() && b->arena_stride && uploads[v].dst == uploads[v - 1].dst + b->arena_stride &&
               (const uint8_t *)uploads[v].src == (const uint8_t *)uploads[v - 1].src + b->arena_stride &&
               same_host_pool(uploads[u].src, uploads[v].src)) {
            bytes = (size_t)(uploads[v].dst - uploads[u].dst) + uploads[v].bytes;
            ++v;
        }
        CK(cudaMemcpyAsync(uploads[u].dst, uploads[u].src, bytes, cudaMemcpyHostToDevice, ds));
        b->last_uploads++;
        u = v;
    }
    CK(cudaMemcpyAsync(st.d_binds, st.h_binds, n * sizeof(Binding), cudaMemcpyHostToDevice, ds));
    if (n_jobs) CK(cudaMemcpyAsync(st.d_jobs, st.h_jobs, n * sizeof(YadifJob), cudaMemcpyHostToDevice, ds));
    if (n_thumb_items) CK(cudaMemcpyAsync(st.d_thumb_items, st.h_thumb_items, n_thumb_items * sizeof(Item), cudaMemcpyHostToDevice, ds));
    if (n_thumb_items16)
        CK(cudaMemcpyAsync(st.d_thumb_items + (b->max_thumb_items - n_thumb_items16), st.h_thumb_items + (b->max_thumb_items - n_thumb_items16),
                           n_thumb_items16 * sizeof(Item), cudaMemcpyHostToDevice, ds));
    const bool any_ladder = (live && (b->n_items || b->n_big_items)) || n_thumb_items || n_thumb_items16;
    if (any_ladder) CK(cudaMemsetAsync(st.d_counters, 0, 6 * sizeof(int), ds));
    CK(cudaEventRecord(st.uploaded, ds));
    CK(cudaStreamWaitEvent(s, st.uploaded, 0));
    if (n_jobs) {
        if (b->timing) CK(cudaEventRecord(st.y0, s));
        CK(launch_yadif(st.d_jobs, n, max_w, max_h, all_fast, s));
        if (b->timing) {
            CK(cudaEventRecord(st.y1, s));
            st.timed_yadif = true;
        }
        b->last_launches++;
    }
    auto make_launch = [&](const Item *items, int count, int which, int bps) {
        LadderLaunch L;
        L.items = items;
        L.binds = st.d_binds;
        L.n_items = count;
        L.bps = bps;
        L.counter = st.d_counters + 2 * which + (bps - 1);
        L.src_max = b->src_max[which];
        L.mid_max = b->mid_max[which];
        L.vc_max = b->vc_max[which];
        L.double_buffer = 0;                             // decided by launch_ladder from the shared-memory budget
        return L;
    };
    if (live && b->n_items) {
        if (b->timing) CK(cudaEventRecord(st.k0, s));
        for (int k = 0, at = 0; k < 2; at += b->n_items_bps[k], ++k)
            if (b->n_items_bps[k]) {
                CK(launch_ladder(make_launch(b->d_items + at, b->n_items_bps[k], 0, k + 1), s));
                b->last_launches++;
            }
        if (b->timing) {
            CK(cudaEventRecord(st.k1, s));
            st.timed = true;
        }
        b->last_items += b->n_items;
    }
    if (live && b->n_big_items) {
        for (int k = 0, at = 0; k < 2; at += b->n_big_items_bps[k], ++k)
            if (b->n_big_items_bps[k]) {
                CK(launch_ladder(make_launch(b->d_big_items + at, b->n_big_items_bps[k], 1, k + 1), s));
                b->last_launches++;
            }
        b->last_items += b->n_big_items;
    }
    if (n_thumb_items) {
        CK(launch_ladder(make_launch(st.d_thumb_items, n_thumb_items, 2, 1), s));
        b->last_launches++;
        b->last_items += n_thumb_items;
    }
    if (n_thumb_items16) {
        CK(launch_ladder(make_launch(st.d_thumb_items + (b->max_thumb_items - n_thumb_items16), n_thumb_items16, 2, 2), s));
        b->last_launches++;
        b->last_items += n_thumb_items16;
    }
    CK(cudaEventRecord(st.done, s));
    if (!b->readbacks.empty()) {
        // results for host destinations leave on their own stream: the next step's kernels do not wait for them
        CK(cudaEventRecord(st.computed, s));
        CK(cudaStreamWaitEvent(b->readback_stream, st.computed, 0));
        // rung by rung; a run of channels whose surfaces and host buffers advance by constant strides is one copy (linear when
        // both strides equal the surface size, else a pitched copy with one "row" per channel)
        std::vector<ottgpu_batch::Readback> &rb = b->readbacks;
        std::stable_sort(rb.begin(), rb.end(), [](const ottgpu_batch::Readback &x, const ottgpu_batch::Readback &y) { return x.rung < y.rung; });
        b->last_readbacks = 0;
        for (size_t u = 0; u < rb.size();) {
            size_t v = u + 1;
            ptrdiff_t ds = 0, hs = 0;
            if (v < rb.size() && rb[v].rung == rb[u].rung && rb[v].bytes == rb[u].bytes) {
                ds = rb[v].dev - rb[u].dev;
                hs = (const uint8_t *)rb[v].host - (const uint8_t *)rb[u].host;
                if (ds >= (ptrdiff_t)rb[u].bytes && hs >= (ptrdiff_t)rb[u].bytes && same_host_pool(rb[u].host, rb[v].host)) {
                    ++v;
                    while (v < rb.size() && rb[v].rung == rb[u].rung && rb[v].bytes == rb[u].bytes && rb[v].dev - rb[v - 1].dev == ds &&
                           (const uint8_t *)rb[v].host - (const uint8_t *)rb[v - 1].host == hs && same_host_pool(rb[u].host, rb[v].host))
                        ++v;
                }
            }
            const size_t count = v - u;
            if (count == 1)
                CK(cudaMemcpyAsync(rb[u].host, rb[u].dev, rb[u].bytes, cudaMemcpyDeviceToHost, b->readback_stream));
            else if (ds == hs)          // the gaps between the surfaces are copied too: both sides own them (arena / pool padding)
                CK(cudaMemcpyAsync(rb[u].host, rb[u].dev, (size_t)ds * (count - 1) + rb[u].bytes, cudaMemcpyDeviceToHost, b->readback_stream));
            else
                CK(cudaMemcpy2DAsync(rb[u].host, (size_t)hs, rb[u].dev, (size_t)ds, rb[u].bytes, count, cudaMemcpyDeviceToHost, b->readback_stream));
            b->last_readbacks++;
            u = v;
        }
        CK(cudaEventRecord(st.read_back, b->readback_stream));
        st.has_readback = true;
    }
    st.in_flight = true;
    if (flags & OTTGPU_SUBMIT_ASYNC) return OTTGPU_OK;
    return ottgpu_batch_wait(b);
}

}  // namespace

/* ---------------------------------------------------------------------------------------------------- pools */

struct ottgpu_pool {
    int gpu = 0, mem = 0;
    size_t size = 0;
    uint8_t *base = nullptr;
    std::vector<void *> free_list;
    std::vector<bool> taken;
    int count = 0;
    std::mutex mu;
};

extern "C" {

int ottgpu_pool_create(int gpu, int mem, int buffer_count, size_t buffer_size, ottgpu_pool **out)
{
    if (!out || buffer_count < 1 || !buffer_size) return fail(OTTGPU_ERR_ARG, "bad pool arguments");
    *out = nullptr;
    int rc = prepare_device(gpu);
    if (rc) return rc;
    std::unique_ptr<ottgpu_pool> p(new ottgpu_pool);
    p->gpu = gpu;
    p->mem = mem;
    p->count = buffer_count;
    p->size = (buffer_size + 255) & ~(size_t)255;      // every buffer 256-byte aligned: vector stores stay legal
    if (mem != OTTGPU_MEM_DEVICE && mem != OTTGPU_MEM_HOST && mem != OTTGPU_MEM_HOST_WC) return fail(OTTGPU_ERR_ARG, "bad pool memory kind");
    cudaError_t e = mem == OTTGPU_MEM_DEVICE ? cudaMalloc((void **)&p->base, p->size * buffer_count)
                    : cudaHostAlloc((void **)&p->base, p->size * buffer_count, mem == OTTGPU_MEM_HOST_WC ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(OTTGPU_ERR_NOMEM, "pool allocation failed");
    }
    p->taken.assign(buffer_count, false);
    for (int i = buffer_count - 1; i >= 0; --i) p->free_list.push_back(p->base + (size_t)i * p->size);
    {
        std::lock_guard<std::mutex> lock(g_pools_mu);
        (mem == OTTGPU_MEM_DEVICE ? g_pools : g_host_pools).push_back({p->base, p->size, buffer_count, gpu});
    }
    *out = p.release();
    return OTTGPU_OK;
}

void *ottgpu_pool_take(ottgpu_pool *p)
{
    if (!p) return nullptr;
    std::lock_guard<std::mutex> lock(p->mu);
    if (p->free_list.empty()) return nullptr;
    void *b = p->free_list.back();
    p->free_list.pop_back();
    p->taken[((uint8_t *)b - p->base) / p->size] = true;
    return b;
}

int ottgpu_pool_return(ottgpu_pool *p, void *buffer)
{
    if (!p || !buffer) return fail(OTTGPU_ERR_ARG, "null argument");
    std::lock_guard<std::mutex> lock(p->mu);
    const uint8_t *b = static_cast<uint8_t *>(buffer);
    if (b < p->base || b >= p->base + p->size * p->count || (b - p->base) % p->size) return fail(OTTGPU_ERR_ARG, "buffer is not from this pool");
    const size_t idx = (b - p->base) / p->size;
    if (!p->taken[idx]) return fail(OTTGPU_ERR_STATE, "buffer returned twice");
    p->taken[idx] = false;
    p->free_list.push_back(buffer);
    return OTTGPU_OK;
}

int ottgpu_pool_unused(ottgpu_pool *p)
{
    if (!p) return 0;
    std::lock_guard<std::mutex> lock(p->mu);
    return (int)p->free_list.size();
}

void ottgpu_pool_destroy(ottgpu_pool *p)
{
    if (!p) return;
    cudaSetDevice(p->gpu);
    {
        std::lock_guard<std::mutex> lock(g_pools_mu);
        std::vector<PoolSpan> &reg = p->mem == OTTGPU_MEM_DEVICE ? g_pools : g_host_pools;
        for (size_t i = 0; i < reg.size(); ++i)
            if (reg[i].base == p->base) { reg.erase(reg.begin() + i); break; }
    }
    if (p->mem == OTTGPU_MEM_DEVICE) cudaFree(p->base);
    else cudaFreeHost(p->base);
    delete p;
}

int ottgpu_memcpy(int gpu, void *dst, int dst_mem, const void *src, int src_mem, size_t bytes)
{
    if (!dst || !src) return fail(OTTGPU_ERR_ARG, "null argument");
    CK(cudaSetDevice(gpu));
    const cudaMemcpyKind kind = dst_mem == OTTGPU_MEM_DEVICE
                                    ? (src_mem == OTTGPU_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)
                                    : (src_mem == OTTGPU_MEM_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost);
    CK(cudaMemcpy(dst, src, bytes, kind));
    return OTTGPU_OK;
}

/* ----------------------------------------------------------------------------------------------- single shots */
int ottgpu_scale_frame(int gpu, const void *src, int src_fmt, int src_w, int src_h, void *dst,
                       const ottgpu_rung_cfg *rung, int filter, int target)
{
    if (!src || !dst || !rung) return fail(OTTGPU_ERR_ARG, "null argument");
    ottgpu_cfg cfg;
    std::memset(&cfg, 0, sizeof(cfg));
    cfg.gpu = gpu;
    cfg.src_width = src_w;
    cfg.src_height = src_h;
    cfg.src_fmt = src_fmt;
    cfg.n_rungs = 1;
    cfg.rung[0] = *rung;
    cfg.filter = filter;
    cfg.target = target;
    cfg.deint = OTTGPU_DEINT_BYPASS;
    ottgpu_channel *ch = nullptr;
    int rc = ottgpu_channel_create(&cfg, &ch);
    if (rc) return rc;
    ottgpu_frame_in in;
    std::memset(&in, 0, sizeof(in));
    in.data = src;
    in.mem = OTTGPU_MEM_HOST;
    ottgpu_frame_out out;
    std::memset(&out, 0, sizeof(out));
    out.dst[0] = dst;
    out.dst_mem[0] = OTTGPU_MEM_HOST;
    rc = ottgpu_channel_submit(ch, &in, &out);
    if (rc == OTTGPU_OK && !out.produced) rc = fail(OTTGPU_ERR_STATE, "no frame produced");
    ottgpu_channel_destroy(ch);
    return rc;
}

int ottgpu_yadif_frame(int gpu, const void *prev, const void *cur, const void *next, void *dst, int fmt, int width,
                       int height, int interlaced, int tff)
{
    if (!prev || !cur || !next || !dst) return fail(OTTGPU_ERR_ARG, "null argument");
    if (fmt != OTTGPU_FMT_I420 && fmt != OTTGPU_FMT_I420P10) return fail(OTTGPU_ERR_ARG, "yadif needs I420 or I420P10");
    if (width < 8 || height < 4) return fail(OTTGPU_ERR_ARG, "frame too small");
    int rc = prepare_device(gpu);
    if (rc) return rc;
    const size_t bytes = packed_frame_bytes(fmt, width, height);
    uint8_t *d = nullptr;
    YadifJob *dj = nullptr;
    if (cudaMalloc((void **)&d, 4 * (bytes + 256)) != cudaSuccess || cudaMalloc((void **)&dj, sizeof(YadifJob)) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(d);
        return fail(OTTGPU_ERR_NOMEM, "cudaMalloc failed");
    }
    const size_t step = (bytes + 255) & ~(size_t)255;
    YadifJob j;
    std::memset(&j, 0, sizeof(j));
    j.prev = d; j.cur = d + step; j.next = d + 2 * step; j.dst = d + 3 * step;
    j.width = width; j.height = height; j.bps = fmt_bps(fmt);
    j.tff = interlaced ? (tff != 0) : 1;
    j.parity = j.tff ^ 1;
    j.active = 1;
    j.fast = j.bps == 1 && width % 16 == 0 && (width * height) % 8 == 0 && ((width / 2) * ((height + 1) / 2)) % 8 == 0;
    cudaError_t e = cudaMemcpy(d, prev, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + step, cur, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + 2 * step, next, bytes, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(dj, &j, sizeof(j), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = launch_yadif(dj, 1, width, height, j.fast != 0, 0);
    if (e == cudaSuccess) e = cudaMemcpy(dst, d + 3 * step, bytes, cudaMemcpyDeviceToHost);
    cudaFree(d);
    cudaFree(dj);
    if (e != cudaSuccess) return fail_cuda(e, "ottgpu_yadif_frame");
    return OTTGPU_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------- decode-side converter
struct ottgpu_converter {
    int gpu = 0, pix_fmt = 0, w = 0, h = 0, target = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int32_t *d_vpos = nullptr;
    int16_t *d_vc = nullptr;
    int vtaps = 1, exact_from = 0;
    // per plane: source rows / bytes per row, destination size
    int src_rows[3] = {0, 0, 0}, row_bytes[3] = {0, 0, 0}, dw[3] = {0, 0, 0}, dh[3] = {0, 0, 0}, mode[3] = {0, 0, 0};
    size_t stage_pitch[3] = {0, 0, 0};
    static constexpr int kSets = 2;                 // a host picture is uploaded while the previous one converts
    uint8_t *stage[kSets][3] = {{nullptr}, {nullptr}};
    uint8_t *out_stage[kSets] = {nullptr, nullptr}; // packed result on the device when the caller wants it in host memory
    cudaEvent_t done[kSets] = {nullptr, nullptr};
    int phase = 0;
    size_t out_bytes = 0;
    ConvertJob *h_jobs[kSets] = {nullptr, nullptr}, *d_jobs[kSets] = {nullptr, nullptr};   // batch calls led by this converter
    int job_capacity = 0;
};

namespace {

const uint8_t kBayer8x8[64] = {36, 68, 60, 92, 34, 66, 58, 90, 100, 4,  124, 28, 98,  2,  122, 26, 52, 84, 44, 76, 50, 82,
                               42, 74, 116, 20, 108, 12, 114, 18, 106, 10, 32, 64, 56,  88, 38,  70, 62, 94, 96, 0,  120, 24,
                               102, 6,  126, 30, 48, 80, 40, 72, 54, 86, 46,  78, 112, 16, 104, 8,  118, 22, 110, 14};

void converter_free(ottgpu_converter *cv)
{
    if (!cv) return;
    cudaSetDevice(cv->gpu);
    if (cv->stream) cudaStreamSynchronize(cv->stream);
    for (int s = 0; s < ottgpu_converter::kSets; ++s) {
        for (int p = 0; p < 3; ++p) cudaFree(cv->stage[s][p]);
        cudaFree(cv->out_stage[s]);
        cudaFreeHost(cv->h_jobs[s]);
        cudaFree(cv->d_jobs[s]);
        if (cv->done[s]) cudaEventDestroy(cv->done[s]);
    }
    cudaFree(cv->d_vpos);
    cudaFree(cv->d_vc);
    if (cv->own_stream && cv->stream) cudaStreamDestroy(cv->stream);
    delete cv;
}

int converter_build(ottgpu_converter *cv, void *stream)
{
    const int w = cv->w, h = cv->h, cw = (w + 1) / 2, ch = (h + 1) / 2;
    const bool is422 = cv->pix_fmt == OTTGPU_PIX_YUV422P || cv->pix_fmt == OTTGPU_PIX_YUV422P10;
    const int bps = (cv->pix_fmt == OTTGPU_PIX_YUV420P10 || cv->pix_fmt == OTTGPU_PIX_YUV422P10) ? 2 : 1;
    cv->dw[0] = w; cv->dh[0] = h;
    cv->dw[1] = cv->dw[2] = cw; cv->dh[1] = cv->dh[2] = ch;
    cv->src_rows[0] = h;
    cv->src_rows[1] = cv->src_rows[2] = is422 ? h : ch;
    cv->row_bytes[0] = w * bps;
    cv->row_bytes[1] = cv->row_bytes[2] = cw * bps;
    cv->mode[0] = bps == 2 ? CONV_REDUCE10 : CONV_COPY8;
    cv->mode[1] = cv->mode[2] = is422 ? (bps == 2 ? CONV_VFILT10 : CONV_VFILT8) : cv->mode[0];
    cv->out_bytes = (size_t)w * h + 2 * (size_t)cw * ch;
    if (stream) cv->stream = (cudaStream_t)stream;
    else {
        CK(cudaStreamCreateWithFlags(&cv->stream, cudaStreamNonBlocking));
        cv->own_stream = true;
    }
    if (is422) {
        // the vertical chroma filter sws_getContext builds for h -> ceil(h/2) rows (chrSrcVSubSample 0 -> 1)
        const PolyphaseTable t = plan_table(h, ch, OTTGPU_FILTER_BICUBIC, cv->target, true);
        if (t.taps > 64) return fail(OTTGPU_ERR_ARG, "vertical chroma filter too long");
        cv->vtaps = t.taps;
        int uv = (h - 2 + 1) / 2;                          // chroma row uv is produced while dstY == 2*uv (swscale.c)
        cv->exact_from = uv < 0 ? 0 : uv;
        CK(cudaMalloc((void **)&cv->d_vpos, std::max<size_t>(t.pos.size() * sizeof(int32_t), 16)));
        CK(cudaMalloc((void **)&cv->d_vc, std::max<size_t>(t.coef.size() * sizeof(int16_t), 16)));
        CK(cudaMemcpy(cv->d_vpos, t.pos.data(), t.pos.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(cv->d_vc, t.coef.data(), t.coef.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
    }
    for (int p = 0; p < 3; ++p) cv->stage_pitch[p] = ((size_t)cv->row_bytes[p] + 15) & ~(size_t)15;
    for (int s = 0; s < ottgpu_converter::kSets; ++s) CK(cudaEventCreateWithFlags(&cv->done[s], cudaEventDisableTiming));
    return OTTGPU_OK;
}

}  // namespace

extern "C" {

int ottgpu_converter_create(int gpu, int pix_fmt, int width, int height, int target, void *stream, ottgpu_converter **out)
{
    if (!out) return fail(OTTGPU_ERR_ARG, "null argument");
    *out = nullptr;
    if (pix_fmt < OTTGPU_PIX_YUV420P || pix_fmt > OTTGPU_PIX_YUV422P10) return fail(OTTGPU_ERR_ARG, "unknown decoder pixel format");
    if (target != OTTGPU_TARGET_A && target != OTTGPU_TARGET_B) return fail(OTTGPU_ERR_ARG, "unknown parity target");
    if (width < 8 || height < 4 || width > 16384 || height > 16384) return fail(OTTGPU_ERR_ARG, "picture size out of range");
    int rc = prepare_device(gpu);
    if (rc) return rc;
    ottgpu_converter *cv = new ottgpu_converter();
    cv->gpu = gpu; cv->pix_fmt = pix_fmt; cv->w = width; cv->h = height; cv->target = target;
    rc = converter_build(cv, stream);
    if (rc) {
        const std::string why = t_err;
        converter_free(cv);
        t_err = why;
        return rc;
    }
    *out = cv;
    return OTTGPU_OK;
}

}  // extern "C"

namespace {

// Validates one picture, uploads host planes into the converter's staging set and fills the kernel's job.
int converter_prepare(ottgpu_converter *cv, const ottgpu_picture *in, void *dst, int dst_mem, int set, cudaStream_t stream, ConvertJob &j)
{
    if (!cv || !in || !dst) return fail(OTTGPU_ERR_ARG, "null argument");
    if ((in->mem != OTTGPU_MEM_HOST && in->mem != OTTGPU_MEM_DEVICE) || (dst_mem != OTTGPU_MEM_HOST && dst_mem != OTTGPU_MEM_DEVICE))
        return fail(OTTGPU_ERR_ARG, "unknown memory kind");
    for (int p = 0; p < 3; ++p) {
        if (!in->plane[p]) return fail(OTTGPU_ERR_ARG, "null plane");
        if (in->stride[p] < cv->row_bytes[p]) return fail(OTTGPU_ERR_ARG, "stride smaller than the row");
    }
    std::memset(&j, 0, sizeof(j));
    for (int p = 0; p < 3; ++p) {
        ConvertPlane &P = j.plane[p];
        if (in->mem == OTTGPU_MEM_HOST) {
            if (!cv->stage[set][p]) CK(cudaMalloc((void **)&cv->stage[set][p], cv->stage_pitch[p] * cv->src_rows[p] + 16));
            CK(cudaMemcpy2DAsync(cv->stage[set][p], cv->stage_pitch[p], in->plane[p], (size_t)in->stride[p], (size_t)cv->row_bytes[p],
                                 (size_t)cv->src_rows[p], cudaMemcpyHostToDevice, stream));
            P.src = cv->stage[set][p];
            P.src_pitch = (int)cv->stage_pitch[p];
        } else {
            P.src = static_cast<const uint8_t *>(in->plane[p]);
            P.src_pitch = in->stride[p];
        }
        P.w = cv->dw[p]; P.h = cv->dh[p];
        P.dst_pitch = cv->dw[p];
        P.mode = cv->mode[p];
        P.dither_offset = p == 2 ? 3 : 0;
    }
    uint8_t *d = static_cast<uint8_t *>(dst);
    if (dst_mem == OTTGPU_MEM_HOST) {
        if (!cv->out_stage[set]) CK(cudaMalloc((void **)&cv->out_stage[set], cv->out_bytes + 16));
        d = cv->out_stage[set];
    }
    j.plane[0].dst = d;
    j.plane[1].dst = d + (size_t)cv->w * cv->h;
    j.plane[2].dst = j.plane[1].dst + (size_t)cv->dw[1] * cv->dh[1];
    for (int p = 0; p < 3; ++p) {
        ConvertPlane &P = j.plane[p];
        // vector path: 16 samples per thread (8 on vertically filtered planes) moved as 8/16-byte words
        const bool vf = cv->mode[p] == CONV_VFILT8 || cv->mode[p] == CONV_VFILT10;
        const int group = vf ? 8 : 16;
        P.fast = P.w % group == 0 && ((uintptr_t)P.src % 16) == 0 && P.src_pitch % 16 == 0 && ((uintptr_t)P.dst % group) == 0;
    }
    j.vpos = cv->d_vpos; j.vc = cv->d_vc;
    j.vtaps = cv->vtaps; j.exact_from = cv->exact_from;
    j.approx = cv->target == OTTGPU_TARGET_A;
    std::memcpy(j.bayer, kBayer8x8, sizeof(j.bayer));
    return OTTGPU_OK;
}

}  // namespace

extern "C" {

int ottgpu_converter_convert(ottgpu_converter *cv, const ottgpu_picture *in, void *dst, int dst_mem, int flags)
{
    if (!cv) return fail(OTTGPU_ERR_ARG, "null argument");
    CK(cudaSetDevice(cv->gpu));
    const int set = cv->phase;
    CK(cudaEventSynchronize(cv->done[set]));                 // the conversion that last used this set has finished
    ConvertJob j;
    int rc = converter_prepare(cv, in, dst, dst_mem, set, cv->stream, j);
    if (rc) return rc;
    cv->phase ^= 1;
    CK(launch_convert(j, cv->stream));
    if (dst_mem == OTTGPU_MEM_HOST) CK(cudaMemcpyAsync(dst, cv->out_stage[set], cv->out_bytes, cudaMemcpyDeviceToHost, cv->stream));
    CK(cudaEventRecord(cv->done[set], cv->stream));
    if (!(flags & OTTGPU_SUBMIT_ASYNC)) CK(cudaStreamSynchronize(cv->stream));
    return OTTGPU_OK;
}

int ottgpu_converter_convert_batch(ottgpu_converter *const *cv, const ottgpu_picture *in, void *const *dst, int n, int dst_mem, int flags)
{
    if (!cv || !in || !dst || n < 1) return fail(OTTGPU_ERR_ARG, "null argument");
    ottgpu_converter *lead = cv[0];
    if (!lead) return fail(OTTGPU_ERR_ARG, "null converter");
    for (int i = 0; i < n; ++i) {
        if (!cv[i]) return fail(OTTGPU_ERR_ARG, "null converter");
        if (cv[i]->gpu != lead->gpu) return fail(OTTGPU_ERR_ARG, "converters of one batch must be on the same gpu");
        for (int k = 0; k < i; ++k)
            if (cv[k] == cv[i]) return fail(OTTGPU_ERR_ARG, "a converter may appear once per batch");
    }
    CK(cudaSetDevice(lead->gpu));
    cudaStream_t s = lead->stream;                            // the whole batch is ordered on the first converter's stream
    const int set = lead->phase;
    for (int i = 0; i < n; ++i) {
        if (cv[i]->phase != set) return fail(OTTGPU_ERR_STATE, "converters of one batch must have been used the same number of times");
        CK(cudaEventSynchronize(cv[i]->done[set]));
    }
    if (lead->job_capacity < n) {                             // job lists (pinned + device, one pair per set) sized on first use
        for (int k = 0; k < ottgpu_converter::kSets; ++k) {
            cudaFreeHost(lead->h_jobs[k]);
            cudaFree(lead->d_jobs[k]);
            lead->h_jobs[k] = nullptr; lead->d_jobs[k] = nullptr;
            CK(cudaMallocHost((void **)&lead->h_jobs[k], (size_t)n * sizeof(ConvertJob)));
            CK(cudaMalloc((void **)&lead->d_jobs[k], (size_t)n * sizeof(ConvertJob)));
        }
        lead->job_capacity = n;
    }
    int max_w = 0, max_rows = 0;
    for (int i = 0; i < n; ++i) {
        int rc = converter_prepare(cv[i], &in[i], dst[i], dst_mem, set, s, lead->h_jobs[set][i]);
        if (rc) return rc;
        max_w = std::max(max_w, cv[i]->w);
        max_rows = std::max(max_rows, cv[i]->dh[0] + cv[i]->dh[1] + cv[i]->dh[2]);
    }
    CK(cudaMemcpyAsync(lead->d_jobs[set], lead->h_jobs[set], (size_t)n * sizeof(ConvertJob), cudaMemcpyHostToDevice, s));
    CK(launch_convert_batch(lead->d_jobs[set], n, max_w, max_rows, s));
    for (int i = 0; i < n; ++i) {
        if (dst_mem == OTTGPU_MEM_HOST) CK(cudaMemcpyAsync(dst[i], cv[i]->out_stage[set], cv[i]->out_bytes, cudaMemcpyDeviceToHost, s));
        cv[i]->phase ^= 1;
        CK(cudaEventRecord(cv[i]->done[set], s));
    }
    if (!(flags & OTTGPU_SUBMIT_ASYNC)) CK(cudaStreamSynchronize(s));
    return OTTGPU_OK;
}

int ottgpu_converter_wait(ottgpu_converter *cv)
{
    if (!cv) return fail(OTTGPU_ERR_ARG, "null converter");
    CK(cudaSetDevice(cv->gpu));
    CK(cudaStreamSynchronize(cv->stream));
    return OTTGPU_OK;
}

void ottgpu_converter_destroy(ottgpu_converter *cv) { converter_free(cv); }

}  // extern "C"

/* ---------------------------------------------------------------------------------------------------- thumbnail JPEG (N4) */
// transvideo.c:164-198 (save_frame_as_jpeg) on the GPU: see jpeg_core.cuh for what is and is not taken from libavcodec.

struct ottgpu_jpeg {
    int gpu = 0, w = 0, h = 0, qscale = 3;
    cudaStream_t stream = nullptr;
    ottgpu::JpegJob job{};
    uint8_t *d_planes = nullptr;                       // packed copy of host planes
    size_t plane_off[3] = {0, 0, 0};
    int plane_pitch[3] = {0, 0, 0}, plane_rows[3] = {0, 0, 0};
    ottgpu::JpegHuff *d_huff = nullptr;
    uint8_t *h_out = nullptr;                          // pinned: the file lands here first
    uint32_t *h_result = nullptr, *h_hist = nullptr;   // pinned
    ottgpu::JpegHuff *h_huff = nullptr;                // pinned
    uint8_t *h_header = nullptr;                       // pinned: SOI .. SOS of this picture
    std::vector<uint8_t> prefix, sof;                  // SOI, COM, DQT | SOF0: the same for every picture; the picture's DHT goes between them
};

namespace {

constexpr size_t kJpegMaxHeader = 2048;               // prefix (~110 bytes) + four DHT segments (<= 4 x 277) + SOS (14)

const unsigned char kMpeg1Intra[64] = {8,  16, 19, 22, 26, 27, 29, 34, 16, 16, 22, 24, 27, 29, 34, 37, 19, 22, 26, 27, 29, 34,
                                       34, 38, 22, 22, 26, 27, 29, 34, 37, 40, 22, 26, 27, 29, 32, 35, 40, 48, 26, 27, 29, 32,
                                       35, 40, 48, 58, 26, 27, 29, 34, 38, 46, 56, 69, 27, 29, 35, 38, 46, 56, 69, 83};
const unsigned char kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                                   41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                                   30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

void put16(std::vector<uint8_t> &v, int x) { v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x); }

// Optimal Huffman table of one symbol histogram, code lengths limited to 16 bits: the procedure of ITU-T T.81 Annex K.2 as
// the IJG library implements it (jpeg_gen_optimal_table): merge the two least frequent entries until one tree is left, with
// one reserved code point so that no code is all ones; then shorten codes longer than 16 bits.
// bits[l-1] = number of codes of length l, vals = the symbols in order of increasing code length.
void jpeg_optimal_table(const uint32_t *hist, int nsym, unsigned char bits[16], unsigned char *vals, int *nvals)
{
    long freq[257];
    int codesize[257], others[257];
    unsigned char nbits[33];
    for (int i = 0; i < 257; ++i) { freq[i] = i < nsym ? (long)hist[i] : 0; codesize[i] = 0; others[i] = -1; }
    std::memset(nbits, 0, sizeof(nbits));
    freq[256] = 1;
    for (;;) {
        int c1 = -1, c2 = -1;
        long v = 1000000000L;
        for (int i = 0; i <= 256; ++i)
            if (freq[i] && freq[i] <= v) { v = freq[i]; c1 = i; }
        v = 1000000000L;
        for (int i = 0; i <= 256; ++i)
            if (freq[i] && freq[i] <= v && i != c1) { v = freq[i]; c2 = i; }
        if (c2 < 0) break;
        freq[c1] += freq[c2];
        freq[c2] = 0;
        for (codesize[c1]++; others[c1] >= 0;) { c1 = others[c1]; codesize[c1]++; }
        others[c1] = c2;
        for (codesize[c2]++; others[c2] >= 0;) { c2 = others[c2]; codesize[c2]++; }
    }
    for (int i = 0; i <= 256; ++i)
        if (codesize[i]) nbits[std::min(codesize[i], 32)]++;
    for (int i = 32; i > 16; --i)
        while (nbits[i] > 0) {
            int j = i - 2;
            while (nbits[j] == 0) --j;
            nbits[i] -= 2;
            nbits[i - 1]++;
            nbits[j + 1] += 2;
            nbits[j]--;
        }
    int i = 16;
    while (i > 0 && nbits[i] == 0) --i;
    if (i > 0) nbits[i]--;                              // the reserved code point
    std::memcpy(bits, nbits + 1, 16);
    int p = 0;
    for (int l = 1; l <= 32; ++l)
        for (int sym = 0; sym < 256; ++sym)
            if (codesize[sym] == l) vals[p++] = (unsigned char)sym;
    *nvals = p;
}

// canonical Huffman codes from the bits / values lists (T.81 Annex C)
void jpeg_codes(const unsigned char *bits, const unsigned char *vals, uint16_t *code, uint8_t *len)
{
    int k = 0;
    uint32_t c = 0;
    for (int l = 1; l <= 16; ++l) {
        for (int i = 0; i < bits[l - 1]; ++i, ++k) {
            code[vals[k]] = (uint16_t)c++;
            len[vals[k]] = (uint8_t)l;
        }
        c <<= 1;
    }
}

void jpeg_free(ottgpu_jpeg *j)
{
    if (!j) return;
    cudaSetDevice(j->gpu);
    if (j->stream) { cudaStreamSynchronize(j->stream); cudaStreamDestroy(j->stream); }
    cudaFree(j->d_planes);
    cudaFree(j->d_huff);
    cudaFree(j->job.coef);
    cudaFree(j->job.bits);
    cudaFree(j->job.stream);
    cudaFree(j->job.out);
    cudaFree(j->job.result);
    cudaFree(j->job.ff_count);
    cudaFree(j->job.hist);
    cudaFreeHost(j->h_out);
    cudaFreeHost(j->h_result);
    cudaFreeHost(j->h_hist);
    cudaFreeHost(j->h_huff);
    cudaFreeHost(j->h_header);
    delete j;
}

int jpeg_build(ottgpu_jpeg *j)
{
    using namespace ottgpu;
    JpegJob &J = j->job;
    J.width = j->w; J.height = j->h;
    J.mcus_x = (j->w + 15) / 16; J.mcus_y = (j->h + 15) / 16;
    J.n_blocks = 6 * J.mcus_x * J.mcus_y;
    for (int i = 0; i < 64; ++i) J.q[i] = (uint16_t)std::min(std::max((kMpeg1Intra[i] * j->qscale) >> 3, 1), 255);
    J.q[0] = 8;
    for (int i = 0; i < 64; ++i) {                    // mpegvideo_enc.c convert_matrix, 16-bit (SIMD) form; intra bias 3/8 = 96 / 256
        const int den = 16 * J.q[i];
        int qm = (2 << 16) / den;
        if (qm == 0 || qm == 128 * 256) qm = 128 * 256 - 1;
        J.qmat16[i] = (uint16_t)qm;
        J.bias16[i] = (uint16_t)((96 * 256 + qm / 2) / qm);
    }
    // ---- the parts of the header every picture shares (segment order as libavcodec writes it: SOI COM DQT DHT SOF0 SOS)
    std::vector<uint8_t> &H = j->prefix;
    H = {0xFF, 0xD8};
    const char com[] = "libottgpu";
    H.push_back(0xFF); H.push_back(0xFE); put16(H, 2 + (int)sizeof(com) - 1);
    H.insert(H.end(), com, com + sizeof(com) - 1);
    H.push_back(0xFF); H.push_back(0xDB); put16(H, 67); H.push_back(0);
    for (int k = 0; k < 64; ++k) H.push_back((uint8_t)J.q[kZigzag[k]]);
    std::vector<uint8_t> &F = j->sof;
    F = {0xFF, 0xC0};
    put16(F, 17); F.push_back(8); put16(F, j->h); put16(F, j->w); F.push_back(3);
    const uint8_t comps[3][3] = {{1, 0x22, 0}, {2, 0x11, 0}, {3, 0x11, 0}};
    for (auto &c : comps) F.insert(F.end(), c, c + 3);
    // ---- buffers
    CK(cudaStreamCreateWithFlags(&j->stream, cudaStreamNonBlocking));
    CK(cudaMalloc((void **)&j->d_huff, sizeof(JpegHuff)));
    J.huff = j->d_huff;
    // worst case of a block: 64 coefficients x (16-bit code + 10 bits) + DC: < 1700 bits; sized at 1728 (a multiple of 32)
    const size_t words = (size_t)J.n_blocks * 54 + 8;
    if (words * 4 > (1u << 30)) return fail(OTTGPU_ERR_ARG, "picture too large for the JPEG encoder");
    J.stream_words = (uint32_t)words;
    J.out_capacity = (uint32_t)(kJpegMaxHeader + 2 * words * 4 + 2);     // every byte could be a stuffed 0xFF
    CK(cudaMalloc((void **)&J.coef, (size_t)J.n_blocks * 64 * sizeof(int16_t)));
    CK(cudaMalloc((void **)&J.bits, ((size_t)J.n_blocks + 1) * 4));
    CK(cudaMalloc((void **)&J.stream, words * 4));
    CK(cudaMalloc((void **)&J.out, J.out_capacity));
    CK(cudaMalloc((void **)&J.result, 8));
    CK(cudaMalloc((void **)&J.ff_count, ((size_t)jpeg_stuff_chunks(J.stream_words) + 1) * 4));
    CK(cudaMalloc((void **)&J.hist, 4 * 256 * sizeof(uint32_t)));
    CK(cudaMallocHost((void **)&j->h_out, J.out_capacity));
    CK(cudaMallocHost((void **)&j->h_result, 8));
    CK(cudaMallocHost((void **)&j->h_hist, 4 * 256 * sizeof(uint32_t)));
    CK(cudaMallocHost((void **)&j->h_huff, sizeof(JpegHuff)));
    CK(cudaMallocHost((void **)&j->h_header, kJpegMaxHeader));
    const int cw = (j->w + 1) / 2, chh = (j->h + 1) / 2;
    const int pw[3] = {j->w, cw, cw}, ph[3] = {j->h, chh, chh};
    size_t at = 0;
    for (int p = 0; p < 3; ++p) {
        j->plane_pitch[p] = (pw[p] + 15) & ~15;
        j->plane_rows[p] = ph[p];
        j->plane_off[p] = at;
        at += (size_t)j->plane_pitch[p] * ph[p];
    }
    CK(cudaMalloc((void **)&j->d_planes, at + 16));
    return OTTGPU_OK;
}

// this picture's tables from its symbol histogram: JpegHuff for the kernels, DHT x4 + SOS appended to the prefix
size_t jpeg_picture_header(ottgpu_jpeg *j)
{
    ottgpu::JpegHuff &hf = *j->h_huff;
    std::memset(&hf, 0, sizeof(hf));
    uint8_t *H = j->h_header;
    size_t n = j->prefix.size();
    std::memcpy(H, j->prefix.data(), n);
    // ONE DHT segment holding the four tables in libavcodec's order: DC luma, DC chroma, AC luma, AC chroma
    const int ids[4] = {0x00, 0x01, 0x10, 0x11};
    const size_t len_at = n + 2;
    H[n++] = 0xFF; H[n++] = 0xC4; n += 2;
    for (int t = 0; t < 4; ++t) {
        unsigned char bits[16], vals[256];
        int nv = 0;
        jpeg_optimal_table(j->h_hist + 256 * t, t < 2 ? 12 : 256, bits, vals, &nv);
        if (t < 2) jpeg_codes(bits, vals, hf.dc_code[t], hf.dc_len[t]);
        else jpeg_codes(bits, vals, hf.ac_code[t - 2], hf.ac_len[t - 2]);
        H[n++] = (uint8_t)ids[t];
        std::memcpy(H + n, bits, 16); n += 16;
        std::memcpy(H + n, vals, nv); n += nv;
    }
    const size_t seg = n - len_at;
    H[len_at] = (uint8_t)(seg >> 8); H[len_at + 1] = (uint8_t)seg;
    std::memcpy(H + n, j->sof.data(), j->sof.size()); n += j->sof.size();
    const uint8_t sos[14] = {0xFF, 0xDA, 0, 12, 3, 1, 0x00, 2, 0x11, 3, 0x11, 0, 63, 0};
    std::memcpy(H + n, sos, sizeof(sos));
    return n + sizeof(sos);
}

}  // namespace

extern "C" {

int ottgpu_jpeg_create(int gpu, int width, int height, int qscale, ottgpu_jpeg **out)
{
    if (!out) return fail(OTTGPU_ERR_ARG, "null argument");
    *out = nullptr;
    if (width < 1 || height < 1 || width > 16384 || height > 16384) return fail(OTTGPU_ERR_ARG, "picture size out of range");
    if (qscale < 1 || qscale > 31) return fail(OTTGPU_ERR_ARG, "qscale out of range (1..31, libavcodec's qmin..qmax)");
    int rc = prepare_device(gpu);
    if (rc) return rc;
    ottgpu_jpeg *j = new ottgpu_jpeg();
    j->gpu = gpu; j->w = width; j->h = height; j->qscale = qscale;
    rc = jpeg_build(j);
    if (rc) {
        const std::string why = t_err;
        jpeg_free(j);
        t_err = why;
        return rc;
    }
    *out = j;
    return OTTGPU_OK;
}

size_t ottgpu_jpeg_max_bytes(const ottgpu_jpeg *j) { return j ? j->job.out_capacity : 0; }

int ottgpu_jpeg_encode(ottgpu_jpeg *j, const ottgpu_picture *pic, void *dst, size_t dst_capacity, size_t *dst_bytes)
{
    if (!j || !pic || !dst || !dst_bytes) return fail(OTTGPU_ERR_ARG, "null argument");
    if (pic->mem != OTTGPU_MEM_HOST && pic->mem != OTTGPU_MEM_DEVICE) return fail(OTTGPU_ERR_ARG, "unknown memory kind");
    CK(cudaSetDevice(j->gpu));
    const int cw = (j->w + 1) / 2;
    const int row_bytes[3] = {j->w, cw, cw};
    ottgpu::JpegJob job = j->job;
    for (int p = 0; p < 3; ++p) {
        if (!pic->plane[p]) return fail(OTTGPU_ERR_ARG, "null plane");
        if (pic->stride[p] < row_bytes[p]) return fail(OTTGPU_ERR_ARG, "stride smaller than the row");
        if (pic->mem == OTTGPU_MEM_HOST) {
            uint8_t *d = j->d_planes + j->plane_off[p];
            CK(cudaMemcpy2DAsync(d, (size_t)j->plane_pitch[p], pic->plane[p], (size_t)pic->stride[p], (size_t)row_bytes[p], (size_t)j->plane_rows[p],
                                 cudaMemcpyHostToDevice, j->stream));
            job.plane[p] = d;
            job.pitch[p] = j->plane_pitch[p];
        } else {
            job.plane[p] = static_cast<const uint8_t *>(pic->plane[p]);
            job.pitch[p] = pic->stride[p];
        }
    }
    // phase 1: coefficients and the symbol histogram; the picture's optimal Huffman tables are built here from 4 KB of counts
    CK(ottgpu::launch_jpeg_analyse(job, j->stream));
    CK(cudaMemcpyAsync(j->h_hist, job.hist, 4 * 256 * sizeof(uint32_t), cudaMemcpyDeviceToHost, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    job.header_bytes = (int)jpeg_picture_header(j);
    CK(cudaMemcpyAsync(j->d_huff, j->h_huff, sizeof(ottgpu::JpegHuff), cudaMemcpyHostToDevice, j->stream));
    CK(cudaMemcpyAsync(job.out, j->h_header, (size_t)job.header_bytes, cudaMemcpyHostToDevice, j->stream));
    // phase 2: lengths, positions, codes, byte stuffing
    CK(ottgpu::launch_jpeg_encode(job, j->stream));
    // the size and, speculatively, the first 24 KB of the file in one round trip (a thumbnail is 3-15 KB)
    const size_t head = std::min<size_t>(job.out_capacity, 24 * 1024);
    CK(cudaMemcpyAsync(j->h_result, job.result, 8, cudaMemcpyDeviceToHost, j->stream));
    CK(cudaMemcpyAsync(j->h_out, job.out, head, cudaMemcpyDeviceToHost, j->stream));
    CK(cudaStreamSynchronize(j->stream));
    if (j->h_result[1]) return fail(OTTGPU_ERR_CUDA, "JPEG stream overflowed its buffer");
    const size_t n = j->h_result[0];
    *dst_bytes = n;
    if (n > dst_capacity) return fail(OTTGPU_ERR_NOMEM, "destination buffer too small for the JPEG file");
    if (n > head) {
        CK(cudaMemcpyAsync(j->h_out + head, job.out + head, n - head, cudaMemcpyDeviceToHost, j->stream));
        CK(cudaStreamSynchronize(j->stream));
    }
    std::memcpy(dst, j->h_out, n);
    return OTTGPU_OK;
}

void ottgpu_jpeg_destroy(ottgpu_jpeg *j) { jpeg_free(j); }

}  // extern "C"
