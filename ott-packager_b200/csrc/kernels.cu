// sm_100a kernels of libottgpu.
//   ladder_kernel : persistent fused H+V polyphase scaler / format writer.  Each CTA pulls output tiles (work items:
//                   all rungs of all channels of a batch) from a global counter; the tile's source rows are fetched by
//                   the TMA engine (one 2-D tensor-map box per plane, mbarrier completion) into the shared source buffer
//                   while the previous tile's V pass is still running.  The source frame is read from HBM once and shared by
//                   the rungs through L2 (items are ordered by source row band), see DESIGN.md.
//   yadif_kernel  : yadif mode 0 first-field deinterlacer on planar 8/16-bit frames.
//   convert_kernel: decode-side 4:2:2 / 10-bit -> packed 8-bit 4:2:0 converter.
// The per-thread work lives in ladder_core.cuh / yadif_core.cuh.
#include "kernels.h"
#include <cstdlib>
#include <cstdio>
#include "ladder_core.cuh"
#include "yadif_core.cuh"
#include "convert_core.cuh"
#include "jpeg_core.cuh"

namespace ottgpu {

// ---------------------------------------------------------------------------------------------- mbarrier / TMA (PTX)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// one box (tensor map coordinates: x in 32-bit elements, y in rows, z = buffer of the arena) -> shared memory, completion on `bar`
__device__ __forceinline__ void tma_box_3d(void *dst_smem, const void *tmap, int x, int y, int z, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// producer-side wait: try_wait suspends the warp in hardware up to the time hint, so a waiting producer takes no issue
// slots from the compute warps and wakes as soon as the phase completes (a nanosleep poll added up to 0.25 us per wait)
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity)
{
    for (;;) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(4000u)
            : "memory");
        if (ok) return;
    }
}
// compute-warp wait of the strip mode: usually the phase is complete already (one test, no loop); a warp that is ahead of
// the others sleeps between polls instead of spinning on the barrier
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity)
{
    for (uint32_t ns = 32;; ns = ns < 256 ? 2 * ns : 256) {
        uint32_t ok;
        asm volatile(
            "{\n"
            ".reg .pred P1;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n"
            "selp.u32 %0, 1, 0, P1;\n"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return;
        __nanosleep(ns);
    }
}
// adds to the pending transaction bytes of the current phase without arriving
__device__ __forceinline__ void mbar_add_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk copy global -> shared (16-byte aligned addresses and size), completion on `bar`
__device__ __forceinline__ void bulk_copy_g2s(void *dst_smem, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void compute_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory"); }

// Experiment build (scripts/build_variant.sh clocks -DOTTGPU_PHASE_CLOCKS, scripts/gpu_clocks.sh): lane 0 of every warp of two
// CTAs adds up where its cycles go -- waiting for a tile's source window, H pass, the CTA barrier, V pass, copy items; the
// producer: waiting for a descriptor slot, fetching the item, resolving it, table copies, waiting for the source buffer,
// issuing the boxes -- and prints the totals of a launch.  Never part of the product build.
#if defined(OTTGPU_PHASE_CLOCKS)
__device__ int g_phase_prints = 0;
#define PC_MARK(acc) { const long long t_ = clock64(); acc += t_ - pc_t0; pc_t0 = t_; }
#define PP_MARK(acc) { const long long t_ = clock64(); acc += t_ - pp_t0; pp_t0 = t_; }
#else
#define PC_MARK(acc)
#define PP_MARK(acc)
#endif
constexpr int kLadderThreads = kThreads + 32;     // 12 compute warps + 1 producer warp

// Warp-specialised persistent kernel.
//   producer warp : pulls work items from the global counter and resolves their geometry into a ring of kGeomSlots
//                   shared descriptors, running ahead of the compute warps; it copies the tile's V tables into the
//                   descriptor's table region, waits for a free source buffer and has the TMA engine fetch the tile's
//                   source window (one 2-D tensor-map box per staged plane); completion (transaction bytes) is
//                   signalled on the descriptor's `ready` mbarrier.
//   compute warps : wait ready[slot] -> H pass (staged rows -> int16 intermediates) -> V pass + store -> release the
//                   descriptor slot and the source buffer.
// One source buffer: it is released after the H pass, tile k+1 streams in while tile k's V pass runs.
template <int BPS>
__global__ void __launch_bounds__(kLadderThreads, 2) ladder_kernel(const LadderLaunch L)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int kGeomSlots = BPS == 2 ? kGeomSlots16 : kGeomSlots8;
    __shared__ __align__(8) uint64_t ready[kGeomSlots], gfree[kGeomSlots], empty[2];
    __shared__ __align__(16) TileGeom s_geom[kGeomSlots];
    __shared__ int s_idx[kGeomSlots], s_what[kGeomSlots], s_buf[kGeomSlots];
    const int tid = threadIdx.x;
    const int nbuf = L.double_buffer ? 2 : 1;
    if (tid == 0) {
        for (int i = 0; i < kGeomSlots; ++i) { mbar_init(&ready[i], 1); mbar_init(&gfree[i], kThreads); }
        for (int i = 0; i < 2; ++i) mbar_init(&empty[i], kThreads);
        fence_mbar_init();
    }
    __syncthreads();

    if (tid >= kThreads) {
        // ------------------------------------------------------------------ producer warp
        const int lane = tid - kThreads;
        uint32_t pg = (1u << kGeomSlots) - 1, pe = 3;    // parity bits; a fresh mbarrier passes a wait on parity 1
        int buf = 0, mid = 0;
        // Work items are claimed kClaim at a time: one atomic on the launch-wide counter per tile (296 CTAs on one address)
        // bounded the whole launch at ~5 ns per tile.  The next claim is issued while the current one is being worked
        // off, so its latency stays hidden.  Consecutive items share source rows, so a CTA's run of tiles also helps L2.
        const int kClaim = L.claim;
        int cur = 0, cur_end = 0, nxt = 0;                   // lane 0: current run [cur, cur_end), start of the prefetched run
        if (lane == 0) {
            cur = atomicAdd(L.counter, kClaim);
            cur_end = cur + kClaim;
            nxt = atomicAdd(L.counter, kClaim);
        }
#if defined(OTTGPU_PHASE_CLOCKS)
        long long pp_gfree = 0, pp_item = 0, pp_prep = 0, pp_tab = 0, pp_empty = 0, pp_tma = 0, pp_t0 = clock64();
#endif
        for (int slot = 0;; slot = (slot + 1) & (kGeomSlots - 1)) {
            int idx = 0;
            if (lane == 0) {
                if (cur == cur_end) {
                    cur = nxt;
                    cur_end = cur + kClaim;
                    if (cur < L.n_items) nxt = atomicAdd(L.counter, kClaim);
                }
                idx = cur++;
            }
            idx = __shfl_sync(0xffffffffu, idx, 0);
            mbar_wait_backoff(&gfree[slot], (pg >> slot) & 1);
            pg ^= 1u << slot;
            PP_MARK(pp_gfree)
            if (idx >= L.n_items) {
                if (lane == 0) {
                    s_idx[slot] = idx;
                    mbar_arrive(&ready[slot]);
                }
#if defined(OTTGPU_PHASE_CLOCKS)
                if (lane == 0 && (blockIdx.x == 0 || blockIdx.x == 150) && L.n_items > 5000 && atomicAdd(&g_phase_prints, 1) < 104)
                    printf("cta %d PRODUCER: gfree %lld  item %lld  prepare %lld  tables %lld  empty %lld  tma %lld (cycles)\n", (int)blockIdx.x,
                           pp_gfree, pp_item, pp_prep, pp_tab, pp_empty, pp_tma);
#endif
                break;
            }
            // the item's static block: eight lanes, one 16-byte vector each, straight into the descriptor slot
            TileGeom &g = s_geom[slot];
            if (lane < (int)(sizeof(Item) / 16))
                reinterpret_cast<uint4 *>(&g)[lane] = __ldg(reinterpret_cast<const uint4 *>(L.items + idx) + lane);
            __syncwarp();
            PP_MARK(pp_item)
            int what = 0;
            if (lane == 0) {
                s_idx[slot] = idx;
                s_buf[slot] = buf;
                what = item_prepare(L.binds[g.chan], L, buf, mid, slot, g);     // frame / surface pointers of this step
                s_what[slot] = what;
            }
            what = __shfl_sync(0xffffffffu, what, 0);
            __syncwarp();
            PP_MARK(pp_prep)
            if (what == 1) {
                if (g.vc_bytes) {                              // V coefficients / row positions: two bulk copies
                    if (lane == 0) {
                        fence_proxy_async();                   // the table region was last read through the generic proxy
                        mbar_add_tx(&ready[slot], (uint32_t)g.vc_bytes + g.vpos_bytes);
                        bulk_copy_g2s(smem + g.off_vc, g.vc_src, g.vc_bytes, &ready[slot]);
                        bulk_copy_g2s(smem + g.off_vpos, g.vpos_src, g.vpos_bytes, &ready[slot]);
                    }
                } else {
                    phase_tables(lane, 32, g, smem);
                    __syncwarp();
                }
                PP_MARK(pp_tab)
                mbar_wait_backoff(&empty[buf], (pe >> buf) & 1);   // the tile that last used this source buffer is done
                pe ^= 1u << buf;
                PP_MARK(pp_empty)
#if defined(OTTGPU_DEBUG_NO_TMA)
                if (false) {
                    if (lane == 0) {
#else
#if defined(OTTGPU_EXPERIMENT_NO_WINDOW_WAIT)   /* timing experiment only (stale windows, no fetch at all): upper bound of a fully hidden source fetch */
                if (false) {
                    if (lane == 0) {
#else
                if (g.bulk_ok) {
                    if (lane == 0) {
#endif
#endif
                        fence_proxy_async();                   // the buffer was last read through the generic proxy
                        mbar_expect_tx(&ready[slot], (uint32_t)(g.staged_planes * g.box_bytes));
                        unsigned char *dst = smem + g.off_src;
                        tma_box_3d(dst, g.tmap0, g.tma_x, g.sy0, g.src_z, &ready[slot]);
                        if (g.staged_planes == 2) tma_box_3d(dst + g.plane_stride, g.tmap1, g.tma_x, g.sy0, g.src_z, &ready[slot]);
                    }
                } else if (lane == 0) {
                    mbar_arrive(&ready[slot]);                 // compute warps stage this tile through registers
                }
                buf = (buf + 1 == nbuf) ? 0 : buf + 1;
                mid = (mid + 1 == L.mid_buffers) ? 0 : mid + 1;
                PP_MARK(pp_tma)
            } else if (lane == 0) {
                mbar_arrive(&ready[slot]);                     // copy item / skipped item: no source buffer involved
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- compute warps
    constexpr bool kStrip = BPS == 1 && OTTGPU_STRIP;
    uint32_t pr = 0;
#if defined(OTTGPU_PHASE_CLOCKS)
    long long pc_ready = 0, pc_h = 0, pc_bar = 0, pc_v = 0, pc_copy = 0, pc_t0 = clock64(), pc_tiles = 0;
#endif
    for (int slot = 0;; slot = (slot + 1) & (kGeomSlots - 1)) {
        // strip mode: the warps of a CTA run apart, and one that is ahead must not spin on the barrier (a spinning warp takes
        // issue slots from the working ones): try_wait with a time hint suspends it in hardware until the phase completes
        if (kStrip) mbar_wait_sleep(&ready[slot], (pr >> slot) & 1);
        else mbar_wait(&ready[slot], (pr >> slot) & 1);
        pr ^= 1u << slot;
        PC_MARK(pc_ready)
        const int idx = s_idx[slot];
        if (idx >= L.n_items) break;
        const int what = s_what[slot];
        const TileGeom &g = s_geom[slot];
        if (what == 2) {
#if !defined(OTTGPU_EXPERIMENT_NO_COPY)     /* timing experiment only: upper bound of moving the copy items out of this kernel */
            item_copy(tid, g, L.binds[g.chan]);
#endif
            PC_MARK(pc_copy)
        } else if (what == 1) {
#if defined(OTTGPU_DEBUG_NO_TMA)
            if (false) {
#else
            if (!g.bulk_ok) {                      // unaligned planes: staged through registers by the compute warps
#endif
                phase_stage(tid, g, smem);
                compute_sync();
            }
            if (BPS == 2 && g.src_shift && g.hsplit != 16) {     // P010 on the dp2a fallback: right-align the 10 significant bits (the MMA path reads the raw bytes)
                phase_fixup16(tid, g, smem);
                compute_sync();
            }
#if !defined(OTTGPU_DEBUG_SKIP_H)
            phase_hpass<BPS>(tid, g, smem);
#endif
            PC_MARK(pc_h)
            // strip mode: the V pass of a warp reads only what the same warp's H pass wrote (and the V tables, complete
            // since ready[slot]); otherwise the H-pass result of the whole CTA must be complete
#if defined(OTTGPU_EXPERIMENT_NO_BARRIER)   /* timing experiment only (races): upper bound of a perfectly balanced H -> V hand-over */
            __syncwarp();
#else
            if (kStrip) __syncwarp();
            else compute_sync();
#endif
            PC_MARK(pc_bar)
            mbar_arrive(&empty[s_buf[slot]]);      // this thread no longer reads the source buffer
#if !defined(OTTGPU_DEBUG_SKIP_V)
            phase_vpass<BPS>(tid, g, smem);
#endif
            PC_MARK(pc_v)
#if defined(OTTGPU_PHASE_CLOCKS)
            ++pc_tiles;
#endif
            // one intermediate region: everybody must be done reading it before the next tile's H pass writes it.  Two
            // regions: the next H pass writes the other one, and the region it will overwrite after that was last read
            // before the barrier above - no wait here, a warp that is done starts the next tile at once
            if (!kStrip && L.mid_buffers == 1) compute_sync();
        }
        mbar_arrive(&gfree[slot]);                 // descriptor slot may be rewritten
    }
#if defined(OTTGPU_PHASE_CLOCKS)
    if ((tid & 31) == 0 && (blockIdx.x == 0 || blockIdx.x == 150) && L.n_items > 5000 && atomicAdd(&g_phase_prints, 1) < 104)
        printf("cta %d warp %d tiles %lld: ready %lld  H %lld  barrier %lld  V %lld  copy %lld  (cycles)\n", (int)blockIdx.x, tid >> 5, pc_tiles,
               pc_ready, pc_h, pc_bar, pc_v, pc_copy);
#endif
}

#if !defined(OTTGPU_YADIF_BLOCKS)
#define OTTGPU_YADIF_BLOCKS 5      /* resident blocks per SM the register budget is set for (48 registers, no spills); measured on
                                     the B200, ms per 64 x 1080i: 3 blocks / 69 registers 0.389, 4 / 64 0.349, 5 / 48 0.337, 6 / 40 (spills) 0.343 */
#endif
__global__ void __launch_bounds__(256, OTTGPU_YADIF_BLOCKS) yadif_kernel(const YadifJob *__restrict__ jobs)
{
    const YadifJob &j = jobs[blockIdx.z];
    if (!j.active) return;
    const int gx = blockIdx.x * 32 + (threadIdx.x & 31), gy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (j.fast) yadif_position_fast(j, gx, gy);          // gx counts 8-sample groups, gy row pairs (luma then chroma)
    else yadif_position(j, gx, gy);
}

// decode-side converter: the job (three plane descriptors + the Bayer table) travels as a kernel parameter
__global__ void __launch_bounds__(256) convert_kernel(const ConvertJob j)
{
    convert_position(j, blockIdx.x * 32 + (threadIdx.x & 31), blockIdx.y * 8 + (threadIdx.x >> 5));
}

// many pictures (different converters) in one launch: blockIdx.z = job
__global__ void __launch_bounds__(256) convert_batch_kernel(const ConvertJob *__restrict__ jobs)
{
    convert_position(jobs[blockIdx.z], blockIdx.x * 32 + (threadIdx.x & 31), blockIdx.y * 8 + (threadIdx.x >> 5));
}

// ---------------------------------------------------------------------------------------------- thumbnail JPEG (N4)
__global__ void __launch_bounds__(128) jpeg_dct_kernel(const JpegJob j) { jpeg_block_dct(j, blockIdx.x * 128 + threadIdx.x); }
__global__ void __launch_bounds__(128) jpeg_hist_kernel(const JpegJob j) { jpeg_block_hist(j, blockIdx.x * 128 + threadIdx.x); }
__global__ void __launch_bounds__(128) jpeg_bits_kernel(const JpegJob j) { jpeg_block_bits(j, blockIdx.x * 128 + threadIdx.x); }
__global__ void __launch_bounds__(128) jpeg_emit_kernel(const JpegJob j) { jpeg_block_emit(j, blockIdx.x * 128 + threadIdx.x); }
__global__ void __launch_bounds__(128) jpeg_stuff_count_kernel(const JpegJob j) { jpeg_stuff_count(j, blockIdx.x * 128 + threadIdx.x); }
__global__ void __launch_bounds__(128) jpeg_stuff_write_kernel(const JpegJob j, int chunks)
{
    const int c = blockIdx.x * 128 + threadIdx.x;
    jpeg_stuff_write(j, c);
    if (c == 0) jpeg_finish(j, chunks);
}
// in-place exclusive scan of v[0..n), total to v[n]; one CTA (the arrays are a few hundred to a few ten thousand entries)
__global__ void __launch_bounds__(1024) jpeg_scan_kernel(uint32_t *v, int n)
{
    __shared__ uint32_t part[1024];
    const int t = threadIdx.x, per = (n + 1023) / 1024, i0 = t * per, i1 = min(i0 + per, n);
    uint32_t sum = 0;
    for (int i = i0; i < i1; ++i) sum += v[i];
    part[t] = sum;
    __syncthreads();
    for (int d = 1; d < 1024; d <<= 1) {
        const uint32_t add = t >= d ? part[t - d] : 0u;
        __syncthreads();
        part[t] += add;
        __syncthreads();
    }
    uint32_t run = part[t] - sum;                               // exclusive prefix of this thread's segment
    for (int i = i0; i < i1; ++i) {
        const uint32_t x = v[i];
        v[i] = run;
        run += x;
    }
    if (t == 1023) v[n] = part[1023];
}

int jpeg_stuff_chunks(uint32_t stream_words) { return (int)(((size_t)stream_words * 4 + kJpegChunk - 1) / kJpegChunk); }

// phase 1: coefficients + symbol histogram (the host then builds the picture's Huffman tables)
cudaError_t launch_jpeg_analyse(const JpegJob &j, cudaStream_t s)
{
    if (j.n_blocks <= 0) return cudaSuccess;
    const int gb = (j.n_blocks + 127) / 128;
    cudaError_t e = cudaMemsetAsync(j.hist, 0, 4 * 256 * sizeof(uint32_t), s);
    if (e != cudaSuccess) return e;
    jpeg_dct_kernel<<<gb, 128, 0, s>>>(j);
    jpeg_hist_kernel<<<gb, 128, 0, s>>>(j);
    return cudaGetLastError();
}

// phase 2: lengths, positions, codes, byte stuffing with the tables in j.huff
cudaError_t launch_jpeg_encode(const JpegJob &j, cudaStream_t s)
{
    if (j.n_blocks <= 0) return cudaSuccess;
    const int gb = (j.n_blocks + 127) / 128;
    cudaError_t e = cudaMemsetAsync(j.stream, 0, (size_t)j.stream_words * 4, s);
    if (e != cudaSuccess) return e;
    jpeg_bits_kernel<<<gb, 128, 0, s>>>(j);
    jpeg_scan_kernel<<<1, 1024, 0, s>>>(j.bits, j.n_blocks);
    jpeg_emit_kernel<<<gb, 128, 0, s>>>(j);
    // the stuffing pass is sized for the largest stream the buffers can hold; chunks beyond the real length exit at once
    const int chunks = jpeg_stuff_chunks(j.stream_words), gc = (chunks + 127) / 128;
    e = cudaMemsetAsync(j.ff_count, 0, (size_t)(chunks + 1) * 4, s);
    if (e != cudaSuccess) return e;
    jpeg_stuff_count_kernel<<<gc, 128, 0, s>>>(j);
    jpeg_scan_kernel<<<1, 1024, 0, s>>>(j.ff_count, chunks);
    jpeg_stuff_write_kernel<<<gc, 128, 0, s>>>(j, chunks);
    return cudaGetLastError();
}

cudaError_t configure_kernels()
{
    cudaError_t e = cudaFuncSetAttribute(ladder_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(ladder_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileSmem);
    return e;
}

// Shared memory layout of one CTA: [source buffer(s)] [intermediate region(s)] [kGeomSlots table regions].
// Preferred: one source buffer + two intermediate regions (the next tile's window is fetched during the V pass, which
// is far longer than the copy, and the V -> next-H hand-over needs no barrier).  Launches whose tiles are too big for
// that (very long filters) run with one of each.
size_t ladder_smem_bytes(int bps, int src_max, int mid_max, int vc_max, int *double_buffer, int *mid_buffers)
{
    const size_t tables = (size_t)geom_slots(bps) * vc_max;     // one table region per descriptor slot
    static const int want_two_src = [] { const char *v = std::getenv("OTTGPU_DOUBLE_SRC"); return v ? std::atoi(v) : 0; }();
    const size_t nsrc = want_two_src && 2 * (size_t)src_max + 2 * (size_t)mid_max + tables + 16 <= kMaxTileSmem ? 2 : 1;
    const size_t two_mid = nsrc * src_max + 2 * (size_t)mid_max + tables + 16, one = (size_t)src_max + mid_max + tables + 16;
    // strip mode (8-bit): one region -- a warp only touches its own columns of it, in program order
    const int mb = (two_mid <= kMaxTileSmem && !(OTTGPU_STRIP && bps == 1)) ? 2 : 1;
    if (double_buffer) *double_buffer = mb == 2 && nsrc == 2;
    if (mid_buffers) *mid_buffers = mb;
    return mb == 2 ? two_mid : one;
}

int ladder_ctas_per_sm(size_t smem_bytes)
{
    int per_sm = 0;
    if (smem_bytes > kMaxTileSmem) return 0;
    // both instances are built with the same launch bounds; the 8-bit one stands for the pair
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ladder_kernel<1>, kLadderThreads, smem_bytes) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return per_sm;
}

cudaError_t launch_ladder(LadderLaunch L, cudaStream_t s)
{
    if (L.n_items <= 0) return cudaSuccess;
    int db = 0, mb = 1;
    const size_t smem_bytes = ladder_smem_bytes(L.bps, L.src_max, L.mid_max, L.vc_max, &db, &mb);
    if (smem_bytes > kMaxTileSmem) return cudaErrorInvalidValue;
    L.double_buffer = db;
    L.mid_buffers = mb;
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess)
        e = L.bps == 2 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ladder_kernel<2>, kLadderThreads, smem_bytes)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ladder_kernel<1>, kLadderThreads, smem_bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidValue;
    const int grid = L.n_items < sms * per_sm ? L.n_items : sms * per_sm;     // one wave of resident CTAs
    // two items per claim on big launches, one on small ones (they keep every CTA busy).  Measured on the final build, ms
    // per 64 frames on cfg4 / cfg2: 1 -> 0.427 / 0.540, 2 -> 0.426 / 0.536, 4 -> 0.431 / 0.539, 8 -> 0.441, 16 -> 0.459 (longer
    // runs of one CTA lose the fine interleaving of heavy and light tiles); OTT_CLAIM=n overrides it for experiments
    L.claim = L.n_items / (grid * 4) >= 2 ? 2 : 1;
    {
        static const int env_claim = [] { const char *e = std::getenv("OTT_CLAIM"); return e ? std::atoi(e) : 0; }();
        if (env_claim > 0 && L.claim == 2) L.claim = env_claim;
    }
    if (L.bps == 2) ladder_kernel<2><<<grid, kLadderThreads, smem_bytes, s>>>(L);
    else ladder_kernel<1><<<grid, kLadderThreads, smem_bytes, s>>>(L);
    return cudaGetLastError();
}

cudaError_t launch_yadif(const YadifJob *jobs, int n_jobs, int max_w, int max_h, bool all_fast, cudaStream_t s)
{
    if (n_jobs <= 0) return cudaSuccess;
    // fast jobs need (w/8) x (row pairs of luma and chroma) positions, slow ones w x h: launch the union (extra CTAs
    // exit at once)
    const int cols = all_fast ? (max_w / 8 + 31) / 32 : (max_w + 31) / 32;
    const int rows = all_fast ? (max_h + 1) / 2 + ((max_h + 1) / 2 + 1) / 2 : max_h + (max_h + 1) / 2;
    dim3 grid(cols, (rows + 7) / 8, n_jobs);
    yadif_kernel<<<grid, 256, 0, s>>>(jobs);
    return cudaGetLastError();
}

cudaError_t launch_convert(const ConvertJob &job, cudaStream_t s)
{
    const int rows = job.plane[0].h + job.plane[1].h + job.plane[2].h;
    const int groups = (job.plane[0].w + kConvertGroup - 1) / kConvertGroup;
    if (rows <= 0 || groups <= 0) return cudaSuccess;
    dim3 grid((groups + 31) / 32, (rows + 7) / 8, 1);
    convert_kernel<<<grid, 256, 0, s>>>(job);
    return cudaGetLastError();
}

cudaError_t launch_convert_batch(const ConvertJob *jobs, int n_jobs, int max_w, int max_rows, cudaStream_t s)
{
    if (n_jobs <= 0 || max_w <= 0 || max_rows <= 0) return cudaSuccess;
    // luma: 16-sample groups over max_w; chroma planes are half as wide, their 8-sample groups fit the same range
    dim3 grid(((max_w + kConvertGroup - 1) / kConvertGroup + 31) / 32, (max_rows + 7) / 8, n_jobs);
    convert_batch_kernel<<<grid, 256, 0, s>>>(jobs);
    return cudaGetLastError();
}

}  // namespace ottgpu
