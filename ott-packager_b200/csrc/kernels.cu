// sm_100a kernels of libottgpu.
//   ladder_kernel : persistent fused H+V polyphase scaler / format writer.  Each CTA pulls output tiles (work items:
//                   all rungs of all channels of a batch) from a global counter; the tile's source rows are fetched by
//                   the TMA engine (cp.async.bulk global->shared, mbarrier completion) into one of two shared buffers
//                   while the previous tile is still computing.  The source frame is read from HBM once and shared by
//                   the rungs through L2 (items are ordered by source row band), see DESIGN.md.
//   yadif_kernel  : yadif mode 0 first-field deinterlacer on planar 8/16-bit frames.
// The per-thread work lives in ladder_core.cuh / yadif_core.cuh.
#include "kernels.h"
#include "ladder_core.cuh"
#include "yadif_core.cuh"

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
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
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

// Fetch the tile's source rows into its shared buffer: TMA row copies when the plane allows it, else through registers.
__device__ __forceinline__ void stage_issue(int tid, const TileGeom &g, unsigned char *smem, uint64_t *bar)
{
    if (!g.bulk_ok) {
        phase_stage(tid, g, smem);
        return;
    }
    if (tid < 32) {
        const int rows = g.nplanes * g.sr;
        fence_proxy_async();           // the buffer was last read through the generic proxy (H pass of an earlier tile)
        if (tid == 0) mbar_expect_tx(bar, (uint32_t)(rows * g.row_copy_bytes));
        __syncwarp();
        for (int r = tid; r < rows; r += 32) {
            const int p = r >= g.sr, rr = r - p * g.sr;
            const uint8_t *src = (p ? g.src1 : g.src0) + (size_t)(g.sy0 + rr) * (p ? g.src_pitch1 : g.src_pitch0) + g.sx0;
            bulk_g2s(smem + g.off_src + r * g.spw * 4, src, (uint32_t)g.row_copy_bytes, bar);
        }
    }
}

__global__ void __launch_bounds__(kThreads, 3) ladder_kernel(const LadderLaunch L)
{
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t mbar[2];
    // Tile descriptors live in shared memory (three rotating slots: previous / current / next) so that no geometry
    // has to stay in registers across the phases; thread 0 resolves the NEXT tile while the others finish the last one.
    __shared__ TileGeom s_geom[3];
    __shared__ int s_idx[3], s_what[3];
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_mbar_init();
        const int first = atomicAdd(L.counter, 1);
        s_idx[0] = first;
        s_what[0] = first < L.n_items ? item_prepare(L.items[first], L.binds[L.items[first].chan], L, 0, s_geom[0]) : 0;
    }
    __syncthreads();
    const int db = L.double_buffer;
    int buf = 0;
    uint32_t parity0 = 0, parity1 = 0;
    if (db && s_what[0] == 1) stage_issue(tid, s_geom[0], smem, &mbar[0]);
    for (int slot = 0;; slot = slot == 2 ? 0 : slot + 1) {
        const int cur = s_idx[slot];
        if (cur >= L.n_items) break;
        const int nslot = slot == 2 ? 0 : slot + 1;
        if (tid == 0) {
            const int nxt = atomicAdd(L.counter, 1);
            s_idx[nslot] = nxt;
            s_what[nslot] = nxt < L.n_items ? item_prepare(L.items[nxt], L.binds[L.items[nxt].chan], L, db ? buf ^ 1 : 0, s_geom[nslot]) : 0;
        }
        __syncthreads();   // (A) next tile resolved; every thread is past the previous tile's V pass and H pass
        if (db && s_what[nslot] == 1) stage_issue(tid, s_geom[nslot], smem, &mbar[buf ^ 1]);   // prefetch
        const int what = s_what[slot];
        const TileGeom &g = s_geom[slot];
        if (what == 2) {
            item_copy(tid, L.items[cur], L.binds[L.items[cur].chan]);
        } else if (what == 1) {
            if (!db) {                                     // single buffer (very large tiles): fetch the current tile now
                stage_issue(tid, g, smem, &mbar[0]);
                if (!g.bulk_ok) __syncthreads();
            }
            if (g.bulk_ok) {
                if (buf == 0) { mbar_wait(&mbar[0], parity0); parity0 ^= 1; }
                else { mbar_wait(&mbar[1], parity1); parity1 ^= 1; }
            }
            phase_tables(tid, g, smem);
            phase_hpass(tid, g, smem);
            __syncthreads();   // (B) H-pass result and V tables complete
            phase_vpass(tid, g, smem);
        }
        buf ^= db;
    }
}

__global__ void __launch_bounds__(256) yadif_kernel(const YadifJob *__restrict__ jobs)
{
    const YadifJob &j = jobs[blockIdx.z];
    if (!j.active) return;
    yadif_position(j, blockIdx.x * 32 + (threadIdx.x & 31), blockIdx.y * 8 + (threadIdx.x >> 5));
}

cudaError_t configure_kernels()
{
    return cudaFuncSetAttribute(ladder_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileSmem);
}

size_t ladder_smem_bytes(int src_max, int mid_max, int vc_max, int *double_buffer)
{
    const size_t two = 2 * (size_t)src_max + mid_max + vc_max + 16, one = (size_t)src_max + mid_max + vc_max + 16;
    const int db = two <= kMaxTileSmem;
    if (double_buffer) *double_buffer = db;
    return db ? two : one;
}

cudaError_t launch_ladder(LadderLaunch L, cudaStream_t s)
{
    if (L.n_items <= 0) return cudaSuccess;
    int db = 0;
    const size_t smem_bytes = ladder_smem_bytes(L.src_max, L.mid_max, L.vc_max, &db);
    if (smem_bytes > kMaxTileSmem) return cudaErrorInvalidValue;
    L.double_buffer = db;
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ladder_kernel, kThreads, smem_bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidValue;
    const int grid = L.n_items < sms * per_sm ? L.n_items : sms * per_sm;     // one wave of resident CTAs
    ladder_kernel<<<grid, kThreads, smem_bytes, s>>>(L);
    return cudaGetLastError();
}

cudaError_t launch_yadif(const YadifJob *jobs, int n_jobs, int max_w, int max_h, cudaStream_t s)
{
    if (n_jobs <= 0) return cudaSuccess;
    dim3 grid((max_w + 31) / 32, (max_h + 7) / 8, n_jobs);
    yadif_kernel<<<grid, 256, 0, s>>>(jobs);
    return cudaGetLastError();
}

}  // namespace ottgpu
