// sm_100a kernels of libottgpu.
//   ladder_kernel : fused H+V polyphase scaler / format writer, one CTA per output tile (work item), all rungs of all
//                   channels of a batch in ONE launch; the source frame is read from HBM once and shared by the rungs
//                   through L2 (items are ordered by source row band), see DESIGN.md.
//   yadif_kernel  : yadif mode 0 first-field deinterlacer on planar 8/16-bit frames.
// The per-thread work lives in ladder_core.cuh / yadif_core.cuh.
#include "kernels.h"
#include "ladder_core.cuh"
#include "yadif_core.cuh"

namespace ottgpu {

__global__ void __launch_bounds__(kThreads, 3) ladder_kernel(const Item *__restrict__ items, const Binding *__restrict__ binds)
{
    extern __shared__ __align__(16) unsigned char smem[];
    const Item it = items[blockIdx.x];
    const Binding &b = binds[it.chan];
    const int tid = threadIdx.x;
    TileGeom g;
    const int what = item_prepare(it, b, g);
    if (what == 0) return;
    if (what == 2) { item_copy(tid, it, b); return; }
    phase_tables(tid, g, smem);
    phase_stage(tid, g, smem);
    __syncthreads();
    phase_hpass(tid, g, smem);
    __syncthreads();
    phase_vpass(tid, g, smem);
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

cudaError_t launch_ladder(const Item *items, int n_items, const Binding *binds, size_t smem_bytes, cudaStream_t s)
{
    if (n_items <= 0) return cudaSuccess;
    if (smem_bytes > kMaxTileSmem) return cudaErrorInvalidValue;
    ladder_kernel<<<n_items, kThreads, smem_bytes, s>>>(items, binds);
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
