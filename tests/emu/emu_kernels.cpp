// TEST INFRASTRUCTURE ONLY: runs the CTA phases of ladder_core.cuh / yadif_core.cuh sequentially on the CPU
// (one "thread" at a time, barriers = loop boundaries).  See cuda_runtime.h in this directory.
#include <vector>
#include "kernels.h"
#include "ladder_core.cuh"
#include "yadif_core.cuh"

namespace ottgpu {

cudaError_t configure_kernels() { return cudaSuccess; }

cudaError_t launch_ladder(const Item *items, int n_items, const Binding *binds, size_t smem_bytes, cudaStream_t)
{
    if (smem_bytes > kMaxTileSmem) return cudaErrorInvalidValue;
    std::vector<unsigned char> smem(smem_bytes + 64);
    for (int blk = 0; blk < n_items; ++blk) {
        const Item it = items[blk];
        const Binding &b = binds[it.chan];
        TileGeom g;
        const int what = item_prepare(it, b, g);
        if (what == 0) continue;
        if (what == 2) {
            for (int tid = 0; tid < kThreads; ++tid) item_copy(tid, it, b);
            continue;
        }
        // every byte the phases touch must lie inside the launch's dynamic shared memory
        if ((size_t)g.off_vpos + (size_t)g.th * 4 > smem_bytes) return cudaErrorInvalidValue;
        std::fill(smem.begin(), smem.end(), 0xA5);       // poison: stale-data reads show up as mismatches
        unsigned char *sm = smem.data() + ((16 - ((uintptr_t)smem.data() & 15)) & 15);
        for (int tid = 0; tid < kThreads; ++tid) { phase_tables(tid, g, sm); phase_stage(tid, g, sm); }
        for (int tid = 0; tid < kThreads; ++tid) phase_hpass(tid, g, sm);
        for (int tid = 0; tid < kThreads; ++tid) phase_vpass(tid, g, sm);
    }
    return cudaSuccess;
}

cudaError_t launch_yadif(const YadifJob *jobs, int n_jobs, int max_w, int max_h, cudaStream_t)
{
    for (int z = 0; z < n_jobs; ++z) {
        if (!jobs[z].active) continue;
        for (int y = 0; y < (max_h + 7) / 8 * 8; ++y)
            for (int x = 0; x < (max_w + 31) / 32 * 32; ++x) yadif_position(jobs[z], x, y);
    }
    return cudaSuccess;
}

}  // namespace ottgpu
