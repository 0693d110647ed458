// TEST INFRASTRUCTURE ONLY: runs the CTA phases of ladder_core.cuh / yadif_core.cuh sequentially on the CPU
// (one "thread" at a time, barriers = loop boundaries).  See cuda_runtime.h in this directory.
#include <vector>
#include "kernels.h"
#include "ladder_core.cuh"
#include "yadif_core.cuh"
#include "convert_core.cuh"
#include "jpeg_core.cuh"

// test hook: how many fast V rows (thread x row x plane) ran as FMAs / on the integer bias path since the last call
extern "C" void ottgpu_emu_vrows(long long *out)
{
    out[0] = ottgpu::ott_emu_vrows[0]; out[1] = ottgpu::ott_emu_vrows[1];
    ottgpu::ott_emu_vrows[0] = ottgpu::ott_emu_vrows[1] = 0;
}

namespace ottgpu {

cudaError_t configure_kernels() { return cudaSuccess; }

size_t ladder_smem_bytes(int bps, int src_max, int mid_max, int vc_max, int *double_buffer, int *mid_buffers)
{
    const size_t tables = (size_t)geom_slots(bps) * vc_max;
    const size_t two_mid = (size_t)src_max + 2 * (size_t)mid_max + tables + 16, one = (size_t)src_max + mid_max + tables + 16;
    const int mb = (two_mid <= kMaxTileSmem && !(OTTGPU_STRIP && bps == 1)) ? 2 : 1;
    if (double_buffer) *double_buffer = 0;
    if (mid_buffers) *mid_buffers = mb;
    return mb == 2 ? two_mid : one;
}

// the B200 figure: 228 KB per SM, 1 KB reserved + ~1 KB static per CTA
int ladder_ctas_per_sm(size_t smem_bytes) { return smem_bytes > kMaxTileSmem ? 0 : (smem_bytes <= 114816 ? 2 : 1); }

cudaError_t launch_ladder(LadderLaunch L, cudaStream_t)
{
    int db = 0, mb = 1;
    const size_t smem_bytes = ladder_smem_bytes(L.bps, L.src_max, L.mid_max, L.vc_max, &db, &mb);
    if (smem_bytes > kMaxTileSmem) return cudaErrorInvalidValue;
    L.double_buffer = db;
    L.mid_buffers = mb;
    std::vector<unsigned char> smem(smem_bytes + 64);
    unsigned char *sm = smem.data() + ((16 - ((uintptr_t)smem.data() & 15)) & 15);
    for (int blk = 0; blk < L.n_items; ++blk) {
        const Item it = L.items[blk];
        const Binding &b = L.binds[it.chan];
        TileGeom g;
        static_cast<Item &>(g) = it;                          // the producer copies the static block, then patches the step's pointers in
        const int what = item_prepare(b, L, db ? (blk & 1) : 0, mb == 2 ? (blk & 1) : 0, blk & (geom_slots(L.bps) - 1), g);   // alternate buffers like the kernel does
        if (what == 0) continue;
        if (what == 2) {
            for (int tid = 0; tid < kThreads; ++tid) item_copy(tid, it, b);
            continue;
        }
        // every byte the phases touch must lie inside the launch's dynamic shared memory
        if ((size_t)g.off_vpos + (size_t)g.th * 4 > smem_bytes) return cudaErrorInvalidValue;
        if ((size_t)(g.staged_planes - 1) * g.plane_stride + (size_t)g.sr * g.spw * 4 > (size_t)L.src_max) return cudaErrorInvalidValue;
        if ((size_t)g.nplanes * g.mid_rows * g.mid_pitch * 2 > (size_t)L.mid_max) return cudaErrorInvalidValue;
        if ((size_t)g.off_mid + L.mid_max > (size_t)g.off_vc) return cudaErrorInvalidValue;
        if (g.nvec * 4 > g.spw) return cudaErrorInvalidValue;
        {                                                 // poison: stale-data reads show up as mismatches (noise, as on the GPU)
            uint32_t lcg = 12345u + (uint32_t)blk;
            for (auto &byte : smem) { lcg = lcg * 1664525u + 1013904223u; byte = (unsigned char)(lcg >> 24); }
        }
        for (int lane = 0; lane < 32; ++lane) phase_tables(lane, 32, g, sm);
        for (int tid = 0; tid < kThreads; ++tid) phase_stage(tid, g, sm);
        for (int tid = 0; tid < kThreads; ++tid) phase_fixup16(tid, g, sm);
        if (g.bps != L.bps) return cudaErrorInvalidValue;   // the launcher partitions items by sample size
        if (L.bps == 2) {
            for (int tid = 0; tid < kThreads; ++tid) phase_hpass<2>(tid, g, sm);
            for (int tid = 0; tid < kThreads; ++tid) phase_vpass<2>(tid, g, sm);
        } else {
            for (int tid = 0; tid < kThreads; ++tid) phase_hpass<1>(tid, g, sm);
            for (int tid = 0; tid < kThreads; ++tid) phase_vpass<1>(tid, g, sm);
        }
    }
    return cudaSuccess;
}

cudaError_t launch_yadif(const YadifJob *jobs, int n_jobs, int max_w, int max_h, bool, cudaStream_t)
{
    for (int z = 0; z < n_jobs; ++z) {
        if (!jobs[z].active) continue;
        for (int y = 0; y < ((max_h + (max_h + 1) / 2) + 7) / 8 * 8; ++y)
            for (int x = 0; x < (max_w + 31) / 32 * 32; ++x) {
                if (jobs[z].fast) yadif_position_fast(jobs[z], x, y);
                else yadif_position(jobs[z], x, y);
            }
    }
    return cudaSuccess;
}

cudaError_t launch_convert_batch(const ConvertJob *jobs, int n_jobs, int max_w, int max_rows, cudaStream_t)
{
    for (int z = 0; z < n_jobs; ++z)
        for (int r = 0; r < (max_rows + 7) / 8 * 8; ++r)
            for (int x = 0; x < ((max_w + kConvertGroup - 1) / kConvertGroup + 31) / 32 * 32; ++x) convert_position(jobs[z], x, r);
    return cudaSuccess;
}

static void emu_scan(uint32_t *v, int n)
{
    uint32_t run = 0;
    for (int i = 0; i < n; ++i) { const uint32_t x = v[i]; v[i] = run; run += x; }
    v[n] = run;
}

int jpeg_stuff_chunks(uint32_t stream_words) { return (int)(((size_t)stream_words * 4 + kJpegChunk - 1) / kJpegChunk); }

cudaError_t launch_jpeg_analyse(const JpegJob &j, cudaStream_t)
{
    if (j.n_blocks <= 0) return cudaSuccess;
    std::memset(j.hist, 0, 4 * 256 * sizeof(uint32_t));
    for (int b = 0; b < (j.n_blocks + 127) / 128 * 128; ++b) jpeg_block_dct(j, b);
    for (int b = 0; b < (j.n_blocks + 127) / 128 * 128; ++b) jpeg_block_hist(j, b);
    return cudaSuccess;
}

cudaError_t launch_jpeg_encode(const JpegJob &j, cudaStream_t)
{
    if (j.n_blocks <= 0) return cudaSuccess;
    std::memset(j.stream, 0, (size_t)j.stream_words * 4);
    for (int b = 0; b < (j.n_blocks + 127) / 128 * 128; ++b) jpeg_block_bits(j, b);
    emu_scan(j.bits, j.n_blocks);
    for (int b = j.n_blocks - 1; b >= 0; --b) jpeg_block_emit(j, b);        // any order: the positions are absolute
    const int chunks = jpeg_stuff_chunks(j.stream_words);
    std::memset(j.ff_count, 0, (size_t)(chunks + 1) * 4);
    for (int c = 0; c < chunks; ++c) jpeg_stuff_count(j, c);
    emu_scan(j.ff_count, chunks);
    for (int c = 0; c < chunks; ++c) jpeg_stuff_write(j, c);
    jpeg_finish(j, chunks);
    return cudaSuccess;
}

cudaError_t launch_convert(const ConvertJob &job, cudaStream_t)
{
    const int rows = job.plane[0].h + job.plane[1].h + job.plane[2].h;
    const int groups = (job.plane[0].w + kConvertGroup - 1) / kConvertGroup;
    for (int r = 0; r < (rows + 7) / 8 * 8; ++r)
        for (int x = 0; x < (groups + 31) / 32 * 32; ++x) convert_position(job, x, r);
    return cudaSuccess;
}

}  // namespace ottgpu
