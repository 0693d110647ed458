// Launch entry points of the sm_100a kernels (kernels.cu).  Host code never sees __global__ symbols.
#pragma once
#include <cuda_runtime.h>
#include "device_types.h"

namespace ottgpu {

// upper bound of dynamic shared memory the ladder kernel is configured for (sm_100 allows 227 KB per CTA)
constexpr size_t kMaxTileSmem = 224 * 1024;

// shared memory one launch needs for the given region maxima; the number of source / intermediate buffers is decided here
size_t ladder_smem_bytes(int bps, int src_max, int mid_max, int vc_max, int *double_buffer, int *mid_buffers = nullptr);

// resident CTAs of the ladder kernel per SM for a launch that needs `smem_bytes` of dynamic shared memory (0 on error)
int ladder_ctas_per_sm(size_t smem_bytes);

// Persistent launch: grid = min(n_items, SMs x resident CTAs); CTAs pull work items from L.counter (must be zero).
cudaError_t launch_ladder(LadderLaunch L, cudaStream_t s);

// yadif: one launch over all jobs (blockIdx.z = job).
// all_fast: every job takes the 8-samples-per-thread path, so the grid only spans (w/8) x (h + h/2) positions
cudaError_t launch_yadif(const YadifJob *jobs, int n_jobs, int max_w, int max_h, bool all_fast, cudaStream_t s);

// decode-side converter: one launch per picture
cudaError_t launch_convert(const ConvertJob &job, cudaStream_t s);
// many pictures in one launch; jobs live in device memory; max_w / max_rows bound the grid (rows = Y + U + V rows)
cudaError_t launch_convert_batch(const ConvertJob *jobs, int n_jobs, int max_w, int max_rows, cudaStream_t s);

// thumbnail JPEG, two phases on `s` with the job passed by value (kernel parameter):
//   analyse: DCT + quantiser -> j.coef, symbol histogram -> j.hist (zeroed here); the host builds the picture's tables
//   encode : entropy length / scan / emit / byte stuffing with the tables in j.huff (j.stream and j.ff_count zeroed here)
struct JpegJob;
cudaError_t launch_jpeg_analyse(const JpegJob &j, cudaStream_t s);
cudaError_t launch_jpeg_encode(const JpegJob &j, cudaStream_t s);
int jpeg_stuff_chunks(uint32_t stream_words);

cudaError_t configure_kernels();

}  // namespace ottgpu
