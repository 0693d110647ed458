// Launch entry points of the sm_100a kernels (kernels.cu).  Host code never sees __global__ symbols.
#pragma once
#include <cuda_runtime.h>
#include "device_types.h"

namespace ottgpu {

// One CTA per work item.  smem_bytes = max tile footprint over the plans referenced by `items`.
cudaError_t launch_ladder(const Item *items, int n_items, const Binding *binds, size_t smem_bytes, cudaStream_t s);

// yadif: one launch over all jobs (blockIdx.z = job).
cudaError_t launch_yadif(const YadifJob *jobs, int n_jobs, int max_w, int max_h, cudaStream_t s);

// upper bound the kernels were configured for (cudaFuncAttributeMaxDynamicSharedMemorySize)
constexpr size_t kMaxTileSmem = 224 * 1024;
cudaError_t configure_kernels();

}  // namespace ottgpu
