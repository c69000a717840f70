// Voxel-grid merge (oracle/SPEC.md s3) -- placeholder interface; the kernels land next.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcs {
struct VoxelScratch { void *buf = nullptr; size_t cap = 0; };
inline void voxel_free(VoxelScratch &s) { cudaFree(s.buf); s.buf = nullptr; s.cap = 0; }
inline int voxel_merge(VoxelScratch &, const int16_t *, int, int, int16_t *, cudaStream_t) { return -4; }
}  // namespace pcs
