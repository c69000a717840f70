// K3: voxel-grid merge of 10-byte records (oracle/SPEC.md s3 -- own integer spec; the
// reference only #includes pcl/filters/voxel_grid.h, src/pcs-multicamera-optimized.cpp:17).
//
//   1. key[i] = (kz, ky, kx) packed most-significant first, idx[i] = i
//   2. stable LSD radix sort of (key, idx), 8 bits per pass, ceil(3*bits/8) passes
//   3. head flags + exclusive scan -> voxel id per sorted element, voxel count
//   4. warp-shuffle segmented reduction of the integer sums; one atomicAdd per
//      warp-segment into the per-voxel accumulators
//   5. finalize: integer means -> one record per voxel, already in ascending key order
//
// All integer, so the result equals the CPU restatement bit for bit regardless of the
// order in which the sums are formed.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <utility>

namespace pcs {

constexpr int RS_THREADS = 256, RS_ITEMS = 8, RS_TILE = RS_THREADS * RS_ITEMS;
constexpr int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

struct VoxelScratch {
    uint8_t *buf = nullptr;
    size_t cap = 0;
    int32_t *h_count = nullptr;   // pinned
    uint8_t *h_tab = nullptr;     // pinned staging for the MSD merge's plan tables
    size_t h_tab_cap = 0;
};

inline void voxel_free(VoxelScratch &s) {
    cudaFree(s.buf);
    if (s.h_count) cudaFreeHost(s.h_count);
    if (s.h_tab) cudaFreeHost(s.h_tab);
    s.buf = nullptr; s.cap = 0; s.h_count = nullptr; s.h_tab = nullptr; s.h_tab_cap = 0;
}

struct VoxelGeom {
    int leaf, kmin, bits;   // key field = k - kmin, `bits` wide
};

__device__ __forceinline__ int floordiv_i(int a, int b) {
    int q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

__global__ void __launch_bounds__(256)
vox_keys(const int16_t *__restrict__ rec, int n, VoxelGeom g, uint64_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int16_t *r = rec + 5 * (size_t)i;
    const uint64_t kx = (uint64_t)(floordiv_i(r[0], g.leaf) - g.kmin);
    const uint64_t ky = (uint64_t)(floordiv_i(r[1], g.leaf) - g.kmin);
    const uint64_t kz = (uint64_t)(floordiv_i(r[2], g.leaf) - g.kmin);
    keys[i] = (kz << (2 * g.bits)) | (ky << g.bits) | kx;
    idx[i] = (uint32_t)i;
}

// ---- radix sort pass: per-tile digit histogram (digit-major table) -----------------
__global__ void __launch_bounds__(RS_THREADS)
rs_hist(const uint64_t *__restrict__ keys, int n, int shift, uint32_t *__restrict__ table, int n_tiles) {
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * RS_TILE;
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int i = base + k * RS_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&hist[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    table[(size_t)threadIdx.x * n_tiles + blockIdx.x] = hist[threadIdx.x];
}

// ---- radix sort pass: stable scatter -------------------------------------------------
__global__ void __launch_bounds__(RS_THREADS)
rs_scatter(const uint64_t *__restrict__ keys_in, const uint32_t *__restrict__ idx_in,
           uint64_t *__restrict__ keys_out, uint32_t *__restrict__ idx_out, int n, int shift,
           const uint32_t *__restrict__ table, int n_tiles) {
    __shared__ uint32_t whist[RS_THREADS / 32][256];
    __shared__ uint32_t gbase[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int k = threadIdx.x; k < (RS_THREADS / 32) * 256; k += RS_THREADS) (&whist[0][0])[k] = 0;
    __syncthreads();
    // warp-blocked order: warp w owns elements [w*256, w*256+256) of the tile, item-major
    const int wbase = blockIdx.x * RS_TILE + warp * (32 * RS_ITEMS);
    uint64_t key[RS_ITEMS];
    uint32_t id[RS_ITEMS], local[RS_ITEMS];
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        const bool valid = i < n;
        key[k] = valid ? keys_in[i] : 0;
        id[k] = valid ? idx_in[i] : 0;
        const uint32_t d = valid ? ((uint32_t)(key[k] >> shift) & 255u) : (256u + lane);
        const uint32_t mask = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(mask) - 1;
        const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
        uint32_t prev = 0;
        if (valid && lane == leader) {
            prev = whist[warp][d];
            whist[warp][d] = prev + __popc(mask);
        }
        __syncwarp();
        prev = __shfl_sync(0xffffffffu, prev, leader);
        local[k] = prev + rank;
    }
    __syncthreads();
    {   // exclusive scan over warps for digit = threadIdx.x
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) {
            const uint32_t c = whist[w][threadIdx.x];
            whist[w][threadIdx.x] = run;
            run += c;
        }
        gbase[threadIdx.x] = table[(size_t)threadIdx.x * n_tiles + blockIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < RS_ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        if (i < n) {
            const uint32_t d = (uint32_t)(key[k] >> shift) & 255u;
            const uint32_t pos = gbase[d] + whist[warp][d] + local[k];
            keys_out[pos] = key[k];
            idx_out[pos] = id[k];
        }
    }
}

// ---- generic exclusive scan of uint32 (three kernels) -------------------------------
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_tot, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        const uint32_t t = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
        uint32_t i2 = t;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, i2, d);
            if (lane >= d) i2 += u;
        }
        warp_tot[lane] = i2 - t;
        if (lane == 31) warp_tot[32] = i2;
    }
    __syncthreads();
    total = warp_tot[32];
    return warp_tot[warp] + inc - v;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_tiles(uint32_t *__restrict__ data, int n, uint32_t *__restrict__ tile_sums) {
    __shared__ uint32_t warp_tot[33];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? data[base + k] : 0;
        sum += v[k];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, warp_tot, total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) data[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// single block; also leaves the grand total in sums[n]
__global__ void __launch_bounds__(1024)
scan_sums(uint32_t *__restrict__ sums, int n) {
    __shared__ uint32_t warp_tot[33];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < n ? sums[i] : 0;
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(v, warp_tot, total);
        const uint32_t carry = carry_s;
        if (i < n) sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[n] = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_add(uint32_t *__restrict__ data, int n, const uint32_t *__restrict__ tile_sums) {
    const uint32_t add = tile_sums[blockIdx.x];
    const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) data[base + k] += add;
}

// exclusive scan of data[0..n) in place; grand total ends up in sums[n_tiles]
inline void exclusive_scan_u32(uint32_t *data, int n, uint32_t *sums, cudaStream_t cs) {
    const int nt = (n + SCAN_TILE - 1) / SCAN_TILE;
    scan_tiles<<<nt, SCAN_THREADS, 0, cs>>>(data, n, sums);
    scan_sums<<<1, 1024, 0, cs>>>(sums, nt);
    scan_add<<<nt, SCAN_THREADS, 0, cs>>>(data, n, sums);
}

// ---- segments ---------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vox_heads(const uint64_t *__restrict__ keys, int n, uint32_t *__restrict__ head) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1u : 0u;
}

// vid_ex = exclusive scan of the head flags: element i belongs to voxel vid_ex[i] + head - 1,
// where head is recomputed from the keys.
__global__ void __launch_bounds__(256)
vox_accumulate(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ idx,
               const uint32_t *__restrict__ vid_ex, const int16_t *__restrict__ rec, int n, VoxelGeom g,
               uint32_t *__restrict__ acc, uint64_t *__restrict__ vkey) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const bool valid = i < n;
    uint32_t v[7] = {0, 0, 0, 0, 0, 0, 0};
    uint32_t vid = 0xFFFFFFFFu;
    if (valid) {
        const uint64_t key = keys[i];
        const bool head = (i == 0) || key != keys[i - 1];
        vid = vid_ex[i] + (head ? 1u : 0u) - 1u;
        if (head) vkey[vid] = key;
        const int16_t *r = rec + 5 * (size_t)idx[i];
        const int x = r[0], y = r[1], z = r[2];
        const uint32_t s3 = (uint16_t)r[3], s4 = (uint16_t)r[4];
        v[0] = (uint32_t)(x - g.leaf * floordiv_i(x, g.leaf));
        v[1] = (uint32_t)(y - g.leaf * floordiv_i(y, g.leaf));
        v[2] = (uint32_t)(z - g.leaf * floordiv_i(z, g.leaf));
        v[3] = s3 & 0xFF; v[4] = s3 >> 8; v[5] = s4 & 0xFF; v[6] = 1;
    }
    // warp-shuffle segmented inclusive scan (segments = runs of equal vid)
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t ovid = __shfl_up_sync(0xffffffffu, vid, d);
        uint32_t o[7];
#pragma unroll
        for (int c = 0; c < 7; ++c) o[c] = __shfl_up_sync(0xffffffffu, v[c], d);
        if (lane >= d && ovid == vid) {
#pragma unroll
            for (int c = 0; c < 7; ++c) v[c] += o[c];
        }
    }
    const uint32_t nvid = __shfl_down_sync(0xffffffffu, vid, 1);
    if (valid && (lane == 31 || nvid != vid)) {
#pragma unroll
        for (int c = 0; c < 7; ++c) atomicAdd(acc + 8 * (size_t)vid + c, v[c]);
    }
}

__global__ void __launch_bounds__(256)
vox_finalize(const uint32_t *__restrict__ acc, const uint64_t *__restrict__ vkey, int nv, VoxelGeom g,
             int16_t *__restrict__ out) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= nv) return;
    const uint32_t *a = acc + 8 * (size_t)v;
    const uint64_t key = vkey[v], m = (1ull << g.bits) - 1;
    const int kx = (int)(key & m) + g.kmin, ky = (int)((key >> g.bits) & m) + g.kmin,
              kz = (int)((key >> (2 * g.bits)) & m) + g.kmin;
    const uint32_t cnt = a[6];
    int16_t *o = out + 5 * (size_t)v;
    o[0] = (int16_t)(g.leaf * kx + (int)(a[0] / cnt));
    o[1] = (int16_t)(g.leaf * ky + (int)(a[1] / cnt));
    o[2] = (int16_t)(g.leaf * kz + (int)(a[2] / cnt));
    o[3] = (int16_t)((a[3] / cnt) | ((a[4] / cnt) << 8));
    o[4] = (int16_t)(a[5] / cnt);
}

// Returns the voxel count (>= 0), -2 on a CUDA error, -3 on allocation failure.
// Synchronises `cs` once (the count has to reach the host); the last two kernels are left in flight.
inline int voxel_merge(VoxelScratch &s, const int16_t *rec, int n, int leaf, int16_t *out, cudaStream_t cs) {
    VoxelGeom g;
    g.leaf = leaf;
    g.kmin = -(32768 + leaf - 1) / leaf;                 // floor(-32768 / leaf)
    const int kmax = 32767 / leaf;
    g.bits = 1;
    while ((1 << g.bits) < (kmax - g.kmin + 1)) ++g.bits;
    const int passes = (3 * g.bits + 7) / 8;
    const int n_tiles = (n + RS_TILE - 1) / RS_TILE;
    const int n_scan_tiles_tab = (256 * n_tiles + SCAN_TILE - 1) / SCAN_TILE;
    const int n_scan_tiles_n = (n + SCAN_TILE - 1) / SCAN_TILE;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t o_k0 = 0, o_k1 = o_k0 + al((size_t)n * 8), o_i0 = o_k1 + al((size_t)n * 8),
                 o_i1 = o_i0 + al((size_t)n * 4), o_tab = o_i1 + al((size_t)n * 4),
                 o_sums = o_tab + al((size_t)256 * n_tiles * 4),
                 o_head = o_sums + al((size_t)(std::max(n_scan_tiles_tab, n_scan_tiles_n) + 2) * 4),
                 o_acc = o_head + al((size_t)n * 4), o_vkey = o_acc + al((size_t)n * 32),
                 total = o_vkey + al((size_t)n * 8);
    if (total > s.cap) {
        cudaFree(s.buf);
        s.buf = nullptr; s.cap = 0;
        if (cudaMalloc(&s.buf, total) != cudaSuccess) { cudaGetLastError(); return -3; }
        s.cap = total;
    }
    if (!s.h_count && cudaHostAlloc(&s.h_count, 64, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return -3; }
    uint64_t *k0 = (uint64_t *)(s.buf + o_k0), *k1 = (uint64_t *)(s.buf + o_k1);
    uint32_t *i0 = (uint32_t *)(s.buf + o_i0), *i1 = (uint32_t *)(s.buf + o_i1);
    uint32_t *tab = (uint32_t *)(s.buf + o_tab), *sums = (uint32_t *)(s.buf + o_sums);
    uint32_t *head = (uint32_t *)(s.buf + o_head), *acc = (uint32_t *)(s.buf + o_acc);
    uint64_t *vkey = (uint64_t *)(s.buf + o_vkey);
    const int nb = (n + 255) / 256;
    vox_keys<<<nb, 256, 0, cs>>>(rec, n, g, k0, i0);
    for (int p = 0; p < passes; ++p) {
        rs_hist<<<n_tiles, RS_THREADS, 0, cs>>>(k0, n, 8 * p, tab, n_tiles);
        exclusive_scan_u32(tab, 256 * n_tiles, sums, cs);
        rs_scatter<<<n_tiles, RS_THREADS, 0, cs>>>(k0, i0, k1, i1, n, 8 * p, tab, n_tiles);
        std::swap(k0, k1);
        std::swap(i0, i1);
    }
    vox_heads<<<nb, 256, 0, cs>>>(k0, n, head);
    exclusive_scan_u32(head, n, sums, cs);
    if (cudaMemcpyAsync(s.h_count, sums + n_scan_tiles_n, 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(cs) != cudaSuccess) return -2;
    const int nv = *s.h_count;
    if (nv < 1 || nv > n) return -2;
    if (cudaMemsetAsync(acc, 0, (size_t)nv * 32, cs) != cudaSuccess) return -2;
    vox_accumulate<<<nb, 256, 0, cs>>>(k0, i0, head, rec, n, g, acc, vkey);
    vox_finalize<<<(nv + 255) / 256, 256, 0, cs>>>(acc, vkey, nv, g, out);
    if (cudaGetLastError() != cudaSuccess) return -2;
    return nv;
}

}  // namespace pcs
