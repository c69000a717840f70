// tools/sort_probe.cu -- measurement aid, NOT part of the product: what the toolkit's own radix sort
// (CUB one-sweep) and a few primitive operations cost on this GPU, to size the voxel merge against.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gpurun_out/sort_probe tools/sort_probe.cu
//   gpurun_out/sort_probe [n]
#include <cub/cub.cuh>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void fill_words(uint64_t *w, int n, int key_bits, int idx_bits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t key = (((uint64_t)mix(i) << 32) | mix(i + 0x9e3779b9u)) & ((1ull << key_bits) - 1);
    w[i] = (key << idx_bits) | (uint64_t)i;
}

// one shared-memory atomicAdd per thread per iteration, random addresses in a table of `bins`
__global__ void k_atoms(uint32_t *out, int iters, int bins) {
    extern __shared__ uint32_t tab[];
    for (int k = threadIdx.x; k < bins; k += blockDim.x) tab[k] = 0;
    __syncthreads();
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x), acc = 0;
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        acc += atomicAdd(&tab[(s >> 8) % bins], 1u);
    }
    __syncthreads();
    if (acc == 0xFFFFFFFFu) out[0] = tab[0];
}
// the same with plain load + store (no atomicity): the non-atomic floor
__global__ void k_ldsts(uint32_t *out, int iters, int bins) {
    extern __shared__ uint32_t tab[];
    for (int k = threadIdx.x; k < bins; k += blockDim.x) tab[k] = 0;
    __syncthreads();
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x), acc = 0;
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        const uint32_t a = (s >> 8) % bins;
        const uint32_t v = tab[a];
        tab[a] = v + 1;
        acc += v;
    }
    __syncthreads();
    if (acc == 0xFFFFFFFFu) out[0] = tab[0];
}
__global__ void k_match(uint32_t *out, int iters, int bins) {
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x), acc = 0;
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        acc += __popc(__match_any_sync(0xffffffffu, (s >> 8) % bins));
    }
    if (acc == 0xFFFFFFFFu) out[0] = acc;
}
__global__ void k_ballot8(uint32_t *out, int iters, int bins) {
    uint32_t s = mix(blockIdx.x * blockDim.x + threadIdx.x), acc = 0;
    for (int i = 0; i < iters; ++i) {
        s = s * 1664525u + 1013904223u;
        const uint32_t d = (s >> 8) % bins;
        uint32_t m = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < 8; ++b) {
            const bool bit = (d >> b) & 1u;
            const uint32_t v = __ballot_sync(0xffffffffu, bit);
            m &= bit ? v : ~v;
        }
        acc += __popc(m);
    }
    if (acc == 0xFFFFFFFFu) out[0] = acc;
}
// scattered 8-byte stores: element i goes to a pseudo-random slot (a permutation: i * odd mod 2^k)
__global__ void k_scatter8(const uint64_t *in, uint64_t *out, int n, uint32_t mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = ((uint32_t)i * 2654435761u) & mask;
    if (j < (uint32_t)n) out[j] = in[i];
}
__global__ void k_gather8(const uint64_t *in, uint64_t *out, int n, uint32_t mask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t j = ((uint32_t)i * 2654435761u) & mask;
    out[i] = j < (uint32_t)n ? in[j] : 0;
}
__global__ void k_copy8(const uint64_t *in, uint64_t *out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
// one global RED per element on `bins` random addresses
__global__ void k_red(const uint64_t *in, uint32_t *hist, int n, int bins) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    atomicAdd(hist + (uint32_t)(in[i] >> 24) % bins, 1u);
}
// one global atomic WITH return per element
__global__ void k_atomg(const uint64_t *in, uint32_t *hist, uint32_t *out, int n, int bins) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = atomicAdd(hist + (uint32_t)(in[i] >> 24) % bins, 1u);
}

template <class F> float timed(F f, int iters = 10) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int i = 0; i < 2; ++i) f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    for (int i = 0; i < iters; ++i) f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, a, b));
    CK(cudaGetLastError());
    return ms / iters;
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 14745600;
    const int idx_bits = 24;
    uint64_t *a, *b, *pristine;
    CK(cudaMalloc(&a, (size_t)n * 8)); CK(cudaMalloc(&b, (size_t)n * 8)); CK(cudaMalloc(&pristine, (size_t)n * 8));
    uint32_t *hist, *o32;
    CK(cudaMalloc(&hist, 4 << 20)); CK(cudaMalloc(&o32, (size_t)n * 4));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"n\": %d", prop.name, sms, n);
    for (int key_bits : {31, 24, 16, 39}) {
        fill_words<<<(n + 255) / 256, 256>>>(pristine, n, key_bits, idx_bits);
        size_t tmp_bytes = 0;
        cub::DoubleBuffer<uint64_t> db(a, b);
        CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, db, n, idx_bits, idx_bits + key_bits));
        void *tmp;
        CK(cudaMalloc(&tmp, tmp_bytes));
        const float ms = timed([&] {
            cudaMemcpyAsync(a, pristine, (size_t)n * 8, cudaMemcpyDeviceToDevice);
            cub::DoubleBuffer<uint64_t> d2(a, b);
            cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, d2, n, idx_bits, idx_bits + key_bits);
        });
        const float ms_copy = timed([&] { cudaMemcpyAsync(a, pristine, (size_t)n * 8, cudaMemcpyDeviceToDevice); });
        printf(", \"cub_sort_u64_%dbits_ms\": %.4f", key_bits, ms - ms_copy);
        CK(cudaFree(tmp));
    }
    // (key32, value32) pairs
    {
        uint32_t *k0 = (uint32_t *)a, *k1 = (uint32_t *)b, *v0 = o32, *v1;
        CK(cudaMalloc(&v1, (size_t)n * 4));
        size_t tmp_bytes = 0;
        cub::DoubleBuffer<uint32_t> dk(k0, k1), dv(v0, v1);
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, dk, dv, n, 0, 31));
        void *tmp;
        CK(cudaMalloc(&tmp, tmp_bytes));
        const float ms = timed([&] {
            cudaMemcpyAsync(k0, pristine, (size_t)n * 4, cudaMemcpyDeviceToDevice);
            cub::DoubleBuffer<uint32_t> d2(k0, k1), e2(v0, v1);
            cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, d2, e2, n, 0, 31);
        });
        const float ms_copy = timed([&] { cudaMemcpyAsync(k0, pristine, (size_t)n * 4, cudaMemcpyDeviceToDevice); });
        printf(", \"cub_sort_pairs_u32_u32_31bits_ms\": %.4f", ms - ms_copy);
        CK(cudaFree(tmp)); CK(cudaFree(v1));
    }
    fill_words<<<(n + 255) / 256, 256>>>(pristine, n, 31, idx_bits);
    uint32_t mask = 1;
    while (mask < (uint32_t)n) mask <<= 1;
    mask -= 1;
    printf(", \"copy8_ms\": %.4f", timed([&] { k_copy8<<<(n + 255) / 256, 256>>>(pristine, a, n); }));
    printf(", \"scatter8_ms\": %.4f", timed([&] { k_scatter8<<<(n + 255) / 256, 256>>>(pristine, a, n, mask); }));
    printf(", \"gather8_ms\": %.4f", timed([&] { k_gather8<<<(n + 255) / 256, 256>>>(pristine, a, n, mask); }));
    for (int bins : {256, 65536, 1 << 20}) {
        printf(", \"red_global_%d_bins_ms\": %.4f", bins, timed([&] { k_red<<<(n + 255) / 256, 256>>>(pristine, hist, n, bins); }));
        printf(", \"atom_global_ret_%d_bins_ms\": %.4f", bins, timed([&] { k_atomg<<<(n + 255) / 256, 256>>>(pristine, hist, o32, n, bins); }));
    }
    // per-SM primitive rates: grid = 4 CTAs x 256 threads per SM, `iters` operations per thread
    const int iters = 4096, grid = sms * 4;
    const double ops = (double)grid * 256 * iters;
    for (int bins : {256, 4096}) {
        float ms = timed([&] { k_atoms<<<grid, 256, bins * 4>>>(hist, iters, bins); }, 3);
        printf(", \"atoms_%d_lanes_per_clk_per_sm\": %.3f", bins, ops / sms / (ms * 1e-3 * prop.clockRate * 1e3));
        ms = timed([&] { k_ldsts<<<grid, 256, bins * 4>>>(hist, iters, bins); }, 3);
        printf(", \"ldsts_%d_lanes_per_clk_per_sm\": %.3f", bins, ops / sms / (ms * 1e-3 * prop.clockRate * 1e3));
    }
    {
        float ms = timed([&] { k_match<<<grid, 256>>>(hist, iters, 256); }, 3);
        printf(", \"match_any_lanes_per_clk_per_sm\": %.3f", ops / sms / (ms * 1e-3 * prop.clockRate * 1e3));
        ms = timed([&] { k_ballot8<<<grid, 256>>>(hist, iters, 256); }, 3);
        printf(", \"ballot8_lanes_per_clk_per_sm\": %.3f", ops / sms / (ms * 1e-3 * prop.clockRate * 1e3));
    }
    printf(", \"clock_khz\": %d}\n", prop.clockRate);
    return 0;
}
