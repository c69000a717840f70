// Multi-GPU voxel merge, sharded BEFORE the exchange ("sharded by voxel-key range with one
// all-to-all").  The reference fans every camera's records into one host (readCloud,
// src/pcs-multicamera-client.cpp:363-371) because the concat (:385-392) needs them all in one place; a voxel
// merge does not -- cameras are independent up to the merge, and a voxel only needs the points of ITS key
// range.  So every GPU keeps its own cameras' records and
//
//   1. xa_zhist    counts its points per z plane of the voxel grid (into symmetric memory);
//   2. xa_plan     every GPU reads all ranks' histograms over NVLink (a few KB each), adds them and cuts the z axis
//                  into n_ranks slabs of equal population -- the same cuts everywhere, no host, no collective;
//   3. xa_scatter  the all-to-all: each tile of records is binned by slab with shared-memory atomics, a run is
//                  reserved in the destination GPU's inbox with one system-scope atomicAdd per (tile, rank), and the
//                  records leave as coalesced peer stores (NVLink carries each record once: 10 B/pt x (N-1)/N);
//   4. every GPU merges its inbox (pcs_voxel_sweep.cuh, point count read on the device).
//
// Slab r of the grid ends up on GPU r; concatenated in rank order the slabs are the single-GPU merge, bit for bit
// (voxels are emitted in ascending (kz, ky, kx) order).  Barriers between the steps are the host's (one after
// step 1, one after step 3).
#pragma once
#include "pcs_voxel_sweep.cuh"

namespace pcs {

constexpr int XA_MAX_RANKS = 8;
constexpr int XA_THREADS = 256, XA_ITEMS = 16, XA_TILE = XA_THREADS * XA_ITEMS;

struct XaPeers {
    int n_ranks, rank;
    uint16_t *inbox[XA_MAX_RANKS];          // records received, per rank (peer-mapped)
    uint32_t *cursor[XA_MAX_RANKS];         // records reserved so far in that inbox
    const uint32_t *zhist[XA_MAX_RANKS];    // points per z plane, per rank
    long long capacity;                     // records an inbox can hold
};

// ---- 1. points per z plane of this rank's records ---------------------------------------------------
__global__ void __launch_bounds__(SW_KH_THREADS)
xa_zhist(const int16_t *__restrict__ rec, int n, SweepGeom g, uint32_t *__restrict__ zhist, int zbins) {
    extern __shared__ uint32_t xa_zh[];
    for (int k = threadIdx.x; k < zbins; k += SW_KH_THREADS) xa_zh[k] = 0;
    __syncthreads();
    const bool aligned = (((uintptr_t)rec) & 15) == 0;
    const int n_tiles = (n + SW_KH_TILE - 1) / SW_KH_TILE;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int first = tile * SW_KH_TILE + threadIdx.x * SW_KH_ITEMS;
        const int cnt = max(0, min(SW_KH_ITEMS, n - first));
        if (cnt == 0) continue;
        uint32_t w[20];
        sw_load8(rec, n, first, cnt == SW_KH_ITEMS && aligned, w);
#pragma unroll
        for (int k = 0; k < SW_KH_ITEMS; ++k)
            if (k < cnt) atomicAdd(xa_zh + sw_q(sw_half(w, 5 * k + 2), g), 1u);
    }
    __syncthreads();
    for (int k = threadIdx.x; k < zbins; k += SW_KH_THREADS) {
        const uint32_t v = xa_zh[k];
        if (v) atomicAdd(zhist + k, v);
    }
}

// ---- 2. equal-population cuts of the z axis, identical on every rank ---------------------------------
// splits[r] .. splits[r + 1] (plane indices floor(z / leaf)) is rank r's slab; zslab[q] = rank owning plane q
// (q = plane + K, the biased index the kernels use).  One block.
__global__ void __launch_bounds__(1024)
xa_plan(XaPeers peers, int zbins, int K, int32_t *__restrict__ splits, uint8_t *__restrict__ zslab) {
    extern __shared__ uint32_t xa_c[];        // inclusive prefix of the summed histogram
    __shared__ uint32_t warp_tot[33];
    __shared__ uint32_t s_carry;
    __shared__ int s_lo, s_hi, s_end[XA_MAX_RANKS + 1];
    if (threadIdx.x == 0) { s_carry = 0; s_lo = zbins; s_hi = -1; }
    __syncthreads();
    // summed histogram -> inclusive prefix, 1024 planes per round
    for (int base = 0; base < zbins; base += 1024) {
        const int q = base + threadIdx.x;
        uint32_t h = 0;
        if (q < zbins)
            for (int r = 0; r < peers.n_ranks; ++r) h += peers.zhist[r][q];
        if (h) { atomicMin(&s_lo, q); atomicMax(&s_hi, q); }
        uint32_t total;
        const uint32_t ex = block_exclusive_scan(h, warp_tot, total);
        const uint32_t carry = s_carry;
        if (q < zbins) xa_c[q] = carry + ex + h;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + total;
        __syncthreads();
    }
    const unsigned long long n = s_carry;
    const int lo = s_lo, hi = s_hi;
    if (threadIdx.x <= peers.n_ranks) {
        const int r = threadIdx.x;         // end of slab r - 1 == start of slab r
        int e;
        if (hi < lo) e = 0;                               // no points at all
        else if (r == 0) e = lo;
        else if (r == peers.n_ranks) e = hi + 1;
        else {
            const unsigned long long target = n * (unsigned long long)r / (unsigned long long)peers.n_ranks;
            if (target == 0) e = lo;
            else {                                        // first plane at which the running count reaches the target
                int a = lo, b = hi;
                while (a < b) {
                    const int mid = (a + b) >> 1;
                    if (xa_c[mid] >= target) b = mid; else a = mid + 1;
                }
                e = a + 1;
            }
        }
        s_end[r] = e;
        splits[r] = e - K;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < zbins; q += 1024) {
        int owner = peers.n_ranks - 1;
        for (int r = peers.n_ranks - 1; r >= 0; --r)
            if (q < s_end[r + 1]) owner = r;
        zslab[q] = (uint8_t)owner;
    }
}

// ---- 3. the all-to-all ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(XA_THREADS)
xa_scatter(const int16_t *__restrict__ rec, int n, SweepGeom g, XaPeers peers, const uint8_t *__restrict__ zslab, int zbins,
           uint32_t *__restrict__ err) {
    extern __shared__ __align__(16) uint8_t xa_smem[];
    uint16_t *stage = reinterpret_cast<uint16_t *>(xa_smem);               // [XA_TILE][5] records, grouped by destination
    uint8_t *stage_d = reinterpret_cast<uint8_t *>(stage + XA_TILE * 5);    // [XA_TILE] destination of every staged record
    uint8_t *zs = stage_d + XA_TILE;                                        // [zbins]
    __shared__ uint32_t cnt[XA_MAX_RANKS], excl[XA_MAX_RANKS + 1], gbase[XA_MAX_RANKS];
    for (int k = threadIdx.x; k < zbins; k += XA_THREADS) zs[k] = zslab[k];
    const bool aligned = (((uintptr_t)rec) & 15) == 0;
    const int n_tiles = (n + XA_TILE - 1) / XA_TILE;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (threadIdx.x < XA_MAX_RANKS) cnt[threadIdx.x] = 0;
        __syncthreads();
        uint32_t w[XA_ITEMS / 8][20];
        uint32_t dr[XA_ITEMS];           // destination | rank << 8, or 0xFFFFFFFF
#pragma unroll
        for (int h = 0; h < XA_ITEMS / 8; ++h) {
            const int first = tile * XA_TILE + threadIdx.x * XA_ITEMS + h * 8;
            const int c = max(0, min(8, n - first));
            if (c > 0) sw_load8(rec, n, first, c == 8 && aligned, w[h]);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                dr[h * 8 + k] = 0xFFFFFFFFu;
                if (k < c) {
                    const uint32_t d = zs[sw_q(sw_half(w[h], 5 * k + 2), g)];
                    dr[h * 8 + k] = d | (atomicAdd(cnt + d, 1u) << 8);
                }
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t run = 0;
            for (int r = 0; r < peers.n_ranks; ++r) { excl[r] = run; run += cnt[r]; }
            excl[peers.n_ranks] = run;
        }
        // one system-scope atomic per (tile, destination) reserves the run in that GPU's inbox
        if (threadIdx.x < peers.n_ranks && cnt[threadIdx.x]) {
            const uint32_t b = atomicAdd_system(peers.cursor[threadIdx.x], cnt[threadIdx.x]);
            gbase[threadIdx.x] = b;
            if ((long long)b + cnt[threadIdx.x] > peers.capacity) atomicExch(err, 1u);
        }
        __syncthreads();
#pragma unroll
        for (int h = 0; h < XA_ITEMS / 8; ++h) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t v = dr[h * 8 + k];
                if (v != 0xFFFFFFFFu) {
                    const uint32_t d = v & 0xFFu, pos = excl[d] + (v >> 8);
#pragma unroll
                    for (int q = 0; q < 5; ++q) stage[pos * 5 + q] = (uint16_t)sw_half(w[h], 5 * k + q);
                    stage_d[pos] = (uint8_t)d;
                }
            }
        }
        __syncthreads();
        const uint32_t total = excl[peers.n_ranks];
        for (uint32_t j = threadIdx.x; j < total * 5; j += XA_THREADS) {
            const uint32_t p = j / 5, q = j - p * 5, d = stage_d[p];
            const long long at = (long long)gbase[d] + (p - excl[d]);
            if (at < peers.capacity) peers.inbox[d][at * 5 + q] = stage[j];
        }
        __syncthreads();
    }
}

constexpr size_t xa_scatter_smem(int zbins) { return (size_t)XA_TILE * 10 + XA_TILE + (size_t)((zbins + 15) & ~15); }

}  // namespace pcs
