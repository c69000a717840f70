// K3, second generation: voxel-grid merge with a one-sweep radix sort (oracle/SPEC.md s3 --
// own integer spec; the reference only #includes pcl/filters/voxel_grid.h,
// src/pcs-multicamera-optimized.cpp:17).
//
//   1. sw_keys_hist   one read of the records: word[i] = (kz,ky,kx) << idx_bits | i  (one u64 carries
//                     the voxel key AND the point index), pay[i] = colour + in-voxel offsets (one u64),
//                     plus the digit histograms of every pass
//   2. sw_hist_scan   exclusive scan of each pass's histogram -> global bin bases
//   3. sw_pass x P    one kernel per digit: each tile is read once and written once; tile-local
//                     ranks from match_any, tile order from a ticket, the cross-tile prefix of every
//                     digit from a decoupled look-back over per-tile status words (16 B/pt per pass)
//   4. sw_chunk_heads / sw_chunk_scan   voxel (segment) boundaries per 256-point chunk and their scan
//   5. sw_reduce      one warp per chunk: gather pay[] through the sorted indices, warp-shuffle
//                     segmented sums, integer means written straight to the output; only the segments
//                     that cross a chunk boundary go through atomics into a per-chunk slot
//   6. sw_finalize_open   integer means of those slots
//
// The sums are integers, so the result equals the CPU restatement bit for bit whatever the order
// of the additions.  Needs 3*bits + ceil(log2 n) <= 64 and 24 + 3*ceil(log2 leaf) <= 64; the caller
// falls back to the (key, idx) pair sort of pcs_voxel.cuh otherwise.
#pragma once
#include "pcs_voxel.cuh"

namespace pcs {

constexpr uint32_t SW_AGG = 1u << 30, SW_PREFIX = 2u << 30, SW_VALUE = (1u << 30) - 1;
constexpr int SW_SPIN_LIMIT = 1 << 20;      // ~0.5 s of polling: a lost predecessor becomes an error, not a hang
constexpr int SW_KH_THREADS = 256, SW_KH_ITEMS = 8, SW_KH_TILE = SW_KH_THREADS * SW_KH_ITEMS;
constexpr int SW_CHUNK_ROUNDS = 8, SW_CHUNK = 32 * SW_CHUNK_ROUNDS;    // points per warp in the reduce

struct SweepGeom {
    int leaf, kmin, bits;   // key field = floor(v / leaf) - kmin, `bits` wide (as VoxelGeom)
    int idx_bits;           // low bits of the sort word hold the point index
    int off_bits;           // width of one in-voxel offset (0 .. leaf-1) in the gather word
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) { return *(const volatile uint32_t *)p; }
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) { *(volatile uint32_t *)p = v; }
__device__ __forceinline__ uint4 ld_volatile_v4(const uint32_t *p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_v4(uint32_t *p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- 1. sort words, gather words and all digit histograms -------------------------------
// pay[i] = R | G<<8 | B<<16 | (x - leaf*kx) << 24 | (y - leaf*ky) << (24+ob) | (z - leaf*kz) << (24+2ob):
// everything the reduce needs from a record in one aligned 8-byte load.
template <int BITS>
__global__ void __launch_bounds__(SW_KH_THREADS)
sw_keys_hist(const int16_t *__restrict__ rec, int n, SweepGeom g, int passes,
             uint64_t *__restrict__ words, uint64_t *__restrict__ pay, uint32_t *__restrict__ ghist) {
    constexpr int BINS = 1 << BITS;
    extern __shared__ uint32_t sw_hist[];    // [passes][BINS]
    for (int k = threadIdx.x; k < passes * BINS; k += SW_KH_THREADS) sw_hist[k] = 0;
    __syncthreads();
    const bool aligned = (((uintptr_t)rec) & 15) == 0;
    const int n_tiles = (n + SW_KH_TILE - 1) / SW_KH_TILE;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int first = tile * SW_KH_TILE + threadIdx.x * SW_KH_ITEMS;
        const int cnt = max(0, min(SW_KH_ITEMS, n - first));
        uint32_t w[20];     // 8 records = 40 halfwords
        if (cnt == SW_KH_ITEMS && aligned) {
            const uint4 *p = reinterpret_cast<const uint4 *>(rec + 5 * (size_t)first);
#pragma unroll
            for (int j = 0; j < 5; ++j) {
                const uint4 q = __ldg(p + j);
                w[4 * j] = q.x; w[4 * j + 1] = q.y; w[4 * j + 2] = q.z; w[4 * j + 3] = q.w;
            }
        } else {
            const size_t lim = 5 * (size_t)n, at = 5 * (size_t)first;
#pragma unroll
            for (int j = 0; j < 20; ++j) {
                const uint32_t lo = at + 2 * j < lim ? (uint16_t)rec[at + 2 * j] : 0u;
                const uint32_t hi = at + 2 * j + 1 < lim ? (uint16_t)rec[at + 2 * j + 1] : 0u;
                w[j] = lo | (hi << 16);
            }
        }
        uint64_t word[SW_KH_ITEMS], pw[SW_KH_ITEMS];
#pragma unroll
        for (int k = 0; k < SW_KH_ITEMS; ++k) {
            const int h = 5 * k;
            const int x = (int16_t)(w[h >> 1] >> (16 * (h & 1)));
            const int y = (int16_t)(w[(h + 1) >> 1] >> (16 * ((h + 1) & 1)));
            const int z = (int16_t)(w[(h + 2) >> 1] >> (16 * ((h + 2) & 1)));
            const uint32_t s3 = (w[(h + 3) >> 1] >> (16 * ((h + 3) & 1))) & 0xFFFFu;
            const uint32_t s4 = (w[(h + 4) >> 1] >> (16 * ((h + 4) & 1))) & 0xFFu;
            const int fx = floordiv_i(x, g.leaf), fy = floordiv_i(y, g.leaf), fz = floordiv_i(z, g.leaf);
            const uint64_t kx = (uint64_t)(fx - g.kmin), ky = (uint64_t)(fy - g.kmin), kz = (uint64_t)(fz - g.kmin);
            word[k] = (((kz << (2 * g.bits)) | (ky << g.bits) | kx) << g.idx_bits) | (uint64_t)(uint32_t)(first + k);
            const uint64_t ox = (uint64_t)(x - g.leaf * fx), oy = (uint64_t)(y - g.leaf * fy), oz = (uint64_t)(z - g.leaf * fz);
            pw[k] = (uint64_t)(s3 | (s4 << 16)) | (ox << 24) | (oy << (24 + g.off_bits)) | (oz << (24 + 2 * g.off_bits));
        }
        if (cnt == SW_KH_ITEMS) {
            uint4 *o = reinterpret_cast<uint4 *>(words + first), *q = reinterpret_cast<uint4 *>(pay + first);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[j] = make_uint4((uint32_t)word[2 * j], (uint32_t)(word[2 * j] >> 32),
                                  (uint32_t)word[2 * j + 1], (uint32_t)(word[2 * j + 1] >> 32));
                q[j] = make_uint4((uint32_t)pw[2 * j], (uint32_t)(pw[2 * j] >> 32),
                                  (uint32_t)pw[2 * j + 1], (uint32_t)(pw[2 * j + 1] >> 32));
            }
        } else {
#pragma unroll
            for (int k = 0; k < SW_KH_ITEMS; ++k)
                if (k < cnt) { words[first + k] = word[k]; pay[first + k] = pw[k]; }
        }
        // Neighbouring pixels mostly share their upper digits: count runs, not elements, and let the
        // lanes whose last run has the same digit add it once (a same-address shared-memory atomic
        // is serialised lane by lane).
        for (int p = 0; p < passes; ++p) {
            const int shift = g.idx_bits + p * BITS;
            uint32_t *h = sw_hist + p * BINS;
            uint32_t prev = cnt > 0 ? ((uint32_t)(word[0] >> shift) & (BINS - 1)) : (uint32_t)(BINS + (threadIdx.x & 31));
            uint32_t run = cnt > 0 ? 1u : 0u;
#pragma unroll
            for (int k = 1; k < SW_KH_ITEMS; ++k) {
                if (k < cnt) {
                    const uint32_t d = (uint32_t)(word[k] >> shift) & (BINS - 1);
                    if (d == prev) {
                        ++run;
                    } else {
                        atomicAdd(h + prev, run);
                        prev = d;
                        run = 1;
                    }
                }
            }
            const uint32_t peers = __match_any_sync(0xffffffffu, prev);
            const uint32_t tot = __reduce_add_sync(peers, run);
            if (cnt > 0 && (threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(h + prev, tot);
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < passes * BINS; k += SW_KH_THREADS) {
        const uint32_t v = sw_hist[k];
        if (v) atomicAdd(ghist + k, v);
    }
}

// ---- 2. bin bases: one block per pass ---------------------------------------------------
template <int BITS>
__global__ void __launch_bounds__(1 << BITS)
sw_hist_scan(uint32_t *__restrict__ ghist) {
    __shared__ uint32_t warp_tot[33];
    uint32_t *h = ghist + (size_t)blockIdx.x * (1 << BITS);
    const uint32_t v = h[threadIdx.x];
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, warp_tot, total);
    h[threadIdx.x] = ex;
}

// ---- 3. one digit pass ------------------------------------------------------------------
template <int BITS, int THREADS, int ITEMS>
struct SweepPassCfg {
    static constexpr int BINS = 1 << BITS, WARPS = THREADS / 32, TILE = THREADS * ITEMS, DPT = BINS / THREADS;
    static constexpr size_t SMEM = (size_t)TILE * 8 + (size_t)(WARPS + 3) * BINS * 4;
    static_assert(DPT >= 1 && DPT * THREADS == BINS, "every thread owns BINS / THREADS digits");
};

// Decoupled look-back for 32 digits at a time, one warp: lane = (group g = lane / 8, quad q = lane % 8)
// reads the status words of digits dq..dq+3 (dq = first digit + 4q) of predecessors
// wstart - (g*M + m), m < M, as 16-byte volatile loads -- a window of 4*M predecessor tiles per step,
// each tile's 32 status words one coalesced 128-byte row.  The window is folded nearest-first:
// counts add up until a tile that carries its full prefix; an unpublished word in front of that
// sends the whole warp back to poll the same window.  A thread-per-digit walk (one predecessor per
// L2 round trip) made the first wave of CTAs, which all start together, wait ~R/2 round trips.
template <int M>
__device__ __forceinline__ void sw_lookback(const uint32_t *status, int bins, uint32_t tile, int dq,
                                            uint32_t (&before)[4], uint32_t *err) {
    const int lane = threadIdx.x & 31, grp = lane >> 3;
    uint32_t sum[4] = {0, 0, 0, 0};
    uint32_t done = 0;                   // bit i: digit dq+i has reached a full prefix
    int wstart = (int)tile - 1;          // nearest predecessor of the current window
    int spins = 0;
    while (true) {
        uint4 s[M];
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const int t = wstart - (grp * M + m);
            s[m] = t >= 0 ? ld_volatile_v4(status + (size_t)t * bins + dq)
                          : make_uint4(SW_PREFIX, SW_PREFIX, SW_PREFIX, SW_PREFIX);   // before tile 0: nothing
        }
        uint32_t wsum[4] = {0, 0, 0, 0}, wdone = 0, winv = 0;
#pragma unroll
        for (int m = 0; m < M; ++m) {
            const uint32_t e[4] = {s[m].x, s[m].y, s[m].z, s[m].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t f = e[i] >> 30;
                if (!(((wdone | winv) >> i) & 1u)) {
                    if (f == 0) {
                        winv |= 1u << i;
                    } else {
                        wsum[i] += e[i] & SW_VALUE;
                        if (f == 2) wdone |= 1u << i;
                    }
                }
            }
        }
        // ordered fold across the four lane groups (the lower group holds the nearer tiles)
#pragma unroll
        for (int step = 8; step <= 16; step <<= 1) {
            const uint32_t mine = wdone | (winv << 4);
            const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, step);
            const bool lower = (lane & step) == 0;
            const uint32_t af = lower ? mine : other, bf = lower ? other : mine;
            uint32_t nd = 0, ni = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint32_t o = __shfl_xor_sync(0xffffffffu, wsum[i], step);
                const uint32_t as = lower ? wsum[i] : o, bs = lower ? o : wsum[i];
                const bool a_closed = ((af >> i) | (af >> (4 + i))) & 1u;
                wsum[i] = a_closed ? as : as + bs;
                const uint32_t src = a_closed ? af : bf;
                nd |= ((src >> i) & 1u) << i;
                ni |= ((src >> (4 + i)) & 1u) << i;
            }
            wdone = nd;
            winv = ni;
        }
        if (__any_sync(0xffffffffu, (winv & ~done) != 0)) {
            if (++spins >= SW_SPIN_LIMIT) {
                if (lane == 0) atomicExch(err, 1u);
                break;
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (!((done >> i) & 1u)) sum[i] += wsum[i];
        done |= wdone;
        if (__all_sync(0xffffffffu, done == 0xFu)) break;
        wstart -= 4 * M;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) before[i] = sum[i];
}

template <int BITS, int THREADS, int ITEMS>
__global__ void __launch_bounds__(THREADS, (BITS <= 8 ? 3 : 2))
sw_pass(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, int n, int shift,
        const uint32_t *__restrict__ gbase, uint32_t *status, uint32_t *ticket, uint32_t *err) {
    using Cfg = SweepPassCfg<BITS, THREADS, ITEMS>;
    constexpr int BINS = Cfg::BINS, WARPS = Cfg::WARPS, TILE = Cfg::TILE, DPT = Cfg::DPT;
    extern __shared__ __align__(16) uint8_t sw_smem[];
    uint64_t *skeys = reinterpret_cast<uint64_t *>(sw_smem);     // [TILE] the tile in digit order
    uint32_t *wh = reinterpret_cast<uint32_t *>(skeys + TILE);   // [WARPS][BINS] per-warp digit counters
    uint32_t *tile_excl = wh + WARPS * BINS;                      // [BINS] first tile-local slot of a digit
    uint32_t *tile_cnt = tile_excl + BINS;                        // [BINS] elements of a digit in this tile
    uint32_t *gdelta = tile_cnt + BINS;                           // [BINS] global slot = gdelta + tile-local slot
    __shared__ uint32_t s_tile, warp_tot[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);   // tiles are taken in the order CTAs start running
    for (int k = threadIdx.x; k < WARPS * BINS; k += THREADS) wh[k] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const int tbase = (int)tile * TILE;
    const int wbase = tbase + warp * (32 * ITEMS);   // warp-blocked: warp w owns 32*ITEMS consecutive words
    uint32_t *mywh = wh + warp * BINS;
    uint64_t key[ITEMS];
    uint32_t local[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int i = wbase + k * 32 + lane;
        key[k] = i < n ? in[i] : ~0ull;
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const bool valid = wbase + k * 32 + lane < n;
        const uint32_t d = valid ? ((uint32_t)(key[k] >> shift) & (BINS - 1)) : (uint32_t)(BINS + lane);
        const uint32_t mask = __match_any_sync(0xffffffffu, d);
        const int leader = __ffs(mask) - 1;
        const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
        uint32_t prev = 0;
        if (valid && lane == leader) {
            prev = mywh[d];
            mywh[d] = prev + __popc(mask);
        }
        __syncwarp();
        prev = __shfl_sync(0xffffffffu, prev, leader);
        local[k] = prev + rank;
    }
    __syncthreads();
    // digit totals of the tile; wh becomes the exclusive prefix over warps
    uint32_t cnt[DPT], tsum = 0;
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const int d = threadIdx.x * DPT + j;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < WARPS; ++w) {
            const uint32_t c = wh[w * BINS + d];
            wh[w * BINS + d] = run;
            run += c;
        }
        cnt[j] = run;
        tsum += run;
        tile_cnt[d] = run;
        // let the successors see this tile's counts as early as possible
        st_volatile_u32(status + (size_t)tile * BINS + d, (tile == 0 ? SW_PREFIX : SW_AGG) | run);
    }
    uint32_t total;
    uint32_t ex = block_exclusive_scan(tsum, warp_tot, total);
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        tile_excl[threadIdx.x * DPT + j] = ex;
        ex += cnt[j];
    }
    __syncthreads();
    // reorder the tile in shared memory so that the global stores are runs of consecutive words
    // (before the look-back: the words leave the registers while the predecessors finish)
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (wbase + k * 32 + lane < n) {
            const uint32_t d = (uint32_t)(key[k] >> shift) & (BINS - 1);
            skeys[tile_excl[d] + mywh[d] + local[k]] = key[k];
        }
    }
    for (int group = warp; group < BINS / 32; group += WARPS) {
        const int dq = group * 32 + 4 * (lane & 7);
        uint32_t before[4] = {0, 0, 0, 0};
        if (tile > 0) sw_lookback<4>(status, BINS, tile, dq, before, err);
        if (lane < 8) {
            uint32_t pub[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                gdelta[dq + i] = gbase[dq + i] + before[i] - tile_excl[dq + i];
                pub[i] = SW_PREFIX | ((before[i] + tile_cnt[dq + i]) & SW_VALUE);
            }
            if (tile > 0) st_volatile_v4(status + (size_t)tile * BINS + dq, make_uint4(pub[0], pub[1], pub[2], pub[3]));
        }
    }
    __syncthreads();
    const int nvalid = min(TILE, n - tbase);
    for (int j = threadIdx.x; j < nvalid; j += THREADS) {
        const uint64_t k = skeys[j];
        const uint32_t d = (uint32_t)(k >> shift) & (BINS - 1);
        out[gdelta[d] + (uint32_t)j] = k;
    }
}

// ---- 4. voxel boundaries per chunk ------------------------------------------------------
// info[c] = (position + 1 of the last head in chunk c, 0 if none) << 32 | heads in chunk c
__global__ void __launch_bounds__(256)
sw_chunk_heads(const uint64_t *__restrict__ sorted, int n, int idx_bits, int n_chunks, uint64_t *__restrict__ info) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n_chunks) return;
    const int base = c * SW_CHUNK;
    uint64_t before = base > 0 ? sorted[base - 1] >> idx_bits : 0;
    uint32_t heads = 0, lastpos = 0;
#pragma unroll
    for (int r = 0; r < SW_CHUNK_ROUNDS; ++r) {
        const int e = base + r * 32 + lane;
        const bool valid = e < n;
        const uint64_t key = valid ? sorted[e] >> idx_bits : 0;
        uint64_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        if (lane == 0) prev = before;
        const bool head = valid && (e == 0 || key != prev);
        const uint32_t hm = __ballot_sync(0xffffffffu, head);
        heads += __popc(hm);
        if (hm) lastpos = (uint32_t)(base + r * 32 + (31 - __clz(hm)) + 1);
        before = __shfl_sync(0xffffffffu, key, 31);
    }
    if (lane == 0) info[c] = ((uint64_t)lastpos << 32) | heads;
}

// single block: chunk_off = exclusive sum of the head counts, ownerpos = exclusive max of lastpos.
// Batches of 4096 chunks, coalesced 32-byte loads per thread, the next batch in flight during the scan.
__global__ void __launch_bounds__(1024)
sw_chunk_scan(const uint64_t *__restrict__ info, int n_chunks, uint32_t *__restrict__ chunk_off,
              uint32_t *__restrict__ ownerpos, int32_t *__restrict__ nv_out) {
    constexpr int IPT = 4;
    __shared__ uint32_t ssum[33], smax[33];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t csum = 0, cmax = 0;     // carried over the batches (block-uniform)
    uint64_t v[IPT], nxt[IPT];
#pragma unroll
    for (int i = 0; i < IPT; ++i) {
        const int c = (int)threadIdx.x * IPT + i;
        nxt[i] = c < n_chunks ? info[c] : 0;
    }
    for (int base = 0; base < n_chunks; base += 1024 * IPT) {
        const int first = base + (int)threadIdx.x * IPT;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            v[i] = nxt[i];
            const int c = first + 1024 * IPT + i;
            nxt[i] = c < n_chunks ? info[c] : 0;
        }
        uint32_t sum = 0, mx = 0;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            sum += (uint32_t)v[i];
            mx = max(mx, (uint32_t)(v[i] >> 32));
        }
        uint32_t isum = sum, imx = mx;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t a = __shfl_up_sync(0xffffffffu, isum, d), b = __shfl_up_sync(0xffffffffu, imx, d);
            if (lane >= d) { isum += a; imx = max(imx, b); }
        }
        uint32_t emx_in_warp = __shfl_up_sync(0xffffffffu, imx, 1);
        if (lane == 0) emx_in_warp = 0;
        if (lane == 31) { ssum[warp] = isum; smax[warp] = imx; }
        __syncthreads();
        if (warp == 0) {
            const uint32_t a = ssum[lane], b = smax[lane];
            uint32_t ia = a, ib = b;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t ta = __shfl_up_sync(0xffffffffu, ia, d), tb = __shfl_up_sync(0xffffffffu, ib, d);
                if (lane >= d) { ia += ta; ib = max(ib, tb); }
            }
            uint32_t eb = __shfl_up_sync(0xffffffffu, ib, 1);
            if (lane == 0) eb = 0;
            ssum[lane] = ia - a;
            smax[lane] = eb;
            if (lane == 31) { ssum[32] = ia; smax[32] = ib; }
        }
        __syncthreads();
        uint32_t esum = csum + ssum[warp] + (isum - sum);
        uint32_t emx = max(cmax, max(smax[warp], emx_in_warp));
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            if (first + i < n_chunks) {
                chunk_off[first + i] = esum;
                ownerpos[first + i] = emx;
            }
            esum += (uint32_t)v[i];
            emx = max(emx, (uint32_t)(v[i] >> 32));
        }
        csum += ssum[32];
        cmax = max(cmax, smax[32]);
        __syncthreads();     // ssum / smax are rewritten by the next batch
    }
    if (threadIdx.x == 0) *nv_out = (int32_t)csum;
}

// ---- 5. segmented sums -> records -------------------------------------------------------
__device__ __forceinline__ void sw_emit(int16_t *__restrict__ out, uint32_t vid, const uint32_t (&a)[7], uint64_t key,
                                        const SweepGeom &g) {
    const uint64_t m = (1ull << g.bits) - 1;
    const int kx = (int)(key & m) + g.kmin, ky = (int)((key >> g.bits) & m) + g.kmin,
              kz = (int)((key >> (2 * g.bits)) & m) + g.kmin;
    const uint32_t cnt = a[6];
    int16_t *o = out + 5 * (size_t)vid;
    o[0] = (int16_t)(g.leaf * kx + (int)(a[0] / cnt));
    o[1] = (int16_t)(g.leaf * ky + (int)(a[1] / cnt));
    o[2] = (int16_t)(g.leaf * kz + (int)(a[2] / cnt));
    o[3] = (int16_t)((a[3] / cnt) | ((a[4] / cnt) << 8));
    o[4] = (int16_t)(a[5] / cnt);
}

// slots[c] = 8 words: 7 sums of chunk c's last voxel when it continues into later chunks, word 7 =
// "finalise me"; slotkey[c] = that voxel's key.  All the chunk's sort words and gather words are
// requested before the first round is reduced (a round is one dependent DRAM round trip otherwise).
__global__ void __launch_bounds__(256)
sw_reduce(const uint64_t *__restrict__ sorted, int n, const uint64_t *__restrict__ pay, SweepGeom g, int n_chunks,
          const uint32_t *__restrict__ chunk_off, const uint32_t *__restrict__ ownerpos,
          uint32_t *__restrict__ slots, uint64_t *__restrict__ slotkey, int16_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= n_chunks) return;
    const int base = c * SW_CHUNK;
    const uint64_t idx_mask = (1ull << g.idx_bits) - 1;
    const uint32_t om = (1u << g.off_bits) - 1;
    uint64_t word[SW_CHUNK_ROUNDS], pw[SW_CHUNK_ROUNDS];
#pragma unroll
    for (int r = 0; r < SW_CHUNK_ROUNDS; ++r) {
        const int e = base + r * 32 + lane;
        word[r] = e < n ? sorted[e] : 0;
    }
#pragma unroll
    for (int r = 0; r < SW_CHUNK_ROUNDS; ++r) {
        const int e = base + r * 32 + lane;
        pw[r] = e < n ? __ldg(pay + (word[r] & idx_mask)) : 0;
    }
    uint32_t run_vid = chunk_off[c] - 1u;     // voxel of the last point before this round
    const uint32_t opos = ownerpos[c];
    const uint32_t owner = opos ? (opos - 1u) / SW_CHUNK : 0u;   // chunk holding the head of the voxel that runs into this one
    uint64_t before = base > 0 ? sorted[base - 1] >> g.idx_bits : 0;
    // the open voxel carried from round to round (warp-uniform)
    uint32_t carry[7] = {0, 0, 0, 0, 0, 0, 0};
    bool carry_has_head = false;
    uint64_t carry_key = before;
#pragma unroll
    for (int r = 0; r < SW_CHUNK_ROUNDS; ++r) {
        const int e = base + r * 32 + lane;
        const bool valid = e < n;
        const uint64_t key = word[r] >> g.idx_bits;
        uint64_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        if (lane == 0) prev = before;
        const bool head = valid && (e == 0 || key != prev);
        const uint32_t hm = __ballot_sync(0xffffffffu, head);
        // the carried voxel ended exactly at the previous round's last lane
        if ((hm & 1u) && r > 0 && lane == 0) {
            if (carry_has_head) {
                sw_emit(out, run_vid, carry, carry_key, g);
            } else {
#pragma unroll
                for (int k = 0; k < 7; ++k) atomicAdd(slots + 8 * (size_t)owner + k, carry[k]);
            }
        }
        const uint32_t le = hm & (0xffffffffu >> (31 - lane));
        const uint32_t vid = run_vid + __popc(le);
        const bool started_here = le != 0;
        const int seg_start = started_here ? 31 - __clz(le) : 0;
        uint32_t v[7] = {0, 0, 0, 0, 0, 0, 0};
        if (valid) {
            const uint32_t lo = (uint32_t)pw[r];
            const uint64_t offs = pw[r] >> 24;
            v[0] = (uint32_t)offs & om;
            v[1] = (uint32_t)(offs >> g.off_bits) & om;
            v[2] = (uint32_t)(offs >> (2 * g.off_bits)) & om;
            v[3] = lo & 0xFF; v[4] = (lo >> 8) & 0xFF; v[5] = (lo >> 16) & 0xFF; v[6] = 1;
        }
        // warp-shuffle segmented inclusive scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t o[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) o[k] = __shfl_up_sync(0xffffffffu, v[k], d);
            if (lane >= d && lane - d >= seg_start) {
#pragma unroll
                for (int k = 0; k < 7; ++k) v[k] += o[k];
            }
        }
        // key of the voxel a lane belongs to (its own key, except on the padding lanes past n)
        const uint64_t key_at_start = __shfl_sync(0xffffffffu, key, seg_start);
        const uint64_t seg_key = started_here ? key_at_start : carry_key;
        const bool is_end = lane < 31 && ((hm >> (lane + 1)) & 1u);
        if (is_end) {
            if (started_here) {
                sw_emit(out, vid, v, seg_key, g);
            } else {
                uint32_t tot[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) tot[k] = carry[k] + v[k];
                if (carry_has_head) {
                    sw_emit(out, vid, tot, seg_key, g);
                } else {
#pragma unroll
                    for (int k = 0; k < 7; ++k) atomicAdd(slots + 8 * (size_t)owner + k, tot[k]);
                }
            }
        }
        // lane 31's voxel stays open
        const bool l31_started = __shfl_sync(0xffffffffu, (int)started_here, 31) != 0;
        const uint64_t l31_key = __shfl_sync(0xffffffffu, seg_key, 31);
#pragma unroll
        for (int k = 0; k < 7; ++k) {
            const uint32_t t = __shfl_sync(0xffffffffu, v[k], 31);
            carry[k] = l31_started ? t : carry[k] + t;
        }
        if (l31_started) {
            carry_has_head = true;
            carry_key = l31_key;
        }
        run_vid += __popc(hm);
        before = __shfl_sync(0xffffffffu, key, 31);
    }
    // the chunk's last voxel: complete if the next chunk starts with a head (or there is none)
    if (lane == 0) {
        const bool complete = base + SW_CHUNK >= n || (sorted[base + SW_CHUNK] >> g.idx_bits) != carry_key;
        if (carry_has_head && complete) {
            sw_emit(out, run_vid, carry, carry_key, g);
        } else {
            const size_t slot = carry_has_head ? (size_t)c : (size_t)owner;
#pragma unroll
            for (int k = 0; k < 7; ++k) atomicAdd(slots + 8 * slot + k, carry[k]);
            if (carry_has_head) {
                slots[8 * slot + 7] = 1u;
                slotkey[c] = carry_key;
            }
        }
    }
}

// ---- 6. the voxels that span chunks -----------------------------------------------------
__global__ void __launch_bounds__(256)
sw_finalize_open(const uint64_t *__restrict__ info, const uint32_t *__restrict__ chunk_off,
                 const uint32_t *__restrict__ slots, const uint64_t *__restrict__ slotkey, int n_chunks,
                 SweepGeom g, int16_t *__restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks || slots[8 * (size_t)c + 7] == 0) return;
    uint32_t a[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) a[k] = slots[8 * (size_t)c + k];
    sw_emit(out, chunk_off[c] + (uint32_t)info[c] - 1u, a, slotkey[c], g);
}

// ---- host side ----------------------------------------------------------------------------
inline int sweep_idx_bits(int n) {
    int b = 1;
    while (b < 31 && (1ll << b) < (long long)n) ++b;
    return b;
}

inline bool sweep_fits(int n, int leaf) {
    const int kmin = -(32768 + leaf - 1) / leaf, kmax = 32767 / leaf;
    int bits = 1;
    while ((1 << bits) < (kmax - kmin + 1)) ++bits;
    int ob = 1;
    while ((1 << ob) < leaf) ++ob;
    return 3 * bits + sweep_idx_bits(n) <= 64 && 24 + 3 * ob <= 64 && n < (1 << 30);
}

template <int BITS, int THREADS, int ITEMS>
inline int sweep_configure() {
    using Cfg = SweepPassCfg<BITS, THREADS, ITEMS>;
    if (cudaFuncSetAttribute(sw_pass<BITS, THREADS, ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Cfg::SMEM) != cudaSuccess) return -2;
    return 0;
}

// Returns the voxel count (>= 0), -2 on a CUDA error, -3 on allocation failure, -4 when the sort
// word does not fit (the caller picks the pair sort then).  Everything is queued before the one
// synchronisation that brings the count back.
template <int BITS, int THREADS, int ITEMS>
inline int voxel_merge_sweep(VoxelScratch &s, const int16_t *rec, int n, int leaf, int16_t *out, cudaStream_t cs,
                             int sm_count) {
    using Cfg = SweepPassCfg<BITS, THREADS, ITEMS>;
    constexpr int BINS = Cfg::BINS;
    if (!sweep_fits(n, leaf)) return -4;
    SweepGeom g;
    g.leaf = leaf;
    g.kmin = -(32768 + leaf - 1) / leaf;
    const int kmax = 32767 / leaf;
    g.bits = 1;
    while ((1 << g.bits) < (kmax - g.kmin + 1)) ++g.bits;
    g.idx_bits = sweep_idx_bits(n);
    g.off_bits = 1;
    while ((1 << g.off_bits) < leaf) ++g.off_bits;
    const int passes = (3 * g.bits + BITS - 1) / BITS;
    const int n_tiles = (n + Cfg::TILE - 1) / Cfg::TILE;
    const int n_chunks = (n + SW_CHUNK - 1) / SW_CHUNK;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // [words0][words1][pay] | zeroed: [ghist][tickets][err][status][slots] | [info][chunk_off][ownerpos][slotkey][nv]
    const size_t o_w0 = 0, o_w1 = o_w0 + al((size_t)n * 8), o_pay = o_w1 + al((size_t)n * 8),
                 o_zero = o_pay + al((size_t)n * 8),
                 o_ghist = o_zero, o_ticket = o_ghist + al((size_t)passes * BINS * 4), o_err = o_ticket + al(64),
                 o_status = o_err + al(64), o_slots = o_status + al((size_t)passes * n_tiles * BINS * 4),
                 o_zero_end = o_slots + al((size_t)n_chunks * 32),
                 o_info = o_zero_end, o_off = o_info + al((size_t)n_chunks * 8), o_owner = o_off + al((size_t)n_chunks * 4),
                 o_skey = o_owner + al((size_t)n_chunks * 4), o_nv = o_skey + al((size_t)n_chunks * 8),
                 total = o_nv + al(64);
    if (total > s.cap) {
        cudaFree(s.buf);
        s.buf = nullptr; s.cap = 0;
        if (cudaMalloc(&s.buf, total) != cudaSuccess) { cudaGetLastError(); return -3; }
        s.cap = total;
    }
    if (!s.h_count && cudaHostAlloc(&s.h_count, 64, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return -3; }
    uint64_t *w0 = (uint64_t *)(s.buf + o_w0), *w1 = (uint64_t *)(s.buf + o_w1), *pay = (uint64_t *)(s.buf + o_pay);
    uint32_t *ghist = (uint32_t *)(s.buf + o_ghist), *ticket = (uint32_t *)(s.buf + o_ticket),
             *err = (uint32_t *)(s.buf + o_err), *status = (uint32_t *)(s.buf + o_status),
             *slots = (uint32_t *)(s.buf + o_slots), *chunk_off = (uint32_t *)(s.buf + o_off),
             *ownerpos = (uint32_t *)(s.buf + o_owner);
    uint64_t *info = (uint64_t *)(s.buf + o_info), *slotkey = (uint64_t *)(s.buf + o_skey);
    int32_t *nv_dev = (int32_t *)(s.buf + o_nv);
    if (cudaMemsetAsync(s.buf + o_zero, 0, o_zero_end - o_zero, cs) != cudaSuccess) return -2;
    const int kh_tiles = (n + SW_KH_TILE - 1) / SW_KH_TILE;
    const int kh_grid = std::min(kh_tiles, std::max(1, sm_count) * 8);
    sw_keys_hist<BITS><<<kh_grid, SW_KH_THREADS, (size_t)passes * BINS * 4, cs>>>(rec, n, g, passes, w0, pay, ghist);
    sw_hist_scan<BITS><<<passes, BINS, 0, cs>>>(ghist);
    for (int p = 0; p < passes; ++p) {
        sw_pass<BITS, THREADS, ITEMS><<<n_tiles, THREADS, Cfg::SMEM, cs>>>(
            w0, w1, n, g.idx_bits + p * BITS, ghist + (size_t)p * BINS, status + (size_t)p * n_tiles * BINS,
            ticket + p, err);
        std::swap(w0, w1);
    }
    const int cblocks = (n_chunks + 7) / 8;
    sw_chunk_heads<<<cblocks, 256, 0, cs>>>(w0, n, g.idx_bits, n_chunks, info);
    sw_chunk_scan<<<1, 1024, 0, cs>>>(info, n_chunks, chunk_off, ownerpos, nv_dev);
    sw_reduce<<<cblocks, 256, 0, cs>>>(w0, n, pay, g, n_chunks, chunk_off, ownerpos, slots, slotkey, out);
    sw_finalize_open<<<(n_chunks + 255) / 256, 256, 0, cs>>>(info, chunk_off, slots, slotkey, n_chunks, g, out);
    if (cudaMemcpyAsync(s.h_count, nv_dev, 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return -2;
    if (cudaMemcpyAsync(s.h_count + 1, err, 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(cs) != cudaSuccess) return -2;
    if (cudaGetLastError() != cudaSuccess) return -2;
    if (s.h_count[1] != 0) return -2;           // a look-back gave up waiting
    const int nv = s.h_count[0];
    if (nv < 1 || nv > n) return -2;
    return nv;
}

}  // namespace pcs
