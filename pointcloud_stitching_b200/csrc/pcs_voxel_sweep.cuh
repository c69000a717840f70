// K3, second generation: voxel-grid merge with a one-sweep radix sort (oracle/SPEC.md s3 --
// own integer spec; the reference only #includes pcl/filters/voxel_grid.h,
// src/pcs-multicamera-optimized.cpp:17).
//
//   0. sw_bounds      occupied box of the cloud in voxel units (and, for the slab plan, a histogram
//                     over z planes); the box fixes the key layout, so it is read back by the host
//   1. sw_keys_hist   one read of the records: word[slot] = key << idx_bits | slot with the mixed-radix
//                     key ((kz-z0)*dy + (ky-y0))*dx + (kx-x0) (one u64 carries the voxel key AND the
//                     point's slot), pay[slot] = colour + in-voxel offsets (one u64), plus the digit
//                     histograms of every pass; with a z-slab filter only the slab's points get slots
//   2. sw_hist_scan   exclusive scan of each pass's histogram -> global bin bases
//   3. sw_pass x P    one kernel per 8-bit digit: each 4096-word tile arrives by one TMA bulk copy, is
//                     ranked (per-bit ballots + per-warp counters), reordered in shared memory and written
//                     once; tile order from a ticket, the cross-tile prefix of every digit from a two-level
//                     decoupled look-back (16 B/pt per pass)
//   4. sw_chunk_heads / sw_chunk_scan   voxel (segment) boundaries per 256-point chunk and their scan
//   5. sw_reduce      one warp per chunk: gather pay[] through the sorted slots, warp-shuffle
//                     segmented sums, integer means written straight to the output; only the segments
//                     that cross a chunk boundary go through atomics into a per-chunk slot
//   6. sw_finalize_open   integer means of those slots
//
// The sums are integers, so the result equals the CPU restatement bit for bit whatever the order
// of the additions.  Needs key bits + slot bits <= 64 and 24 + 3*ceil(log2 leaf) <= 64; the caller
// falls back to the (key, idx) pair sort of pcs_voxel.cuh otherwise.  Measurements and the tuning
// history: profiles/r01_voxel.md.
#pragma once
#include <vector>

#include "pcs_voxel.cuh"

namespace pcs {

constexpr uint32_t SW_AGG = 1u << 30, SW_PREFIX = 2u << 30, SW_VALUE = (1u << 30) - 1;
constexpr int SW_SPIN_LIMIT = 1 << 20;      // ~0.5 s of polling: a lost predecessor becomes an error, not a hang
constexpr int SW_KH_THREADS = 256, SW_KH_ITEMS = 8, SW_KH_TILE = SW_KH_THREADS * SW_KH_ITEMS;
constexpr int SW_GROUP = 16;                // tiles per look-back group (= one look-back window)
constexpr int SW_CHUNK_ROUNDS = 8, SW_CHUNK = 32 * SW_CHUNK_ROUNDS;    // points per warp in the reduce

// Key geometry.  With K = ceil(32768 / leaf) every coordinate v becomes u = v + leaf*K >= 0 and
// q = floor(u / leaf) = floor(v / leaf) + K  (one multiply-high: u < 2^17, magic = floor(2^40 / leaf) + 1
// is exact for u * leaf < 2^40).  A first pass over the records finds the occupied box
// [x0, x1] x [y0, y1] x [z0, z1] in q units -- what PCL's VoxelGrid does with getMinMax3D before it
// builds its indices -- so the sort key only spends the bits the scene needs:
//     key = ((qz - z0) * dy + (qy - y0)) * dx + (qx - x0)        dx, dy = box extents (mixed radix)
// A z-slab filter [z_lo, z_hi) (q units) restricts the merge to a horizontal slice of the grid: the
// slices of disjoint slabs, concatenated in slab order, are the full result (multi-GPU sharding).
struct SweepGeom {
    int leaf, K, bias;            // bias = leaf * K
    unsigned long long magic;     // floor(2^40 / leaf) + 1
    int x0, y0, z0;               // box origin, q units
    unsigned dx, dy;              // box extents along x and y (the key is mixed radix)
    unsigned long long mdx, mdy;  // floor(2^64 / d) + 1 for d = dx, dy (0 when d == 1): key / d = umul64hi(key, m)
    int idx_bits;                 // low bits of the sort word hold the point's slot
    int off_bits;                 // width of one in-voxel offset (0 .. leaf-1) in the gather word
    int z_lo, z_hi;               // keep qz in [z_lo, z_hi)
};

// Everything the sort needs that depends on the data (the occupied box fixes the key layout): written by
// sw_plan ON THE DEVICE and copied into a constant-memory slot, so that the merge needs no host round trip
// and the kernels still read their geometry as constant-bank operands.  Launch grids are sized on the
// host from upper bounds (n points, the widest possible key); blocks past the plan's counts exit.
struct SweepPlan {
    SweepGeom g;
    int m;            // points to sort (== n without a slab filter)
    int passes, n_tiles, n_chunks;
    int status;       // 0, or -4 when the (key, slot) word does not fit 64 bits
    int pre;          // 1: runs of neighbouring same-voxel points are merged into one sort element (m counts runs)
    int pad[2];
};
constexpr int SW_PLAN_SLOTS = 8;     // one per context (contexts sharing a slot must not merge concurrently)
__constant__ SweepPlan c_sweep_plan[SW_PLAN_SLOTS];

__device__ __forceinline__ uint32_t sw_q(int v, const SweepGeom &g) {
    return (uint32_t)(((unsigned long long)(uint32_t)(v + g.bias) * g.magic) >> 40);
}

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

// eight consecutive records (40 halfwords) as 20 words; past-the-end halfwords read as 0
__device__ __forceinline__ void sw_load8(const int16_t *__restrict__ rec, int n, int first, bool fast, uint32_t (&w)[20]) {
    if (fast) {
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
}
__device__ __forceinline__ int sw_half(const uint32_t (&w)[20], int h) {   // halfword h, sign-extended
    return (int16_t)(w[h >> 1] >> (16 * (h & 1)));
}

// ---- 0. occupied box, slab population, optional histogram over qz ---------------------------
// bounds[0..2] = min qx, qy, qz; bounds[3..5] = max; bounds[6] = points inside the slab; bounds[7] = runs (see below).
// zhist (HIST): points per qz plane over ALL points (the plan for cutting slabs).
template <bool HIST>
__global__ void __launch_bounds__(SW_KH_THREADS)
sw_bounds(const int16_t *__restrict__ rec, int n, SweepGeom g, uint32_t *__restrict__ bounds,
          uint32_t *__restrict__ zhist, int zbins, int zhist_in_smem, const int32_t *__restrict__ n_dev = nullptr) {
    extern __shared__ uint32_t sw_zh[];
    if (n_dev) n = max(0, min(n, *n_dev));       // the point count lives on the device (an inbox filled by peers)
    __shared__ uint32_t red[8];
    if (threadIdx.x < 8) red[threadIdx.x] = threadIdx.x < 3 ? 0xFFFFFFFFu : 0u;
    uint32_t *zh = (HIST && zhist_in_smem) ? sw_zh : zhist;
    if (HIST && zhist_in_smem)
        for (int k = threadIdx.x; k < zbins; k += SW_KH_THREADS) sw_zh[k] = 0;
    __syncthreads();
    const bool aligned = (((uintptr_t)rec) & 15) == 0;
    const int n_tiles = (n + SW_KH_TILE - 1) / SW_KH_TILE;
    uint32_t mn[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, mx[3] = {0, 0, 0}, inside = 0, heads = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int first = tile * SW_KH_TILE + threadIdx.x * SW_KH_ITEMS;
        const int cnt = max(0, min(SW_KH_ITEMS, n - first));
        if (cnt == 0) continue;
        uint32_t w[20];
        sw_load8(rec, n, first, cnt == SW_KH_ITEMS && aligned, w);
        uint32_t prev = 0, run = 0;
        uint32_t lx = 0, ly = 0, lz = 0;
        bool lin = false;        // the previous item of this thread lies inside the slab (and its voxel is lx, ly, lz)
#pragma unroll
        for (int k = 0; k < SW_KH_ITEMS; ++k) {
            if (k < cnt) {
                const uint32_t qx = sw_q(sw_half(w, 5 * k), g), qy = sw_q(sw_half(w, 5 * k + 1), g),
                               qz = sw_q(sw_half(w, 5 * k + 2), g);
                const bool in = (int)qz >= g.z_lo && (int)qz < g.z_hi;
                if (in) {
                    mn[0] = min(mn[0], qx); mn[1] = min(mn[1], qy); mn[2] = min(mn[2], qz);
                    mx[0] = max(mx[0], qx); mx[1] = max(mx[1], qy); mx[2] = max(mx[2], qz);
                    ++inside;
                    // a run = consecutive items of one thread in the same voxel (neighbouring pixels of a surface):
                    // sw_keys_hist can merge a run into one sort element
                    if (!(lin && qx == lx && qy == ly && qz == lz)) ++heads;
                }
                lin = in; lx = qx; ly = qy; lz = qz;
                if (HIST) {
                    if (run && qz == prev) {
                        ++run;
                    } else {
                        if (run) atomicAdd(zh + prev, run);
                        prev = qz;
                        run = 1;
                    }
                }
            }
        }
        if (HIST && run) atomicAdd(zh + prev, run);
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        mn[a] = __reduce_min_sync(0xffffffffu, mn[a]);
        mx[a] = __reduce_max_sync(0xffffffffu, mx[a]);
    }
    inside = __reduce_add_sync(0xffffffffu, inside);
    heads = __reduce_add_sync(0xffffffffu, heads);
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(red + a, mn[a]);
            atomicMax(red + 3 + a, mx[a]);
        }
        atomicAdd(red + 6, inside);
        atomicAdd(red + 7, heads);
    }
    __syncthreads();
    if (threadIdx.x < 3) atomicMin(bounds + threadIdx.x, red[threadIdx.x]);
    else if (threadIdx.x < 6) atomicMax(bounds + threadIdx.x, red[threadIdx.x]);
    else if (threadIdx.x < 8 && red[threadIdx.x]) atomicAdd(bounds + threadIdx.x, red[threadIdx.x]);
    if (HIST && zhist_in_smem)
        for (int k = threadIdx.x; k < zbins; k += SW_KH_THREADS) {
            const uint32_t v = sw_zh[k];
            if (v) atomicAdd(zhist + k, v);
        }
}

// ---- 0b. the plan -----------------------------------------------------------------------------
__device__ __forceinline__ int sw_bits_for_dev(unsigned long long count) {     // bits for values 0 .. count-1 (>= 1)
    int b = 1;
    while (b < 63 && (1ull << b) < count) ++b;
    return b;
}
__global__ void sw_plan(const uint32_t *__restrict__ bounds, SweepGeom base, int n, int slab, int tile, int bits,
                        int prereduce, SweepPlan *__restrict__ plan, int32_t *__restrict__ nv_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    SweepPlan p;
    p.g = base;
    p.pad[0] = p.pad[1] = 0;
    p.status = 0;
    p.pre = 0;
    p.m = (int)bounds[6];
    // Pre-reduction: raster neighbours on a surface share a voxel; when merging the runs inside every thread's eight
    // records drops at least 10 % of the elements, sort the runs (their partial sums ride in the gather word: count 4
    // bits, three colour sums of 11 bits, three offset sums of off_bits + 3 bits <= 64 for leaf <= 64 mm)
    const int heads = (int)bounds[7];
    if (prereduce && p.m > 0 && p.m <= n && heads > 0 && heads <= p.m && base.off_bits <= 6 &&
        (long long)heads * 10 <= (long long)p.m * 9) {
        p.pre = 1;
        p.m = heads;
    }
    p.passes = p.n_tiles = p.n_chunks = 0;
    if (p.m > 0 && p.m <= n) {
        SweepGeom &g = p.g;
        g.x0 = (int)bounds[0]; g.y0 = (int)bounds[1]; g.z0 = (int)bounds[2];
        g.dx = bounds[3] - bounds[0] + 1;
        g.dy = bounds[4] - bounds[1] + 1;
        g.mdx = g.dx > 1 ? ~0ull / g.dx + 1ull : 0ull;
        g.mdy = g.dy > 1 ? ~0ull / g.dy + 1ull : 0ull;
        const unsigned long long dz = (unsigned long long)bounds[5] - bounds[2] + 1;
        g.idx_bits = sw_bits_for_dev((unsigned long long)((slab || p.pre) ? p.m : n));
        const int key_bits = sw_bits_for_dev((unsigned long long)g.dx * g.dy * dz);   // < 2^48 always
        // the multiply-high divisions in sw_emit are exact while key * d stays below 2^64
        const double vol = (double)g.dx * (double)g.dy * (double)dz * (double)max(g.dx, g.dy);
        if (key_bits + g.idx_bits > 64 || vol >= 9.0e18) {
            p.status = -4;
        } else {
            p.passes = (key_bits + bits - 1) / bits;
            p.n_tiles = (p.m + tile - 1) / tile;
            p.n_chunks = (p.m + SW_CHUNK - 1) / SW_CHUNK;
        }
    } else if (p.m > n) {
        p.status = -2;
    }
    if (p.status != 0) p.pre = 0;
    *plan = p;
    if (p.status != 0 || p.m == 0) *nv_out = p.status;     // nothing else will write the result
}

// ---- 1. sort words, gather words and all digit histograms -------------------------------
// pay[slot] = R | G<<8 | B<<16 | (x - leaf*kx) << 24 | (y - leaf*ky) << (24+ob) | (z - leaf*kz) << (24+2ob):
// everything the reduce needs from a record in one aligned 8-byte load.  Without a slab filter a
// point's slot is its index; with one, the tile's survivors take consecutive slots reserved with one
// atomicAdd per tile (their order is irrelevant: the slot only ties the sort word to its gather word,
// and the per-voxel sums are integers).
// PRE: runs of consecutive same-voxel items inside a thread's eight records become ONE sort element whose gather
// word carries the run's partial sums: count (4 bits) | R, G, B sums (11 bits each) | x, y, z offset sums (off_bits + 3
// bits each).  The host launches both forms; the plan (made on the device) says which one runs.
template <int BITS, bool FILTER, bool PRE>
__global__ void __launch_bounds__(SW_KH_THREADS)
sw_keys_hist(const int16_t *__restrict__ rec, int n, int plan_slot,
             uint64_t *__restrict__ words, uint64_t *__restrict__ pay, uint32_t *__restrict__ ghist,
             uint32_t *__restrict__ slot_counter, const int32_t *__restrict__ n_dev = nullptr) {
    constexpr int BINS = 1 << BITS;
    if (n_dev) n = max(0, min(n, *n_dev));
    const SweepPlan &plan = c_sweep_plan[plan_slot];
    const SweepGeom &g = plan.g;
    const int passes = plan.passes;
    if (passes == 0 || (plan.pre != 0) != PRE) return;
    extern __shared__ uint32_t sw_hist[];    // [passes][BINS]
    __shared__ uint32_t warp_cnt[SW_KH_THREADS / 32], s_base;
    for (int k = threadIdx.x; k < passes * BINS; k += SW_KH_THREADS) sw_hist[k] = 0;
    __syncthreads();
    const bool aligned = (((uintptr_t)rec) & 15) == 0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_tiles = (n + SW_KH_TILE - 1) / SW_KH_TILE;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int first = tile * SW_KH_TILE + threadIdx.x * SW_KH_ITEMS;
        const int cnt = max(0, min(SW_KH_ITEMS, n - first));
        uint32_t w[20];
        if (cnt > 0) sw_load8(rec, n, first, cnt == SW_KH_ITEMS && aligned, w);
        uint64_t word[SW_KH_ITEMS], pw[SW_KH_ITEMS];
        uint32_t keep = 0;     // bit k: item k exists and lies inside the slab (PRE: and starts a run)
        uint32_t cont = 0;     // PRE, bit k: item k continues the run of item k - 1
        const int pw_ = g.off_bits + 3;
#pragma unroll
        for (int k = 0; k < SW_KH_ITEMS; ++k) {
            if (k < cnt) {
                const int x = sw_half(w, 5 * k), y = sw_half(w, 5 * k + 1), z = sw_half(w, 5 * k + 2);
                const uint32_t s3 = (uint32_t)sw_half(w, 5 * k + 3) & 0xFFFFu, s4 = (uint32_t)sw_half(w, 5 * k + 4) & 0xFFu;
                const uint32_t qx = sw_q(x, g), qy = sw_q(y, g), qz = sw_q(z, g);
                if (!(FILTER || PRE) || ((int)qz >= g.z_lo && (int)qz < g.z_hi)) keep |= 1u << k;
                word[k] = ((uint64_t)(qz - (uint32_t)g.z0) * g.dy + (uint64_t)(qy - (uint32_t)g.y0)) * g.dx +
                          (uint64_t)(qx - (uint32_t)g.x0);
                const uint64_t ox = (uint64_t)(uint32_t)(x + g.bias - g.leaf * (int)qx),
                               oy = (uint64_t)(uint32_t)(y + g.bias - g.leaf * (int)qy),
                               oz = (uint64_t)(uint32_t)(z + g.bias - g.leaf * (int)qz);
                if (PRE) {
                    pw[k] = 1ull | ((uint64_t)(s3 & 0xFFu) << 4) | ((uint64_t)(s3 >> 8) << 15) | ((uint64_t)s4 << 26) |
                            (ox << 37) | (oy << (37 + pw_)) | (oz << (37 + 2 * pw_));
                    // the same voxel as the item before (both inside): equal keys <=> equal voxels inside the box
                    if (k > 0 && ((keep >> (k - 1)) & 3u) == 3u && word[k] == word[k - 1]) cont |= 1u << k;
                } else {
                    pw[k] = (uint64_t)(s3 | (s4 << 16)) | (ox << 24) | (oy << (24 + g.off_bits)) | (oz << (24 + 2 * g.off_bits));
                }
            } else {
                word[k] = 0; pw[k] = 0;
            }
        }
        if (PRE) {
            // back to front: a run's sums collect in its first item (field widths leave room for eight points)
            uint64_t acc = 0;
#pragma unroll
            for (int k = SW_KH_ITEMS - 1; k >= 0; --k) {
                if ((keep >> k) & 1u) {
                    acc += pw[k];
                    if (!((cont >> k) & 1u)) { pw[k] = acc; acc = 0; }
                }
            }
            keep &= ~cont;
        }
        uint32_t slot0 = (uint32_t)first;     // slot of this thread's first surviving item
        if (FILTER || PRE) {
            const uint32_t c = __popc(keep);
            uint32_t inc = c;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            if (lane == 31) warp_cnt[warp] = inc;
            __syncthreads();
            uint32_t before = 0, total = 0;
#pragma unroll
            for (int v = 0; v < SW_KH_THREADS / 32; ++v) {
                const uint32_t t = warp_cnt[v];
                if (v < warp) before += t;
                total += t;
            }
            if (threadIdx.x == 0) s_base = total ? atomicAdd(slot_counter, total) : 0u;
            __syncthreads();
            slot0 = s_base + before + inc - c;
        }
        if (!(FILTER || PRE) && cnt == SW_KH_ITEMS) {
#pragma unroll
            for (int k = 0; k < SW_KH_ITEMS; ++k) word[k] = (word[k] << g.idx_bits) | (uint64_t)(slot0 + k);
            uint4 *o = reinterpret_cast<uint4 *>(words + first), *q = reinterpret_cast<uint4 *>(pay + first);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                o[j] = make_uint4((uint32_t)word[2 * j], (uint32_t)(word[2 * j] >> 32),
                                  (uint32_t)word[2 * j + 1], (uint32_t)(word[2 * j + 1] >> 32));
                q[j] = make_uint4((uint32_t)pw[2 * j], (uint32_t)(pw[2 * j] >> 32),
                                  (uint32_t)pw[2 * j + 1], (uint32_t)(pw[2 * j + 1] >> 32));
            }
        } else {
            uint32_t slot = slot0;
#pragma unroll
            for (int k = 0; k < SW_KH_ITEMS; ++k)
                if ((keep >> k) & 1u) {
                    word[k] = (word[k] << g.idx_bits) | (uint64_t)slot;
                    words[slot] = word[k];
                    pay[slot] = pw[k];
                    ++slot;
                }
        }
        // Neighbouring pixels mostly share their upper digits: count runs, not elements, and let the
        // lanes whose last run has the same digit add it once (a same-address shared-memory atomic
        // is serialised lane by lane).
        for (int p = 0; p < passes; ++p) {
            const int shift = g.idx_bits + p * BITS;
            uint32_t *h = sw_hist + p * BINS;
            uint32_t prev = (uint32_t)(BINS + lane), run = 0;
#pragma unroll
            for (int k = 0; k < SW_KH_ITEMS; ++k) {
                if ((keep >> k) & 1u) {
                    const uint32_t d = (uint32_t)(word[k] >> shift) & (BINS - 1);
                    if (run && d == prev) {
                        ++run;
                    } else {
                        if (run) atomicAdd(h + prev, run);
                        prev = d;
                        run = 1;
                    }
                }
            }
            // merge the lanes' last runs: consecutive lanes with the same digit add once (a masked
            // __reduce_add_sync over match_any groups compiles to a loop over the groups -- 15 trips
            // per call on this data -- so the merge is a plain inclusive scan + head flags instead)
            const uint32_t up = __shfl_up_sync(0xffffffffu, prev, 1);
            const bool head = lane == 0 || up != prev;
            uint32_t inc = run;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            const uint32_t hm = __ballot_sync(0xffffffffu, head);
            const uint32_t above = lane == 31 ? 0u : (hm & (0xFFFFFFFEu << lane));     // heads after this lane
            const int last = above ? __ffs(above) - 2 : 31;                             // last lane of my segment
            const uint32_t seg_end = __shfl_sync(0xffffffffu, inc, last);
            const uint32_t tot = seg_end - (inc - run);
            if (head && tot) atomicAdd(h + prev, tot);
        }
        if (FILTER || PRE) __syncthreads();   // warp_cnt / s_base are rewritten by the next tile
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
    static constexpr size_t SMEM = (size_t)TILE * 16 + (size_t)(WARPS + 2) * BINS * 4;     // raw + ordered tile, counters
    static_assert(DPT >= 1 && DPT * THREADS == BINS, "every thread owns BINS / THREADS digits");
};

// Decoupled look-back, two levels.  Every tile publishes its digit counts (tile words).  The last
// tile of each group of SW_GROUP has to add up the 15 tiles before it anyway: that sum plus its own
// counts is the group's total, which it publishes as a group word, followed by the group's inclusive
// prefix once it knows what lies before the group.  A tile's digits-before = the earlier tiles of
// its own group + a walk over group words, 16 tiles per word: with ~450 tiles in flight a one-level
// walk was most of the kernel.  One thread per digit, a step's loads issued together.
template <int BITS, int THREADS, int ITEMS, bool BALLOT>
__global__ void __launch_bounds__(THREADS, (BITS > 8 || THREADS > 256 ? 2 : (ITEMS <= 8 ? 4 : 3)))
sw_pass(const uint64_t *__restrict__ in, uint64_t *__restrict__ out, int plan_slot, int pass,
        const uint32_t *__restrict__ gbase, uint32_t *status, uint32_t *gstatus,
        uint32_t *ticket, uint32_t *err, int probe, int poll_ns) {
    using Cfg = SweepPassCfg<BITS, THREADS, ITEMS>;
    const SweepPlan &plan = c_sweep_plan[plan_slot];
    if (pass >= plan.passes || (int)blockIdx.x >= plan.n_tiles) return;    // grids are sized for the worst case
    const int n = plan.m, shift = plan.g.idx_bits + pass * BITS;
    constexpr int BINS = Cfg::BINS, WARPS = Cfg::WARPS, TILE = Cfg::TILE, DPT = Cfg::DPT;
    extern __shared__ __align__(16) uint8_t sw_smem[];
    uint64_t *skeys = reinterpret_cast<uint64_t *>(sw_smem);     // [TILE] the tile in digit order
    uint64_t *raw = skeys + TILE;                                 // [TILE] the tile as it lies in `in`
    uint32_t *wh = reinterpret_cast<uint32_t *>(raw + TILE);     // [WARPS][BINS] per-warp digit counters
    uint32_t *tile_excl = wh + WARPS * BINS;                      // [BINS] first tile-local slot of a digit
    uint32_t *gdelta = tile_excl + BINS;                          // [BINS] global slot = gdelta + tile-local slot
    __shared__ uint32_t s_tile, s_bulk, warp_tot[33];
    __shared__ __align__(8) uint64_t s_bar;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // The tile is one contiguous span of `in`: a single bulk copy (TMA) brings it to shared memory while
    // the counters are cleared.  (Sixteen register loads per thread were sunk by ptxas next to their
    // uses to hold the register budget -- twelve exposed DRAM round trips per tile.)
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&s_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t t = atomicAdd(ticket, 1u);   // tiles are taken in the order CTAs start running
        s_tile = t;
        const long long first = (long long)t * TILE;
        const uint32_t bytes = (uint32_t)min((long long)TILE, (long long)n - first) * 8u;
        const bool bulk = (bytes & 15u) == 0 && ((reinterpret_cast<uintptr_t>(in + first)) & 15) == 0;
        s_bulk = bulk ? 1u : 0u;
        if (bulk) {
            mbar_expect_tx(smem_u32(&s_bar), bytes);
            bulk_load(smem_u32(raw), in + first, bytes, smem_u32(&s_bar));
        }
    }
    for (int k = threadIdx.x; k < WARPS * BINS; k += THREADS) wh[k] = 0;
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t group = tile / SW_GROUP;
    const int group_first = (int)(group * SW_GROUP);
    const int tbase = (int)tile * TILE;
    const int wbase = tbase + warp * (32 * ITEMS);   // warp-blocked: warp w owns 32*ITEMS consecutive words
    const uint64_t *wraw = raw + warp * (32 * ITEMS) + lane;
    uint32_t *mywh = wh + warp * BINS;
    uint32_t local[ITEMS];
    if (s_bulk) {
        mbar_wait(smem_u32(&s_bar), 0);
    } else {      // ragged last tile
        for (int j = threadIdx.x; j < TILE; j += THREADS)
            if (tbase + j < n) raw[j] = in[tbase + j];
        __syncthreads();
    }
    // lanes holding the same digit, from one ballot per digit bit (MATCH.ANY, PCS_SW_BALLOT=0, measured
    // the same: 151 vs 145 us per pass -- the pass is latency-bound).  All masks first -- the ballots of
    // different items are independent -- then the serial walk over the warp's counters.
    uint32_t mask[ITEMS];
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const bool valid = wbase + k * 32 + lane < n;
        const uint32_t d = (uint32_t)(wraw[k * 32] >> shift) & (BINS - 1);
        if (BALLOT) {
            uint32_t m = __ballot_sync(0xffffffffu, valid);
#pragma unroll
            for (int b = 0; b < BITS; ++b) {
                const bool bit = (d >> b) & 1u;
                const uint32_t v = __ballot_sync(0xffffffffu, bit);
                m &= bit ? v : ~v;
            }
            mask[k] = m;
        } else {
            mask[k] = __match_any_sync(0xffffffffu, valid ? d : (uint32_t)(BINS + lane));
        }
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        if (probe & 4) { local[k] = 0; continue; }
        const bool valid = wbase + k * 32 + lane < n;
        const uint32_t d = (uint32_t)(wraw[k * 32] >> shift) & (BINS - 1);
        const int leader = __ffs(mask[k]) - 1;
        const uint32_t rank = __popc(mask[k] & ((1u << lane) - 1u));
        uint32_t prev = 0;
        if (valid && lane == leader) {
            prev = mywh[d];
            mywh[d] = prev + __popc(mask[k]);
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
        // let the successors see this tile's counts as early as possible
        st_volatile_u32(status + (size_t)tile * BINS + d, SW_AGG | run);
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
            const uint64_t key = wraw[k * 32];
            const uint32_t d = (uint32_t)(key >> shift) & (BINS - 1);
            skeys[tile_excl[d] + mywh[d] + local[k]] = key;
        }
    }
    // The LAST tile of a group speaks for the group: what it finds before itself inside the group plus
    // its own counts are the group's totals (published as soon as known), and once it knows what lies
    // before the group, the group's inclusive prefix.  No counters, no atomics.
    const bool speaker = (int)tile == group_first + SW_GROUP - 1;
    // digits before this tile = the group's earlier tiles + everything before the group.  One thread
    // per digit; the loads of a step are issued together (one L2 round trip per step, not per word):
    // all <= 15 tile words of the group at once, then the group words 8 at a time, nearest first,
    // until one carries a full prefix.  A row of words is one coalesced line per warp.
    const int n_before = (int)tile - group_first;
    constexpr int GB = 16;      // group words per step
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const int d = threadIdx.x * DPT + j;
        uint32_t in_group = 0, before_group = 0;
        bool need_tiles = n_before > 0 && !(probe & 1), need_groups = group > 0 && !(probe & 1);
        int gs = (int)group - 1;
        for (int spins = 0; need_tiles || need_groups;) {
            uint32_t tv[SW_GROUP - 1], gv[GB];
            if (need_tiles) {
#pragma unroll
                for (int t = 0; t < SW_GROUP - 1; ++t)
                    tv[t] = t < n_before ? ld_volatile_u32(status + (size_t)((int)tile - 1 - t) * BINS + d) : SW_AGG;
            }
            if (need_groups) {
#pragma unroll
                for (int t = 0; t < GB; ++t)
                    gv[t] = gs - t >= 0 ? ld_volatile_u32(gstatus + (size_t)(gs - t) * BINS + d) : SW_PREFIX;
            }
            bool progress = false;
            if (need_tiles) {
                uint32_t acc = 0, flags = 0xFFFFFFFFu;
#pragma unroll
                for (int t = 0; t < SW_GROUP - 1; ++t) {
                    acc += tv[t] & SW_VALUE;
                    flags &= tv[t];
                }
                if (flags & SW_AGG) {      // every tile word carries SW_AGG once published
                    in_group = acc;
                    need_tiles = false;
                    progress = true;
                    if (speaker)
                        st_volatile_u32(gstatus + (size_t)group * BINS + d,
                                        (group == 0 ? SW_PREFIX : SW_AGG) | (in_group + cnt[j]));
                }
            }
            if (need_groups) {
                uint32_t acc = 0;
                bool ready = true, found = false;
#pragma unroll
                for (int t = 0; t < GB; ++t) {
                    if (!found) {
                        const uint32_t f = gv[t] >> 30;
                        if (f == 0) ready = false;
                        acc += gv[t] & SW_VALUE;
                        if (f == 2) found = true;
                    }
                }
                if (ready) {     // else: an unpublished group in front of the first full prefix, poll again
                    before_group += acc;
                    gs -= GB;
                    if (found) need_groups = false;
                    progress = true;
                }
            }
            if (!progress) {
                if (++spins >= SW_SPIN_LIMIT) { atomicExch(err, 1u); break; }
                if (poll_ns) __nanosleep(poll_ns);      // polling in a tight loop starves the tiles being waited for
            }
        }
        gdelta[d] = gbase[d] + before_group + in_group - tile_excl[d];
        if (speaker && group > 0)
            st_volatile_u32(gstatus + (size_t)group * BINS + d, SW_PREFIX | ((before_group + in_group + cnt[j]) & SW_VALUE));
    }
    __syncthreads();
    const int nvalid = min(TILE, n - tbase);
    for (int j = threadIdx.x; j < nvalid; j += THREADS) {
        const uint64_t k = skeys[j];
        const uint32_t d = (uint32_t)(k >> shift) & (BINS - 1);
        if (!(probe & 2)) out[gdelta[d] + (uint32_t)j] = k;
        else if (k == 0x123456789abcdefull) out[j] = k;
    }
}

// ---- 4. voxel boundaries per chunk ------------------------------------------------------
// info[c] = (position + 1 of the last head in chunk c, 0 if none) << 32 | heads in chunk c
__global__ void __launch_bounds__(256)
sw_chunk_heads(const uint64_t *__restrict__ w0, const uint64_t *__restrict__ w1, int plan_slot, uint64_t *__restrict__ info) {
    const SweepPlan &plan = c_sweep_plan[plan_slot];
    const uint64_t *__restrict__ sorted = (plan.passes & 1) ? w1 : w0;     // where the last pass left the words
    const int n = plan.m, idx_bits = plan.g.idx_bits, n_chunks = plan.n_chunks;
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
sw_chunk_scan(const uint64_t *__restrict__ info, int plan_slot, uint32_t *__restrict__ chunk_off,
              uint32_t *__restrict__ ownerpos, int32_t *__restrict__ nv_out) {
    constexpr int IPT = 4;
    const int n_chunks = c_sweep_plan[plan_slot].n_chunks;
    if (n_chunks == 0) return;
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
    // exact: key * d < 2^64 (key < 2^48, d < 2^16)
    const uint64_t t = g.mdx ? __umul64hi(key, g.mdx) : key;           // key / dx
    const uint64_t u = g.mdy ? __umul64hi(t, g.mdy) : t;               // key / (dx * dy)
    const int qx = (int)(key - t * g.dx) + g.x0, qy = (int)(t - u * g.dy) + g.y0, qz = (int)u + g.z0;
    const uint32_t cnt = a[6];
    // floor(a / cnt) for six sums: one reciprocal, then a multiply, a truncation and a +-1 fix-up
    // each.  a / cnt <= max(255, leaf - 1) < 2^15, so the float product is off by far less than 1.
    const float r = __frcp_rn((float)cnt);
    uint32_t q[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        // a + cnt < 2^32 (the ABI bounds n * max(256, leaf)), so the remainder is exact modulo 2^32
        uint32_t e = (uint32_t)__float2int_rz(__fmul_rn((float)a[k], r));
        const int rem = (int)(a[k] - e * cnt);
        if (rem < 0) --e;
        else if ((uint32_t)rem >= cnt) ++e;
        q[k] = e;
    }
    int16_t *o = out + 5 * (size_t)vid;
    o[0] = (int16_t)(g.leaf * qx - g.bias + (int)q[0]);
    o[1] = (int16_t)(g.leaf * qy - g.bias + (int)q[1]);
    o[2] = (int16_t)(g.leaf * qz - g.bias + (int)q[2]);
    o[3] = (int16_t)(q[3] | (q[4] << 8));
    o[4] = (int16_t)q[5];
}

// slots[c] = 8 words: 7 sums of chunk c's last voxel when it continues into later chunks, word 7 =
// "finalise me"; slotkey[c] = that voxel's key.  All the chunk's sort words and gather words are
// requested before the first round is reduced (a round is one dependent DRAM round trip otherwise).
// PACKED (off_bits <= 5, i.e. leaf <= 32): the seven running sums of a round fit three words
// (3 x 10-bit offset sums | R 13, G 13, count 6 | B 13), so the segmented scan moves 3 registers
// per step instead of 7.
template <bool PACKED>
__global__ void __launch_bounds__(256)
sw_reduce(const uint64_t *__restrict__ w0, const uint64_t *__restrict__ w1, const uint64_t *__restrict__ pay, int plan_slot,
          const uint32_t *__restrict__ chunk_off, const uint32_t *__restrict__ ownerpos,
          uint32_t *__restrict__ slots, uint64_t *__restrict__ slotkey, int16_t *__restrict__ out) {
    const SweepPlan &plan = c_sweep_plan[plan_slot];
    const SweepGeom &g = plan.g;
    // the host launches both forms: packed sums need single points of a small leaf, merged runs need the wide scan
    if (PACKED != (plan.pre == 0 && g.off_bits <= 5)) return;
    const uint64_t *__restrict__ sorted = (plan.passes & 1) ? w1 : w0;
    const int n = plan.m, n_chunks = plan.n_chunks;
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
    bool carry_has_head = false, carry_closed = false;
    uint64_t carry_key = before;
    const uint64_t next_chunk_key = base + SW_CHUNK < n ? sorted[base + SW_CHUNK] >> g.idx_bits : 0;
#pragma unroll
    for (int r = 0; r < SW_CHUNK_ROUNDS; ++r) {
        const int e = base + r * 32 + lane;
        const bool valid = e < n;
        const uint64_t key = word[r] >> g.idx_bits;
        uint64_t prev = __shfl_up_sync(0xffffffffu, key, 1);
        if (lane == 0) prev = before;
        const bool head = valid && (e == 0 || key != prev);
        const uint32_t hm = __ballot_sync(0xffffffffu, head);
        const uint32_t le = hm & (0xffffffffu >> (31 - lane));
        const uint32_t vid = run_vid + __popc(le);
        const bool started_here = le != 0;
        const int seg_start = started_here ? 31 - __clz(le) : 0;
        uint32_t v[7] = {0, 0, 0, 0, 0, 0, 0};
        if (PACKED) {
            uint32_t A = 0, B = 0, Cc = 0;
            if (valid) {
                const uint32_t lo = (uint32_t)pw[r];
                const uint32_t offs = (uint32_t)(pw[r] >> 24);
                A = (offs & om) | (((offs >> g.off_bits) & om) << 10) | (((offs >> (2 * g.off_bits)) & om) << 20);
                B = (lo & 0xFF) | (((lo >> 8) & 0xFF) << 13) | (1u << 26);
                Cc = (lo >> 16) & 0xFF;
            }
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t a = __shfl_up_sync(0xffffffffu, A, d), b = __shfl_up_sync(0xffffffffu, B, d),
                               c = __shfl_up_sync(0xffffffffu, Cc, d);
                if (lane >= d && lane - d >= seg_start) { A += a; B += b; Cc += c; }
            }
            v[0] = A & 1023u; v[1] = (A >> 10) & 1023u; v[2] = A >> 20;
            v[3] = B & 8191u; v[4] = (B >> 13) & 8191u; v[6] = B >> 26; v[5] = Cc;
        } else {
            if (valid && plan.pre) {         // a merged run: count | R, G, B sums | offset sums
                const int pw_ = g.off_bits + 3;
                const uint32_t pm = (1u << pw_) - 1;
                const uint64_t offs = pw[r] >> 37;
                v[0] = (uint32_t)offs & pm;
                v[1] = (uint32_t)(offs >> pw_) & pm;
                v[2] = (uint32_t)(offs >> (2 * pw_)) & pm;
                v[3] = (uint32_t)(pw[r] >> 4) & 2047u; v[4] = (uint32_t)(pw[r] >> 15) & 2047u;
                v[5] = (uint32_t)(pw[r] >> 26) & 2047u; v[6] = (uint32_t)pw[r] & 15u;
            } else if (valid) {
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
        }
        // key of the voxel a lane belongs to (its own key, except on the padding lanes past n)
        const uint64_t key_at_start = __shfl_sync(0xffffffffu, key, seg_start);
        const uint64_t seg_key = started_here ? key_at_start : carry_key;
        // lane 31's voxel ends with this round when the next element (next round, or next chunk) starts
        // a new voxel or does not exist; every voxel end of the round goes through ONE emit site
        // (divergent copies of the emit were each paid in full by the whole warp)
        uint64_t nxt_key;
        bool nxt_valid;
        if (r + 1 < SW_CHUNK_ROUNDS) {
            nxt_key = __shfl_sync(0xffffffffu, word[r + 1 < SW_CHUNK_ROUNDS ? r + 1 : r] >> g.idx_bits, 0);
            nxt_valid = base + (r + 1) * 32 < n;
        } else {
            nxt_key = next_chunk_key;
            nxt_valid = base + SW_CHUNK < n;
        }
        const bool end31 = valid && (!nxt_valid || nxt_key != key);
        const bool is_end = lane < 31 ? (((hm >> (lane + 1)) & 1u) != 0) : end31;
        const bool whole = started_here || carry_has_head;     // the voxel's head is in this chunk
        if (is_end) {
            uint32_t tot[7];
#pragma unroll
            for (int k = 0; k < 7; ++k) tot[k] = started_here ? v[k] : carry[k] + v[k];
            if (whole) {
                sw_emit(out, vid, tot, seg_key, g);
            } else {
#pragma unroll
                for (int k = 0; k < 7; ++k) atomicAdd(slots + 8 * (size_t)owner + k, tot[k]);
            }
        }
        if (__shfl_sync(0xffffffffu, (int)end31, 31)) {
            // nothing stays open
#pragma unroll
            for (int k = 0; k < 7; ++k) carry[k] = 0;
            carry_has_head = false;
            carry_closed = true;
        } else if (base + r * 32 < n) {
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
            carry_closed = false;
        }
        run_vid += __popc(hm);
        before = __shfl_sync(0xffffffffu, key, 31);
    }
    // a voxel still open after the last round: it runs into the next chunk, or the data ended inside
    // a round (then it is complete)
    if (lane == 0 && !carry_closed) {
        const bool complete = base + SW_CHUNK >= n || next_chunk_key != carry_key;
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
                 const uint32_t *__restrict__ slots, const uint64_t *__restrict__ slotkey, int plan_slot,
                 int16_t *__restrict__ out) {
    const SweepPlan &plan = c_sweep_plan[plan_slot];
    const SweepGeom &g = plan.g;
    const int n_chunks = plan.n_chunks;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks || slots[8 * (size_t)c + 7] == 0) return;
    uint32_t a[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) a[k] = slots[8 * (size_t)c + k];
    sw_emit(out, chunk_off[c] + (uint32_t)info[c] - 1u, a, slotkey[c], g);
}

// the result word: a negative status beats the count (a look-back that gave up, inconsistent counts)
__global__ void sw_result(int32_t *__restrict__ nv, const uint32_t *__restrict__ err, int plan_slot) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const SweepPlan &plan = c_sweep_plan[plan_slot];
    if (plan.status != 0 || plan.m == 0) return;                 // sw_plan wrote the result already
    if (*err != 0 || *nv < 1 || *nv > plan.m) *nv = -2;
}

// ---- host side ----------------------------------------------------------------------------
inline int sweep_bits_for(long long count) {    // bits needed for values 0 .. count-1 (at least 1)
    int b = 1;
    while (b < 62 && (1ll << b) < count) ++b;
    return b;
}

inline SweepGeom sweep_base_geom(int leaf) {
    SweepGeom g{};
    g.leaf = leaf;
    g.K = (32768 + leaf - 1) / leaf;
    g.bias = leaf * g.K;
    g.magic = ((1ull << 40) / (unsigned long long)leaf) + 1ull;
    g.off_bits = sweep_bits_for(leaf);
    g.z_lo = 0;
    g.z_hi = 0x7FFFFFFF;
    return g;
}
inline int sweep_zbins(const SweepGeom &g) { return (32767 + g.bias) / g.leaf + 1; }

template <int BITS, int THREADS, int ITEMS>
inline int sweep_configure() {
    using Cfg = SweepPassCfg<BITS, THREADS, ITEMS>;
    if (cudaFuncSetAttribute(sw_pass<BITS, THREADS, ITEMS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Cfg::SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(sw_pass<BITS, THREADS, ITEMS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Cfg::SMEM) != cudaSuccess) return -2;
    return 0;
}

inline int sweep_reserve(VoxelScratch &s, size_t total) {
    if (total > s.cap) {
        cudaFree(s.buf);
        s.buf = nullptr; s.cap = 0;
        if (cudaMalloc(&s.buf, total) != cudaSuccess) { cudaGetLastError(); return -3; }
        s.cap = total;
    }
    if (!s.h_count && cudaHostAlloc(&s.h_count, 64, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return -3; }
    return 0;
}

// Cuts the qz axis into n_slabs slabs of (nearly) equal population.  kz_splits[0 .. n_slabs] are
// plane indices floor(z / leaf): slab r holds kz_splits[r] <= floor(z / leaf) < kz_splits[r + 1];
// slab_points[r] (optional) is its population.  Identical inputs give identical cuts, so ranks
// that hold the same stitched cloud agree on the plan without talking to each other.
inline int voxel_slab_plan(VoxelScratch &s, const int16_t *rec, int n, int leaf, int n_slabs, int32_t *kz_splits,
                           int32_t *slab_points, cudaStream_t cs, int sm_count) {
    SweepGeom g = sweep_base_geom(leaf);
    const int zbins = sweep_zbins(g);
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t total = al(64) + al((size_t)zbins * 4);
    int rc = sweep_reserve(s, total);
    if (rc) return rc;
    uint32_t *bounds = (uint32_t *)s.buf, *zhist = (uint32_t *)(s.buf + al(64));
    if (cudaMemsetAsync(s.buf, 0, total, cs) != cudaSuccess || cudaMemsetAsync(bounds, 0xFF, 12, cs) != cudaSuccess) return -2;
    const int in_smem = (size_t)zbins * 4 <= 40 * 1024 ? 1 : 0;
    const int tiles = (n + SW_KH_TILE - 1) / SW_KH_TILE;
    const int grid = std::max(1, std::min(tiles, std::max(1, sm_count) * 8));
    sw_bounds<true><<<grid, SW_KH_THREADS, in_smem ? (size_t)zbins * 4 : 0, cs>>>(rec, n, g, bounds, zhist, zbins, in_smem);
    std::vector<uint32_t> h(zbins);
    if (cudaMemcpyAsync(h.data(), zhist, (size_t)zbins * 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(cs) != cudaSuccess || cudaGetLastError() != cudaSuccess) return -2;
    int lo = 0, hi = zbins - 1;
    while (lo < zbins && !h[lo]) ++lo;
    while (hi >= 0 && !h[hi]) --hi;
    if (lo > hi) { lo = 0; hi = -1; }     // no points
    // slab r ends at the first plane where the running count reaches (r + 1) * n / n_slabs
    long long cum = 0;
    int q = lo;
    kz_splits[0] = lo - g.K;
    for (int r = 0; r < n_slabs; ++r) {
        const long long target = (long long)n * (r + 1) / n_slabs;
        const long long start = cum;
        while (q <= hi && (cum < target || r == n_slabs - 1)) cum += h[q++];
        kz_splits[r + 1] = q - g.K;
        if (slab_points) slab_points[r] = (int32_t)(cum - start);
    }
    return 0;
}

// Enqueues the whole merge on `cs` without any host synchronisation.  *nv_dev_out (device int32) receives the
// voxel count when the stream gets there, or -4 when the sort word does not fit (the caller picks another
// variant), or -2 (a look-back gave up / inconsistent counts).  Returns 0, -2 on a CUDA error, -3 on allocation
// failure, -4 when the limits are known to be exceeded up front.
// slab: merge only the points with kz_lo <= floor(z / leaf) < kz_hi.
template <int BITS, int THREADS, int ITEMS>
inline int voxel_merge_sweep_enqueue(VoxelScratch &s, const int16_t *rec, int n, int leaf, int16_t *out, cudaStream_t cs,
                                     int sm_count, int plan_slot, bool slab, int kz_lo, int kz_hi, int32_t **nv_dev_out,
                                     const int32_t *n_dev = nullptr) {
    using Cfg = SweepPassCfg<BITS, THREADS, ITEMS>;
    constexpr int BINS = Cfg::BINS;
    SweepGeom g = sweep_base_geom(leaf);
    const int zbins = sweep_zbins(g);
    if (24 + 3 * g.off_bits > 64 || n >= (1 << 30)) return -4;
    if (slab) {
        g.z_lo = std::max(0, std::min(zbins, kz_lo + g.K));
        g.z_hi = std::max(0, std::min(zbins, kz_hi + g.K));
    }
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // sized for the worst case (every point inside, full-range keys): the plan only exists on the device
    // [bounds][plan][nv, err][words0][words1][pay] | zeroed: [ghist][tickets][slot counter][group words][status][slots] |
    // [info][chunk_off][ownerpos][slotkey]
    const int passes_max = (3 * sweep_bits_for(zbins) + BITS - 1) / BITS;
    const int tiles_max = (n + Cfg::TILE - 1) / Cfg::TILE, chunks_max = (n + SW_CHUNK - 1) / SW_CHUNK;
    const int groups_max = (tiles_max + SW_GROUP - 1) / SW_GROUP;
    const size_t o_bounds = 0, o_plan = al(64), o_nv = o_plan + al(sizeof(SweepPlan)), o_w0 = o_nv + al(64),
                 o_w1 = o_w0 + al((size_t)n * 8), o_pay = o_w1 + al((size_t)n * 8),
                 o_zero = o_pay + al((size_t)n * 8),
                 o_ghist = o_zero, o_ticket = o_ghist + al((size_t)passes_max * BINS * 4), o_cnt = o_ticket + al(64),
                 o_gstat = o_cnt + al(64),
                 o_status = o_gstat + al((size_t)passes_max * groups_max * BINS * 4),
                 o_slots = o_status + al((size_t)passes_max * tiles_max * BINS * 4),
                 o_zero_end = o_slots + al((size_t)chunks_max * 32),
                 o_info = o_zero_end, o_off = o_info + al((size_t)chunks_max * 8), o_owner = o_off + al((size_t)chunks_max * 4),
                 o_skey = o_owner + al((size_t)chunks_max * 4), total = o_skey + al((size_t)chunks_max * 8);
    int rc = sweep_reserve(s, total);
    if (rc) return rc;
    uint32_t *bounds = (uint32_t *)(s.buf + o_bounds);
    SweepPlan *d_plan = (SweepPlan *)(s.buf + o_plan);
    int32_t *nv_dev = (int32_t *)(s.buf + o_nv);
    uint32_t *err = (uint32_t *)(s.buf + o_nv) + 1;
    *nv_dev_out = nv_dev;
    if (cudaMemsetAsync(bounds, 0, o_w0, cs) != cudaSuccess || cudaMemsetAsync(bounds, 0xFF, 12, cs) != cudaSuccess ||
        cudaMemsetAsync(s.buf + o_zero, 0, o_zero_end - o_zero, cs) != cudaSuccess)
        return -2;
    if (slab && g.z_lo >= g.z_hi) return 0;        // empty slab: *nv_dev is 0
    // ---- occupied box (and the slab's population), then the plan, on the device
    const int kh_tiles = (n + SW_KH_TILE - 1) / SW_KH_TILE;
    const int kh_grid = std::min(kh_tiles, std::max(1, sm_count) * 8);
    sw_bounds<false><<<kh_grid, SW_KH_THREADS, 0, cs>>>(rec, n, g, bounds, nullptr, 0, 0, n_dev);
    // (with a device-side count the slots are the record indices below that count: index bits from m, as for a slab)
    static const int prereduce = pipe_knob("PCS_SW_PREREDUCE", 1, 0, 1);     // 0: never merge runs before the sort
    sw_plan<<<1, 32, 0, cs>>>(bounds, g, n, (slab || n_dev) ? 1 : 0, Cfg::TILE, BITS, prereduce, d_plan, nv_dev);
    if (cudaMemcpyToSymbolAsync(c_sweep_plan, d_plan, sizeof(SweepPlan), (size_t)plan_slot * sizeof(SweepPlan),
                                cudaMemcpyDeviceToDevice, cs) != cudaSuccess)
        return -2;
    uint64_t *w0 = (uint64_t *)(s.buf + o_w0), *w1 = (uint64_t *)(s.buf + o_w1), *pay = (uint64_t *)(s.buf + o_pay);
    uint32_t *ghist = (uint32_t *)(s.buf + o_ghist), *ticket = (uint32_t *)(s.buf + o_ticket),
             *slot_counter = (uint32_t *)(s.buf + o_cnt),
             *status = (uint32_t *)(s.buf + o_status), *slots = (uint32_t *)(s.buf + o_slots),
             *gstatus = (uint32_t *)(s.buf + o_gstat),
             *chunk_off = (uint32_t *)(s.buf + o_off), *ownerpos = (uint32_t *)(s.buf + o_owner);
    uint64_t *info = (uint64_t *)(s.buf + o_info), *slotkey = (uint64_t *)(s.buf + o_skey);
    // one of the two runs (the plan decides on the device whether runs are merged)
    if (slab)
        sw_keys_hist<BITS, true, false><<<kh_grid, SW_KH_THREADS, (size_t)passes_max * BINS * 4, cs>>>(rec, n, plan_slot, w0, pay, ghist, slot_counter, n_dev);
    else
        sw_keys_hist<BITS, false, false><<<kh_grid, SW_KH_THREADS, (size_t)passes_max * BINS * 4, cs>>>(rec, n, plan_slot, w0, pay, ghist, slot_counter, n_dev);
    if (prereduce)
        sw_keys_hist<BITS, true, true><<<kh_grid, SW_KH_THREADS, (size_t)passes_max * BINS * 4, cs>>>(rec, n, plan_slot, w0, pay, ghist, slot_counter, n_dev);
    sw_hist_scan<BITS><<<passes_max, BINS, 0, cs>>>(ghist);
    static const bool ballot = pipe_knob("PCS_SW_BALLOT", 1, 0, 1) != 0;     // 0: MATCH.ANY ranking (tuning knob)
#ifdef PCS_SW_PROBES      // timing probes skip phases of sw_pass (WRONG results): only in builds made for tools/probe_vox.py
    static const int probe = pipe_knob("PCS_SW_PROBE", 0, 0, 15);
#else
    constexpr int probe = 0;
#endif
    static const int poll_ns = pipe_knob("PCS_SW_POLL_NS", 512, 0, 4096);     // pause between look-back polls (0/64/200/600/1500 ns: 1.118/1.101/1.092/1.067/1.074 ms)
    for (int p = 0; p < passes_max; ++p) {
        auto kern = ballot ? sw_pass<BITS, THREADS, ITEMS, true> : sw_pass<BITS, THREADS, ITEMS, false>;
        kern<<<tiles_max, THREADS, Cfg::SMEM, cs>>>(
            w0, w1, plan_slot, p, ghist + (size_t)p * BINS, status + (size_t)p * tiles_max * BINS,
            gstatus + (size_t)p * groups_max * BINS, ticket + p, err, probe, poll_ns);
        std::swap(w0, w1);
    }
    // the chunk kernels pick the buffer the last pass wrote (passes is only known on the device)
    uint64_t *wa = (uint64_t *)(s.buf + o_w0), *wb = (uint64_t *)(s.buf + o_w1);
    const int cblocks = (chunks_max + 7) / 8;
    sw_chunk_heads<<<cblocks, 256, 0, cs>>>(wa, wb, plan_slot, info);
    sw_chunk_scan<<<1, 1024, 0, cs>>>(info, plan_slot, chunk_off, ownerpos, nv_dev);
    if (g.off_bits <= 5)
        sw_reduce<true><<<cblocks, 256, 0, cs>>>(wa, wb, pay, plan_slot, chunk_off, ownerpos, slots, slotkey, out);
    if (g.off_bits > 5 || prereduce)
        sw_reduce<false><<<cblocks, 256, 0, cs>>>(wa, wb, pay, plan_slot, chunk_off, ownerpos, slots, slotkey, out);
    sw_finalize_open<<<(chunks_max + 255) / 256, 256, 0, cs>>>(info, chunk_off, slots, slotkey, plan_slot, out);
    sw_result<<<1, 32, 0, cs>>>(nv_dev, err, plan_slot);
    if (cudaGetLastError() != cudaSuccess) return -2;
    return 0;
}

// The synchronous form: the voxel count (>= 0), -2 on a CUDA error, -3 on allocation failure, -4 when the sort
// word does not fit (the caller picks another variant).  One host synchronisation, at the end.
template <int BITS, int THREADS, int ITEMS>
inline int voxel_merge_sweep(VoxelScratch &s, const int16_t *rec, int n, int leaf, int16_t *out, cudaStream_t cs,
                             int sm_count, int plan_slot, bool slab = false, int kz_lo = 0, int kz_hi = 0) {
    int32_t *nv_dev = nullptr;
    int rc = voxel_merge_sweep_enqueue<BITS, THREADS, ITEMS>(s, rec, n, leaf, out, cs, sm_count, plan_slot, slab, kz_lo, kz_hi, &nv_dev);
    if (rc) return rc;
    if (cudaMemcpyAsync(s.h_count, nv_dev, 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(cs) != cudaSuccess || cudaGetLastError() != cudaSuccess) return -2;
    const int nv = s.h_count[0];
    if (nv > n) return -2;
    return nv;          // a count, or the negative status the device reported
}

}  // namespace pcs
