// K3, third generation: voxel-grid merge as a two-level MSD partition + a sort-free local stage
// (oracle/SPEC.md s3 -- own integer spec; the reference only #includes pcl/filters/voxel_grid.h,
// src/pcs-multicamera-optimized.cpp:17).
//
// What the measurements on this part said (profiles/r02_sort_probe.json): an LSD one-sweep pass over
// 14.7 M 64-bit words costs ~150 us whoever writes it (CUB: 165 us), ballot / MATCH ranking runs at
// ~1 lane per clock per SM, a scattered 8-byte store stream at 1/7 of the copy rate -- but shared-memory
// atomics retire 6.6 lanes per clock per SM.  So: no stable ranking, no look-back, no sort at all.
//
//   0. sw_bounds<true>  occupied box in voxel units + points per z plane (one read of the records); the
//                       host turns the z histogram into SLABS: runs of consecutive z planes holding
//                       ~n/256 points and at most 2^24 voxel keys, each cut into <= 1024 SUB-BUCKETS of
//                       2^shift consecutive keys (shift chosen per slab for ~256 points per sub-bucket)
//   1. vm_l1            records -> 64-bit words  (key inside the slab) << PAYB | in-voxel offsets | RGB,
//                       partitioned by slab: shared-memory atomics hand out the in-tile ranks (the order
//                       inside a slab is irrelevant -- per-voxel sums are integers), one global atomicAdd per
//                       (tile, slab) reserves the run, the tile leaves through shared memory as runs
//   2. vm_hist2         points per sub-bucket (tiles never straddle slabs)     3. exclusive scan
//   4. vm_l2            the same partition step inside every slab, by sub-bucket
//   5. vm_local         one warp per sub-bucket: a BITMAP over its 2^shift keys (atomicOr), popcount prefix
//                       -> every occupied voxel's rank, in key order, with no sorting; per-voxel sums by
//                       shared-memory atomicAdd into rank-indexed accumulators; integer means -> records,
//                       written at the sub-bucket's point offset of a staging buffer.  Sub-buckets with more
//                       points than the packed accumulators can count (> ~4000, e.g. the voxel at a camera's
//                       origin that receives every depth hole) are taken by whole CTAs with wide accumulators.
//   6. scan of the voxel counts, vm_compact: staging -> output, ascending key order, no gaps
//
// Bit-exact with the CPU restatement for any input (all sums are integers).  Needs: leaf <= 32 mm
// (offsets fit 5 bits), a z-plane of the occupied box <= 2^24 voxels and <= 1024 slabs; the caller
// falls back to the one-sweep sort (pcs_voxel_sweep.cuh) otherwise.
#pragma once
#include <vector>

#include "pcs_voxel_sweep.cuh"

namespace pcs {

constexpr int VM_LB = 14;            // a sub-bucket spans at most 2^VM_LB voxel keys (2 KB bitmap)
constexpr int VM_F2 = 11;            // and a slab at most 2^VM_F2 sub-buckets
constexpr int VM_MAX_SLABS = 1024;
constexpr int VM_THREADS = 256, VM_ITEMS = 16, VM_TILE = VM_THREADS * VM_ITEMS;
constexpr int VM_SPAN = 8;           // sub-buckets per local task
constexpr int VM_LWARPS = 8;         // warps per CTA in vm_local
constexpr int VM_SLOTS = 256;        // accumulator slots per warp and round
constexpr int VM_BIG_SLOTS = 1024;   // accumulator slots per CTA and round (big sub-buckets)
constexpr int VM_BM_WORDS = (1 << VM_LB) / 32;
constexpr size_t VM_L2_SMEM = (size_t)VM_TILE * 10 + 3 * (size_t)(1 << VM_F2) * 4;

struct VmSlab {
    uint32_t base;      // first word of the slab in the partitioned arrays
    uint32_t sb0;       // index of its first sub-bucket
    uint32_t nsub;      // sub-buckets (<= 1024)
    uint16_t zfirst;    // first z plane, relative to the box (qz - z0)
    uint16_t shift;     // sub-bucket = 2^shift keys
};
struct VmTile {         // a tile of <= VM_TILE words inside one slab
    uint32_t begin;
    uint16_t slab, count;
};
struct VmTask {         // <= VM_SPAN consecutive sub-buckets of one slab
    uint32_t sb;
    uint16_t slab, nsb;
};

struct VmGeom {
    SweepGeom g;
    int payb;           // payload bits below the key: 24 + 3 * off_bits
    int n_slabs, dz;    // dz: z planes of the box
    int big;            // sub-buckets with more points go to the wide (CTA) path
    uint32_t m32;       // floor(2^32 / leaf) + 1: floor(u / leaf) = umulhi(u, m32) for u < 2^17 (0: leaf == 1)
    float rdx, rdy;     // 1 / dx, 1 / dy (vm_unkey)
};

// floor((v + bias) / leaf) with one 32-bit multiply-high.  m32 * leaf = 2^32 + e, 0 < e <= leaf, so the
// quotient is exact while u * e < 2^32; here u < 2^17 and leaf < 2^15.
__device__ __forceinline__ uint32_t vm_q(int v, const VmGeom &vg) {
    const uint32_t u = (uint32_t)(v + vg.g.bias);
    return vg.m32 ? __umulhi(u, vg.m32) : u;
}

// ---- 1. records -> words, partitioned by slab -------------------------------------------------
__global__ void __launch_bounds__(VM_THREADS)
vm_l1(const int16_t *__restrict__ rec, int n, VmGeom vg, const VmSlab *__restrict__ slabs,
      const uint16_t *__restrict__ zslab, uint32_t *__restrict__ cursor, uint64_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t vm_smem[];
    uint64_t *stage_w = reinterpret_cast<uint64_t *>(vm_smem);                   // [VM_TILE]
    uint16_t *stage_s = reinterpret_cast<uint16_t *>(stage_w + VM_TILE);          // [VM_TILE]
    uint32_t *cnt = reinterpret_cast<uint32_t *>(stage_s + VM_TILE);              // [n_slabs] (padded to 4)
    const int ns4 = (vg.n_slabs + 3) & ~3;
    uint32_t *excl = cnt + ns4, *gbase = excl + ns4;
    uint16_t *zfirst = reinterpret_cast<uint16_t *>(gbase + ns4);                 // [n_slabs]
    uint16_t *zs = zfirst + ns4;                                                  // [dz]
    __shared__ uint32_t warp_tot[33];
    const SweepGeom &g = vg.g;
    for (int k = threadIdx.x; k < vg.n_slabs; k += VM_THREADS) zfirst[k] = slabs[k].zfirst;
    for (int k = threadIdx.x; k < vg.dz; k += VM_THREADS) zs[k] = zslab[k];
    const bool aligned = (((uintptr_t)rec) & 15) == 0;
    const int n_tiles = (n + VM_TILE - 1) / VM_TILE;
    constexpr int SPT = (VM_MAX_SLABS + VM_THREADS - 1) / VM_THREADS;             // slabs per thread in the scan
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int k = threadIdx.x; k < ns4; k += VM_THREADS) cnt[k] = 0;
        __syncthreads();
        uint64_t word[VM_ITEMS];
        uint32_t sr[VM_ITEMS];      // slab | rank << 16, or 0xFFFFFFFF (no point / outside the z range)
#pragma unroll
        for (int h = 0; h < VM_ITEMS / 8; ++h) {
            const int first = tile * VM_TILE + threadIdx.x * VM_ITEMS + h * 8;
            const int c = max(0, min(8, n - first));
            uint32_t w[20];
            if (c > 0) sw_load8(rec, n, first, c == 8 && aligned, w);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int i = h * 8 + k;
                sr[i] = 0xFFFFFFFFu;
                word[i] = 0;
                if (k < c) {
                    const int x = sw_half(w, 5 * k), y = sw_half(w, 5 * k + 1), z = sw_half(w, 5 * k + 2);
                    const uint32_t s3 = (uint32_t)sw_half(w, 5 * k + 3) & 0xFFFFu, s4 = (uint32_t)sw_half(w, 5 * k + 4) & 0xFFu;
                    const uint32_t qx = vm_q(x, vg), qy = vm_q(y, vg), qz = vm_q(z, vg);
                    const uint32_t zr = qz - (uint32_t)g.z0;
                    const uint32_t s = zr < (uint32_t)vg.dz ? zs[zr] : 0xFFFFu;
                    if (s != 0xFFFFu) {
                        // a slab spans at most 2^24 keys: 32-bit arithmetic throughout
                        const uint32_t key = ((zr - zfirst[s]) * g.dy + (qy - (uint32_t)g.y0)) * g.dx + (qx - (uint32_t)g.x0);
                        const uint32_t ox = (uint32_t)(x + g.bias) - (uint32_t)g.leaf * qx,
                                       oy = (uint32_t)(y + g.bias) - (uint32_t)g.leaf * qy,
                                       oz = (uint32_t)(z + g.bias) - (uint32_t)g.leaf * qz;
                        // payload below the key: RGB (24 bits) | ox | oy | oz (off_bits each, <= 15 bits together)
                        const uint32_t offs = ox | (oy << g.off_bits) | (oz << (2 * g.off_bits));
                        word[i] = ((uint64_t)key << vg.payb) | ((uint64_t)offs << 24) | (uint64_t)(s3 | (s4 << 16));
                        sr[i] = s | (atomicAdd(cnt + s, 1u) << 16);
                    }
                }
            }
        }
        __syncthreads();
        // exclusive scan over the slabs; one global atomicAdd per (tile, non-empty slab) reserves its run
        uint32_t c[SPT], sum = 0;
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
            const int e = threadIdx.x * SPT + j;
            c[j] = e < vg.n_slabs ? cnt[e] : 0;
            sum += c[j];
        }
        uint32_t total;
        uint32_t run = block_exclusive_scan(sum, warp_tot, total);
#pragma unroll
        for (int j = 0; j < SPT; ++j) {
            const int e = threadIdx.x * SPT + j;
            if (e < vg.n_slabs) {
                excl[e] = run;
                if (c[j]) gbase[e] = slabs[e].base + atomicAdd(cursor + e, c[j]);
            }
            run += c[j];
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < VM_ITEMS; ++i) {
            if (sr[i] != 0xFFFFFFFFu) {
                const uint32_t s = sr[i] & 0xFFFFu, pos = excl[s] + (sr[i] >> 16);
                stage_w[pos] = word[i];
                stage_s[pos] = (uint16_t)s;
            }
        }
        __syncthreads();
        for (uint32_t j = threadIdx.x; j < total; j += VM_THREADS) {
            const uint32_t s = stage_s[j];
            out[gbase[s] + (j - excl[s])] = stage_w[j];
        }
        __syncthreads();
    }
}

// ---- 2. points per sub-bucket --------------------------------------------------------------------
__global__ void __launch_bounds__(VM_THREADS)
vm_hist2(const uint64_t *__restrict__ in, VmGeom vg, const VmSlab *__restrict__ slabs, const VmTile *__restrict__ tiles,
         uint32_t *__restrict__ hist2) {
    __shared__ uint32_t cnt[1 << VM_F2];
    const VmTile t = tiles[blockIdx.x];
    const VmSlab sl = slabs[t.slab];
    for (int k = threadIdx.x; k < (int)sl.nsub; k += VM_THREADS) cnt[k] = 0;
    __syncthreads();
    const int sh = vg.payb + sl.shift;
    for (int j = threadIdx.x; j < (int)t.count; j += VM_THREADS) atomicAdd(cnt + (uint32_t)(in[t.begin + j] >> sh), 1u);
    __syncthreads();
    for (int k = threadIdx.x; k < (int)sl.nsub; k += VM_THREADS) {
        const uint32_t v = cnt[k];
        if (v) atomicAdd(hist2 + sl.sb0 + k, v);
    }
}

// ---- 4. partition inside the slabs, by sub-bucket ------------------------------------------------
__global__ void __launch_bounds__(VM_THREADS)
vm_l2(const uint64_t *__restrict__ in, VmGeom vg, const VmSlab *__restrict__ slabs, const VmTile *__restrict__ tiles,
      const uint32_t *__restrict__ base2, uint32_t *__restrict__ cursor2, uint64_t *__restrict__ out) {
    extern __shared__ __align__(16) uint8_t vm_smem[];
    uint64_t *stage_w = reinterpret_cast<uint64_t *>(vm_smem);                   // [VM_TILE]
    uint16_t *stage_d = reinterpret_cast<uint16_t *>(stage_w + VM_TILE);          // [VM_TILE]
    uint32_t *cnt = reinterpret_cast<uint32_t *>(stage_d + VM_TILE), *excl = cnt + (1 << VM_F2), *gbase = excl + (1 << VM_F2);
    __shared__ uint32_t warp_tot[33];
    const VmTile t = tiles[blockIdx.x];
    const VmSlab sl = slabs[t.slab];
    for (int k = threadIdx.x; k < (int)sl.nsub; k += VM_THREADS) cnt[k] = 0;
    __syncthreads();
    const int sh = vg.payb + sl.shift;
    uint64_t word[VM_ITEMS];
    uint32_t dr[VM_ITEMS];
#pragma unroll
    for (int i = 0; i < VM_ITEMS; ++i) {
        const int j = threadIdx.x + i * VM_THREADS;
        dr[i] = 0xFFFFFFFFu;
        word[i] = 0;
        if (j < (int)t.count) {
            word[i] = in[t.begin + j];
            const uint32_t d = (uint32_t)(word[i] >> sh);
            dr[i] = d | (atomicAdd(cnt + d, 1u) << 16);
        }
    }
    __syncthreads();
    constexpr int DPT = (1 << VM_F2) / VM_THREADS;
    uint32_t c[DPT], sum = 0;
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const int e = threadIdx.x * DPT + j;
        c[j] = e < (int)sl.nsub ? cnt[e] : 0;
        sum += c[j];
    }
    uint32_t total;
    uint32_t run = block_exclusive_scan(sum, warp_tot, total);
#pragma unroll
    for (int j = 0; j < DPT; ++j) {
        const int e = threadIdx.x * DPT + j;
        if (e < (int)sl.nsub) {
            excl[e] = run;
            if (c[j]) gbase[e] = base2[sl.sb0 + e] + atomicAdd(cursor2 + sl.sb0 + e, c[j]);
        }
        run += c[j];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < VM_ITEMS; ++i) {
        if (dr[i] != 0xFFFFFFFFu) {
            const uint32_t d = dr[i] & 0xFFFFu, pos = excl[d] + (dr[i] >> 16);
            stage_w[pos] = word[i];
            stage_d[pos] = (uint16_t)d;
        }
    }
    __syncthreads();
    for (uint32_t j = threadIdx.x; j < total; j += VM_THREADS) {
        const uint32_t d = stage_d[j];
        out[gbase[d] + (j - excl[d])] = stage_w[j];
    }
}

// ---- 5. the local stage ---------------------------------------------------------------------------
// key (inside the slab, < 2^24) -> voxel coordinates.  key / dx by float reciprocal: (float)key is exact, the
// product is off by far less than 1, a +-1 fix-up on the integer remainder makes it exact.
__device__ __forceinline__ void vm_unkey(uint32_t key, uint32_t zfirst, const VmGeom &vg, int &qx, int &qy, int &qz) {
    const SweepGeom &g = vg.g;
    uint32_t t = (uint32_t)__float2int_rz(__fmul_rn((float)key, vg.rdx));
    int rx = (int)(key - t * g.dx);
    if (rx < 0) { --t; rx += (int)g.dx; } else if ((uint32_t)rx >= g.dx) { ++t; rx -= (int)g.dx; }
    uint32_t u = (uint32_t)__float2int_rz(__fmul_rn((float)t, vg.rdy));
    int ry = (int)(t - u * g.dy);
    if (ry < 0) { --u; ry += (int)g.dy; } else if ((uint32_t)ry >= g.dy) { ++u; ry -= (int)g.dy; }
    qx = rx + g.x0;
    qy = ry + g.y0;
    qz = (int)u + (int)zfirst + g.z0;
}

// One voxel record from its sums.
__device__ __forceinline__ void vm_record(uint16_t *o, uint32_t key, uint32_t zfirst, const uint32_t (&a)[6], uint32_t cnt,
                                          const VmGeom &vg) {
    const SweepGeom &g = vg.g;
    int qx, qy, qz;
    vm_unkey(key, zfirst, vg, qx, qy, qz);
    const float r = __frcp_rn((float)cnt);
    uint32_t q[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {     // floor(a / cnt): a < 2^21, a / cnt < 2^8 (colours) or < leaf
        uint32_t e = (uint32_t)__float2int_rz(__fmul_rn((float)a[k], r));
        const int rem = (int)(a[k] - e * cnt);
        if (rem < 0) --e;
        else if ((uint32_t)rem >= cnt) ++e;
        q[k] = e;
    }
    o[0] = (uint16_t)(g.leaf * qx - g.bias + (int)q[0]);
    o[1] = (uint16_t)(g.leaf * qy - g.bias + (int)q[1]);
    o[2] = (uint16_t)(g.leaf * qz - g.bias + (int)q[2]);
    o[3] = (uint16_t)(q[3] | (q[4] << 8));
    o[4] = (uint16_t)q[5];
}

// wide version: sums up to 2^28 points (colour sums as 4-bit halves), exact 64-bit division
__device__ __forceinline__ void vm_record_wide(uint16_t *o, uint32_t key, uint32_t zfirst, const uint32_t *a, const VmGeom &vg) {
    const SweepGeom &g = vg.g;
    int qx, qy, qz;
    vm_unkey(key, zfirst, vg, qx, qy, qz);
    const uint64_t cnt = a[9];
    const uint32_t mx = (uint32_t)((uint64_t)a[0] / cnt), my = (uint32_t)((uint64_t)a[1] / cnt), mz = (uint32_t)((uint64_t)a[2] / cnt);
    const uint32_t r = (uint32_t)((16ull * a[3] + a[4]) / cnt), gg = (uint32_t)((16ull * a[5] + a[6]) / cnt),
                   b = (uint32_t)((16ull * a[7] + a[8]) / cnt);
    o[0] = (uint16_t)(g.leaf * qx - g.bias + (int)mx);
    o[1] = (uint16_t)(g.leaf * qy - g.bias + (int)my);
    o[2] = (uint16_t)(g.leaf * qz - g.bias + (int)mz);
    o[3] = (uint16_t)(r | (gg << 8));
    o[4] = (uint16_t)b;
}

// Per warp: bm[VM_BM_WORDS] bitmap, pre[VM_BM_WORDS] popcount prefix (u16 would do; u32 keeps the code simple),
// acc[VM_SLOTS][5] packed sums (ox | oy << 16, oz | count << 16, R, G, B).
struct VmWarpSmem {
    uint32_t bm[VM_BM_WORDS];
    uint16_t pre[VM_BM_WORDS];
    uint32_t acc[VM_SLOTS * 5];     // a slot's record (10 B) overwrites the start of its own sums (20 B)
    uint16_t skey[VM_SLOTS];        // the key (inside the sub-bucket) of every slot of the round
};
struct VmBigSmem {
    uint32_t bm[VM_BM_WORDS];
    uint32_t pre[VM_BM_WORDS];
    uint32_t acc[VM_BIG_SLOTS * 10];
    uint32_t warp_tot[33];
};
union VmLocalSmem {
    VmWarpSmem w[VM_LWARPS];
    VmBigSmem big;
};

// blocks [0, big_blocks): big sub-buckets (whole CTA each, strided over the list); the others: warp tasks.
// stage[base2[sb] + j] receives voxel j of sub-bucket sb, vout[sb] the number of voxels.
__global__ void __launch_bounds__(VM_LWARPS * 32)
vm_local(const uint64_t *__restrict__ in, VmGeom vg, const VmSlab *__restrict__ slabs, const VmTask *__restrict__ tasks,
         int n_tasks, const uint32_t *__restrict__ base2, const uint32_t *__restrict__ biglist,
         const uint32_t *__restrict__ nbig_dev, int big_blocks, uint16_t *__restrict__ stage, uint32_t *__restrict__ vout) {
    extern __shared__ __align__(16) uint8_t vm_smem[];
    VmLocalSmem &sm = *reinterpret_cast<VmLocalSmem *>(vm_smem);
    const SweepGeom &g = vg.g;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t om = (1u << g.off_bits) - 1;
    if ((int)blockIdx.x < big_blocks) {
        // ---- wide path: every thread of the CTA works on one sub-bucket
        VmBigSmem &b = sm.big;
        const uint32_t nbig = *nbig_dev;
        for (uint32_t bi = blockIdx.x; bi < nbig; bi += big_blocks) {
            const uint32_t packed = biglist[bi];          // slab << 20 | sub-bucket inside the slab
            const VmSlab sl = slabs[packed >> 20];
            const uint32_t d2 = packed & 0xFFFFFu, sb = sl.sb0 + d2;
            const uint32_t beg = base2[sb], m = base2[sb + 1] - beg;
            const int W = max(1, (1 << sl.shift) >> 5);
            const uint32_t kmask = (1u << sl.shift) - 1;
            for (int k = threadIdx.x; k < W; k += blockDim.x) b.bm[k] = 0;
            __syncthreads();
            for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
                const uint32_t k = (uint32_t)(in[beg + i] >> vg.payb) & kmask;
                // most of a big sub-bucket's points repeat a few keys: look before touching the word again
                if (!((b.bm[k >> 5] >> (k & 31)) & 1u)) atomicOr(b.bm + (k >> 5), 1u << (k & 31));
            }
            __syncthreads();
            // popcount prefix over the bitmap words (W <= 512: two words per thread at most)
            uint32_t V;
            {
                uint32_t c0 = 0, c1 = 0;
                const int k0 = threadIdx.x * 2;
                if (k0 < W) c0 = __popc(b.bm[k0]);
                if (k0 + 1 < W) c1 = __popc(b.bm[k0 + 1]);
                const uint32_t ex = block_exclusive_scan(c0 + c1, b.warp_tot, V);
                if (k0 < W) b.pre[k0] = ex;
                if (k0 + 1 < W) b.pre[k0 + 1] = ex + c0;
            }
            __syncthreads();
            for (uint32_t lo = 0; lo < V; lo += VM_BIG_SLOTS) {
                const uint32_t cntr = min((uint32_t)VM_BIG_SLOTS, V - lo);
                for (uint32_t k = threadIdx.x; k < cntr * 10; k += blockDim.x) b.acc[k] = 0;
                __syncthreads();
                // Same-address shared-memory atomics serialise (~4 cycles each): a voxel that holds 100 000 depth
                // holes would cost milliseconds.  The lanes of a warp that hit the same voxel add up first (three
                // packed REDUX: offsets 3 x 10 bits, colour nibbles 2 x 3 x 10 bits), one lane per voxel commits.
                for (uint32_t i0 = 0; i0 < m; i0 += blockDim.x) {
                    const uint32_t i = i0 + threadIdx.x;
                    uint32_t slot = 0xFFFFFFFFu, p0 = 0, p1 = 0, p2 = 0;
                    if (i < m) {
                        const uint64_t w = in[beg + i];
                        const uint32_t k = (uint32_t)(w >> vg.payb) & kmask;
                        const uint32_t sl_ = b.pre[k >> 5] + __popc(b.bm[k >> 5] & ((1u << (k & 31)) - 1u)) - lo;
                        if (sl_ < cntr) {
                            slot = sl_;
                            const uint32_t lo32 = (uint32_t)w, offs = (uint32_t)(w >> 24);
                            p0 = (offs & om) | (((offs >> g.off_bits) & om) << 10) | (((offs >> (2 * g.off_bits)) & om) << 20);
                            p1 = ((lo32 >> 4) & 15u) | ((lo32 & 15u) << 10) | (((lo32 >> 12) & 15u) << 20);
                            p2 = ((lo32 >> 8) & 15u) | (((lo32 >> 20) & 15u) << 10) | (((lo32 >> 16) & 15u) << 20);
                        }
                    }
                    uint32_t todo = __ballot_sync(0xffffffffu, slot != 0xFFFFFFFFu);
                    while (todo) {
                        const int leader = __ffs(todo) - 1;
                        const uint32_t s_ = __shfl_sync(0xffffffffu, slot, leader);
                        const bool mine = slot == s_;
                        const uint32_t grp = __ballot_sync(0xffffffffu, mine);
                        const uint32_t r0 = __reduce_add_sync(0xffffffffu, mine ? p0 : 0u), r1 = __reduce_add_sync(0xffffffffu, mine ? p1 : 0u),
                                       r2 = __reduce_add_sync(0xffffffffu, mine ? p2 : 0u);
                        if (lane == leader) {
                            uint32_t *a = b.acc + s_ * 10;
                            atomicAdd(a + 0, r0 & 1023u); atomicAdd(a + 1, (r0 >> 10) & 1023u); atomicAdd(a + 2, r0 >> 20);
                            atomicAdd(a + 3, r1 & 1023u); atomicAdd(a + 4, (r1 >> 10) & 1023u);
                            atomicAdd(a + 5, r1 >> 20);   atomicAdd(a + 6, r2 & 1023u);
                            atomicAdd(a + 7, (r2 >> 10) & 1023u); atomicAdd(a + 8, r2 >> 20);
                            atomicAdd(a + 9, (uint32_t)__popc(grp));
                        }
                        todo &= ~grp;
                    }
                }
                __syncthreads();
                for (int wd = threadIdx.x; wd < W; wd += blockDim.x) {
                    uint32_t bits = b.bm[wd], slot = b.pre[wd] - lo;
                    while (bits) {
                        const int bit = __ffs(bits) - 1;
                        bits &= bits - 1;
                        if (slot < cntr) {
                            const uint32_t key = (d2 << sl.shift) | (uint32_t)(wd * 32 + bit);
                            uint16_t r[5];
                            vm_record_wide(r, key, sl.zfirst, b.acc + slot * 10, vg);
                            uint16_t *o = stage + 5 * ((size_t)beg + lo + slot);
#pragma unroll
                            for (int q = 0; q < 5; ++q) o[q] = r[q];
                        }
                        ++slot;
                    }
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) vout[sb] = V;
            __syncthreads();
        }
        return;
    }
    // ---- warp path
    VmWarpSmem &ws = sm.w[warp];
    const int wid = ((int)blockIdx.x - big_blocks) * VM_LWARPS + warp, nwarps = ((int)gridDim.x - big_blocks) * VM_LWARPS;
    constexpr int CH = 8;                              // words per lane and chunk: 256 points are in flight at once
    constexpr uint64_t NONE = ~0ull;                   // no word looks like this (bit 63 of a word is always 0)
    for (int ti = wid; ti < n_tasks; ti += nwarps) {
        const VmTask task = tasks[ti];
        const VmSlab sl = slabs[task.slab];
        const int W = max(1, (1 << sl.shift) >> 5);
        const uint32_t kmask = (1u << sl.shift) - 1;
        for (int q = 0; q < (int)task.nsb; ++q) {
            const uint32_t sb = task.sb + q, d2 = sb - sl.sb0;
            const uint32_t beg = base2[sb], m = base2[sb + 1] - beg;
            if (m == 0 || m > (uint32_t)vg.big) continue;         // empty: vout stays 0; big: the wide path writes it
            for (int k = lane; k < W; k += 32) ws.bm[k] = 0;
            __syncwarp();
            uint64_t w[CH];
            // pass 1: which keys occur (a sub-bucket of <= 256 points stays in registers for pass 2)
            for (uint32_t c0 = 0; c0 < m; c0 += 32 * CH) {
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    const uint32_t i = c0 + j * 32 + lane;
                    w[j] = i < m ? in[beg + i] : NONE;
                }
#pragma unroll
                for (int j = 0; j < CH; ++j) {
                    if (w[j] != NONE) {
                        const uint32_t k = (uint32_t)(w[j] >> vg.payb) & kmask;
                        if (!((ws.bm[k >> 5] >> (k & 31)) & 1u)) atomicOr(ws.bm + (k >> 5), 1u << (k & 31));
                    }
                }
            }
            __syncwarp();
            // popcount prefix: lane l owns words [l * per, (l + 1) * per)
            const int per = (W + 31) >> 5;
            uint32_t mine = 0;
            for (int k = 0; k < per; ++k) {
                const int wd = lane * per + k;
                if (wd < W) mine += __popc(ws.bm[wd]);
            }
            uint32_t inc = mine;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
                if (lane >= d) inc += o;
            }
            const uint32_t V = __shfl_sync(0xffffffffu, inc, 31);
            uint32_t run = inc - mine;
            for (int k = 0; k < per; ++k) {
                const int wd = lane * per + k;
                if (wd < W) {
                    ws.pre[wd] = (uint16_t)run;
                    run += __popc(ws.bm[wd]);
                }
            }
            __syncwarp();
            for (uint32_t lo = 0; lo < V; lo += VM_SLOTS) {
                const uint32_t cntr = min((uint32_t)VM_SLOTS, V - lo);
                for (uint32_t k = lane; k < cntr * 5; k += 32) ws.acc[k] = 0;
                __syncwarp();
                // pass 2: every point adds itself to its voxel's sums (rank = occupied keys below it)
                for (uint32_t c0 = 0; c0 < m; c0 += 32 * CH) {
                    if (m > 32 * CH) {
#pragma unroll
                        for (int j = 0; j < CH; ++j) {
                            const uint32_t i = c0 + j * 32 + lane;
                            w[j] = i < m ? in[beg + i] : NONE;
                        }
                    }
#pragma unroll
                    for (int j = 0; j < CH; ++j) {
                        if (w[j] != NONE) {
                            const uint32_t k = (uint32_t)(w[j] >> vg.payb) & kmask;
                            const uint32_t slot = (uint32_t)ws.pre[k >> 5] + __popc(ws.bm[k >> 5] & ((1u << (k & 31)) - 1u)) - lo;
                            if (slot < cntr) {
                                uint32_t *a = ws.acc + slot * 5;
                                const uint32_t lo32 = (uint32_t)w[j], offs = (uint32_t)(w[j] >> 24);
                                atomicAdd(a + 0, (offs & om) | (((offs >> g.off_bits) & om) << 16));
                                atomicAdd(a + 1, ((offs >> (2 * g.off_bits)) & om) | (1u << 16));
                                atomicAdd(a + 2, lo32 & 0xFFu);
                                atomicAdd(a + 3, (lo32 >> 8) & 0xFFu);
                                atomicAdd(a + 4, (lo32 >> 16) & 0xFFu);
                                ws.skey[slot] = (uint16_t)k;      // every point of the voxel stores the same value
                            }
                        }
                    }
                }
                __syncwarp();
                // one lane per voxel: sums -> record, written over the voxel's own sums
                for (uint32_t s_ = lane; s_ < cntr; s_ += 32) {
                    uint32_t *a = ws.acc + s_ * 5;
                    const uint32_t sums[6] = {a[0] & 0xFFFFu, a[0] >> 16, a[1] & 0xFFFFu, a[2], a[3], a[4]};
                    const uint32_t cnt = a[1] >> 16;
                    vm_record(reinterpret_cast<uint16_t *>(a), (d2 << sl.shift) | (uint32_t)ws.skey[s_], sl.zfirst, sums, cnt, vg);
                }
                __syncwarp();
                uint16_t *o = stage + 5 * ((size_t)beg + lo);
                const uint16_t *rs = reinterpret_cast<const uint16_t *>(ws.acc);
                for (uint32_t k = lane; k < cntr * 5; k += 32) o[k] = rs[(k / 5) * 10 + k % 5];
                __syncwarp();
            }
            if (lane == 0) vout[sb] = V;
        }
    }
}

// biglist: sub-buckets with more than `big` points, as slab << 20 | index inside the slab
__global__ void __launch_bounds__(256)
vm_find_big(const uint32_t *__restrict__ hist2, const VmSlab *__restrict__ slabs, int n_slabs, uint32_t big,
            uint32_t *__restrict__ biglist, uint32_t *__restrict__ nbig) {
    const int s = blockIdx.x;
    if (s >= n_slabs) return;
    const VmSlab sl = slabs[s];
    for (uint32_t d = threadIdx.x; d < sl.nsub; d += blockDim.x)
        if (hist2[sl.sb0 + d] > big) biglist[atomicAdd(nbig, 1u)] = ((uint32_t)s << 20) | d;
}

// ---- 6. staging -> output ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vm_compact(const uint16_t *__restrict__ stage, const uint32_t *__restrict__ base2, const uint32_t *__restrict__ vout,
           const uint32_t *__restrict__ vbase, int nsb, int16_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = gridDim.x * (blockDim.x >> 5);
    for (int sb = w; sb < nsb; sb += nw) {
        const uint32_t v = vout[sb];
        if (!v) continue;
        const uint16_t *src = stage + 5 * (size_t)base2[sb];
        uint16_t *dst = reinterpret_cast<uint16_t *>(out) + 5 * (size_t)vbase[sb];
        for (uint32_t k = lane; k < v * 5; k += 32) dst[k] = src[k];
    }
}

// the exclusive scan leaves the grand total behind the tile sums: copy it next to the data
__global__ void vm_store_total(const uint32_t *__restrict__ sums, int n_tiles, uint32_t *__restrict__ dst) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *dst = sums[n_tiles];
}

// ---- host side -----------------------------------------------------------------------------------------
struct VmPlan {
    std::vector<VmSlab> slabs;
    std::vector<uint16_t> zslab;
    std::vector<VmTile> tiles;
    std::vector<VmTask> tasks;
    uint32_t nsb = 0;
    long long m = 0;     // points inside the z range
};

inline int vm_avg() {       // tuning knob: average points per sub-bucket (PCS_VM_AVG)
    static const int v = pipe_knob("PCS_VM_AVG", 32, 1, 4096);
    return v;
}

// Slabs from the z histogram.  Returns false when the cloud does not fit this path's limits.
inline bool vm_make_plan(const SweepGeom &g, const uint32_t *zhist, int z0, int dz, VmPlan &p) {
    const unsigned long long plane = (unsigned long long)g.dx * g.dy;
    if (plane > (1ull << 24)) return false;                  // a slab's keys are 24 bits (float-exact, 32-bit arithmetic)
    const int pmax = (int)std::max<unsigned long long>(1, (1ull << 24) / plane);
    long long m = 0;
    for (int k = 0; k < dz; ++k)
        if (z0 + k >= g.z_lo && z0 + k < g.z_hi) m += zhist[z0 + k];
    p.m = m;
    p.zslab.assign(dz, 0xFFFF);
    p.slabs.clear(); p.tiles.clear(); p.tasks.clear();
    const long long target = std::max<long long>(m / 256, 8192);
    uint32_t base = 0, sb0 = 0;
    int k = 0;
    while (k < dz) {
        const bool inside = z0 + k >= g.z_lo && z0 + k < g.z_hi;
        if (!inside || zhist[z0 + k] == 0) { ++k; continue; }     // a slab starts at an occupied plane
        long long cnt = 0;
        int planes = 0, last_occupied = 0;
        while (k + planes < dz && planes < pmax && cnt < target && z0 + k + planes < g.z_hi) {
            const uint32_t h = zhist[z0 + k + planes];
            cnt += h;
            ++planes;
            if (h) last_occupied = planes;
        }
        planes = last_occupied;                                  // no trailing empty planes
        if ((int)p.slabs.size() >= VM_MAX_SLABS) return false;
        const unsigned long long range = (unsigned long long)planes * plane;
        // Sub-bucket width: a scan's points sit on surfaces, so most of a slab's key range is empty and the occupied
        // sub-buckets hold several times the average -- aim at VM_AVG points per sub-bucket over the whole range
        // (occupied ones then hold ~100-400, what one warp handles well); at least 128 keys, at most 2^VM_LB keys
        // and 2^VM_F2 sub-buckets
        int shift = 7;
        while (shift < VM_LB && ((range >> shift) > (1ull << VM_F2) || (double)cnt * (double)(1ull << shift) < (double)vm_avg() * (double)range)) ++shift;
        if ((range + (1ull << shift) - 1) >> shift > (1ull << VM_F2)) return false;
        VmSlab s{};
        s.base = base;
        s.sb0 = sb0;
        s.nsub = (uint32_t)((range + (1ull << shift) - 1) >> shift);
        s.zfirst = (uint16_t)k;
        s.shift = (uint16_t)shift;
        const int si = (int)p.slabs.size();
        for (int q = 0; q < planes; ++q) p.zslab[k + q] = (uint16_t)si;
        for (long long off = 0; off < cnt; off += VM_TILE)
            p.tiles.push_back(VmTile{base + (uint32_t)off, (uint16_t)si, (uint16_t)std::min<long long>(VM_TILE, cnt - off)});
        for (uint32_t d = 0; d < s.nsub; d += VM_SPAN)
            p.tasks.push_back(VmTask{sb0 + d, (uint16_t)si, (uint16_t)std::min<uint32_t>(VM_SPAN, s.nsub - d)});
        p.slabs.push_back(s);
        base += (uint32_t)cnt;
        sb0 += s.nsub;
        k += planes;
    }
    p.nsb = sb0;
    return true;
}

inline int vm_configure() {
    if (cudaFuncSetAttribute(vm_l1, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(vm_l2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)VM_L2_SMEM) != cudaSuccess ||
        cudaFuncSetAttribute(vm_local, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VmLocalSmem)) != cudaSuccess)
        return -2;
    return 0;
}

// Returns 0 and leaves the voxel count in *nv_dev (device), -2 CUDA error, -3 allocation failure, -4 when
// the cloud does not fit this path (the caller takes the one-sweep sort).  One host synchronisation (the
// box and the z histogram fix the partition plan).
inline int voxel_merge_msd(VoxelScratch &s, const int16_t *rec, int n, int leaf, int16_t *out, cudaStream_t cs,
                           int sm_count, bool slab, int kz_lo, int kz_hi, int32_t **nv_dev_out) {
    SweepGeom g = sweep_base_geom(leaf);
    const int zbins = sweep_zbins(g);
    // wide accumulators are uint32: offset sums (leaf - 1) * n and nibble sums 15 * n must fit
    if (g.off_bits > 5 || (long long)n * std::max(15, leaf - 1) > 0xFFFFFFFFll) return -4;
    if (slab) {
        g.z_lo = std::max(0, std::min(zbins, kz_lo + g.K));
        g.z_hi = std::max(0, std::min(zbins, kz_hi + g.K));
    }
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    // worst-case sizes, so that nothing is allocated after the plan is known
    const size_t nsb_max = (size_t)VM_MAX_SLABS << VM_F2;
    const size_t tiles_max = (size_t)n / VM_TILE + VM_MAX_SLABS + 1, tasks_max = nsb_max / VM_SPAN + VM_MAX_SLABS;
    const size_t o_bounds = 0, o_zhist = al(64), o_nv = o_zhist + al((size_t)zbins * 4), o_nbig = o_nv + al(64),
                 o_tab = o_nbig + al(64),
                 tab_bytes = al(sizeof(VmSlab) * VM_MAX_SLABS) + al((size_t)zbins * 2) + al(sizeof(VmTile) * tiles_max) +
                             al(sizeof(VmTask) * tasks_max),
                 o_cur1 = o_tab + tab_bytes, o_hist2 = o_cur1 + al(VM_MAX_SLABS * 4),
                 o_cur2 = o_hist2 + al((nsb_max + 1) * 4), o_vout = o_cur2 + al(nsb_max * 4),
                 o_sums = o_vout + al((nsb_max + 1) * 4), o_big = o_sums + al((nsb_max / SCAN_TILE + 2) * 4),
                 o_b = o_big + al(((size_t)n / 512 + 2) * 4), o_c = o_b + al((size_t)n * 8), o_stage = o_c + al((size_t)n * 8),
                 total = o_stage + al((size_t)n * 10);
    int rc = sweep_reserve(s, total);
    if (rc) return rc;
    if (!s.h_tab || s.h_tab_cap < tab_bytes + (size_t)zbins * 4 + 64) {
        if (s.h_tab) cudaFreeHost(s.h_tab);
        s.h_tab = nullptr;
        s.h_tab_cap = tab_bytes + (size_t)zbins * 4 + 64;
        if (cudaHostAlloc(&s.h_tab, s.h_tab_cap, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); s.h_tab = nullptr; return -3; }
    }
    uint32_t *bounds = (uint32_t *)(s.buf + o_bounds), *zhist = (uint32_t *)(s.buf + o_zhist);
    int32_t *nv_dev = (int32_t *)(s.buf + o_nv);
    *nv_dev_out = nv_dev;
    if (cudaMemsetAsync(s.buf, 0, o_tab, cs) != cudaSuccess || cudaMemsetAsync(bounds, 0xFF, 12, cs) != cudaSuccess) return -2;
    const int in_smem = (size_t)zbins * 4 <= 40 * 1024 ? 1 : 0;
    const int kh_tiles = (n + SW_KH_TILE - 1) / SW_KH_TILE;
    const int kh_grid = std::max(1, std::min(kh_tiles, std::max(1, sm_count) * 8));
    sw_bounds<true><<<kh_grid, SW_KH_THREADS, in_smem ? (size_t)zbins * 4 : 0, cs>>>(rec, n, g, bounds, zhist, zbins, in_smem);
    // bounds (28 B) and the z histogram come back in one copy: [bounds: 64 B][zhist]
    uint32_t *hb = (uint32_t *)(s.h_tab + tab_bytes);
    if (cudaMemcpyAsync(hb, bounds, 64, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return -2;
    uint32_t *hz = hb + 16;
    if (cudaMemcpyAsync(hz, zhist, (size_t)zbins * 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess) return -2;
    if (cudaStreamSynchronize(cs) != cudaSuccess || cudaGetLastError() != cudaSuccess) return -2;
    const long long m = hb[6];
    if (m == 0) return 0;          // *nv_dev is 0 (cleared above)
    if (m > n) return -2;
    g.x0 = (int)hb[0]; g.y0 = (int)hb[1]; g.z0 = (int)hb[2];
    g.dx = hb[3] - hb[0] + 1;
    g.dy = hb[4] - hb[1] + 1;
    g.mdx = g.dx > 1 ? ~0ull / g.dx + 1ull : 0ull;
    g.mdy = g.dy > 1 ? ~0ull / g.dy + 1ull : 0ull;
    const int dz = (int)(hb[5] - hb[2] + 1);
    VmPlan plan;
    if (!vm_make_plan(g, hz, g.z0, dz, plan) || plan.m != m) return -4;
    VmGeom vg{};
    vg.g = g;
    vg.payb = 24 + 3 * g.off_bits;
    vg.n_slabs = (int)plan.slabs.size();
    vg.dz = dz;
    // the packed accumulators count to 65535 and hold offset sums up to 65535; above ~1000 points one warp is
    // too slow anyway (a sub-bucket that large is mostly repeats of a few voxels: same-address atomics)
    vg.big = std::min(1024, 65535 / std::max(1, leaf - 1));
    vg.rdx = 1.0f / (float)g.dx;
    vg.rdy = 1.0f / (float)g.dy;
    vg.m32 = leaf == 1 ? 0u : (uint32_t)((1ull << 32) / (unsigned long long)leaf) + 1u;
    // tables: one upload
    uint8_t *ht = s.h_tab;
    const size_t t_slab = 0, t_z = al(sizeof(VmSlab) * VM_MAX_SLABS), t_tile = t_z + al((size_t)zbins * 2),
                 t_task = t_tile + al(sizeof(VmTile) * tiles_max);
    if (plan.tiles.size() > tiles_max || plan.tasks.size() > tasks_max) return -4;
    memcpy(ht + t_slab, plan.slabs.data(), sizeof(VmSlab) * plan.slabs.size());
    memcpy(ht + t_z, plan.zslab.data(), 2 * plan.zslab.size());
    memcpy(ht + t_tile, plan.tiles.data(), sizeof(VmTile) * plan.tiles.size());
    memcpy(ht + t_task, plan.tasks.data(), sizeof(VmTask) * plan.tasks.size());
    uint8_t *dt = s.buf + o_tab;
    // (four copies of exactly what is used; the regions are far apart in the worst-case layout)
    if (cudaMemcpyAsync(dt + t_slab, ht + t_slab, sizeof(VmSlab) * plan.slabs.size(), cudaMemcpyHostToDevice, cs) != cudaSuccess ||
        cudaMemcpyAsync(dt + t_z, ht + t_z, 2 * plan.zslab.size(), cudaMemcpyHostToDevice, cs) != cudaSuccess ||
        cudaMemcpyAsync(dt + t_tile, ht + t_tile, sizeof(VmTile) * plan.tiles.size(), cudaMemcpyHostToDevice, cs) != cudaSuccess ||
        cudaMemcpyAsync(dt + t_task, ht + t_task, sizeof(VmTask) * plan.tasks.size(), cudaMemcpyHostToDevice, cs) != cudaSuccess)
        return -2;
    const VmSlab *d_slabs = (const VmSlab *)(dt + t_slab);
    const uint16_t *d_zslab = (const uint16_t *)(dt + t_z);
    const VmTile *d_tiles = (const VmTile *)(dt + t_tile);
    const VmTask *d_tasks = (const VmTask *)(dt + t_task);
    uint32_t *cur1 = (uint32_t *)(s.buf + o_cur1), *hist2 = (uint32_t *)(s.buf + o_hist2), *cur2 = (uint32_t *)(s.buf + o_cur2),
             *vout = (uint32_t *)(s.buf + o_vout), *sums = (uint32_t *)(s.buf + o_sums), *biglist = (uint32_t *)(s.buf + o_big),
             *nbig = (uint32_t *)(s.buf + o_nbig);
    uint64_t *B = (uint64_t *)(s.buf + o_b), *Cw = (uint64_t *)(s.buf + o_c);
    uint16_t *stage = (uint16_t *)(s.buf + o_stage);
    const int nsb = (int)plan.nsb;
    if (cudaMemsetAsync(cur1, 0, VM_MAX_SLABS * 4, cs) != cudaSuccess ||
        cudaMemsetAsync(hist2, 0, ((size_t)nsb + 1) * 4, cs) != cudaSuccess ||
        cudaMemsetAsync(cur2, 0, (size_t)nsb * 4, cs) != cudaSuccess ||
        cudaMemsetAsync(vout, 0, ((size_t)nsb + 1) * 4, cs) != cudaSuccess)
        return -2;
    const int ns4 = (vg.n_slabs + 3) & ~3;
    const size_t smem1 = (size_t)VM_TILE * 10 + (size_t)ns4 * 12 + (size_t)ns4 * 2 + (size_t)((dz + 7) & ~7) * 2 + 16;
    const int n_tiles1 = (n + VM_TILE - 1) / VM_TILE;
    const int grid1 = std::max(1, std::min(n_tiles1, std::max(1, sm_count) * 3));
    vm_l1<<<grid1, VM_THREADS, smem1, cs>>>(rec, n, vg, d_slabs, d_zslab, cur1, B);
    const int n_tiles2 = (int)plan.tiles.size();
    vm_hist2<<<n_tiles2, VM_THREADS, 0, cs>>>(B, vg, d_slabs, d_tiles, hist2);
    vm_find_big<<<vg.n_slabs, 256, 0, cs>>>(hist2, d_slabs, vg.n_slabs, (uint32_t)vg.big, biglist, nbig);
    // hist2[0 .. nsb] -> exclusive prefix (entry nsb = m)
    exclusive_scan_u32(hist2, nsb + 1, sums, cs);
    vm_l2<<<n_tiles2, VM_THREADS, VM_L2_SMEM, cs>>>(B, vg, d_slabs, d_tiles, hist2, cur2, Cw);
    const int big_blocks = (int)std::min<long long>(m / std::max(1, vg.big) + 1, 2 * std::max(1, sm_count));
    const int n_tasks = (int)plan.tasks.size();
    const int warp_blocks = std::max(1, std::min((n_tasks + VM_LWARPS - 1) / VM_LWARPS, std::max(1, sm_count) * 8));
    vm_local<<<big_blocks + warp_blocks, VM_LWARPS * 32, sizeof(VmLocalSmem), cs>>>(
        Cw, vg, d_slabs, d_tasks, n_tasks, hist2, biglist, nbig, big_blocks, stage, vout);
    // voxel counts -> output offsets; the total is the result
    uint32_t *vbase = cur2;      // cursor2 is free again
    if (cudaMemcpyAsync(vbase, vout, ((size_t)nsb + 1) * 4, cudaMemcpyDeviceToDevice, cs) != cudaSuccess) return -2;
    exclusive_scan_u32(vbase, nsb + 1, sums, cs);
    vm_store_total<<<1, 32, 0, cs>>>(sums, (nsb + 1 + SCAN_TILE - 1) / SCAN_TILE, (uint32_t *)nv_dev);
    const int cgrid = std::max(1, std::min((nsb + 7) / 8, std::max(1, sm_count) * 16));
    vm_compact<<<cgrid, 256, 0, cs>>>(stage, hist2, vout, vbase, nsb, out);
    if (cudaGetLastError() != cudaSuccess) return -2;
    return 0;
}

}  // namespace pcs
