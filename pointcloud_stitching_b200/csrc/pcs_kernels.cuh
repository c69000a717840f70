// Kernels of the hot path, "direct" variants (register loads, shared-memory
// transposed stores).  The bulk-async pipelined K1 lives in pcs_k1_pipe.cuh.
//
//   k1_direct        z16 + RGB8 -> records      (rs2::pointcloud::calculate +
//                                                copyPointCloudXYZRGBToBufferSIMD,
//                                                src/pcs-camera-optimized.cpp:288,363-616)
//   k1a_vertices     vertex/texcoord -> records (copyPointCloudXYZRGBToBufferSIMD itself)
//   compact_*        order-preserving -c compaction (src/pcs-camera-optimized.cpp:499-577)
//   stitch_raw       concat + stride decimation (src/pcs-multicamera-client.cpp:385-395)
//   stitch_pcl       unpack + transform + append + repack
//                                               (src/pcs-multicamera-optimized.cpp:226-265,289,366)
#pragma once
#include "pcs_device.cuh"

namespace pcs {

constexpr int K1_THREADS = 256;          // one octet (8 pixels) per thread
constexpr int K1_TILE_PTS = K1_THREADS * 8;
constexpr int MAX_CAMS = 32;

struct DevJob {
    const uint16_t *z16;
    const uint8_t *color;
    int16_t *payload;
    float *xyzrgb;     // optional 16 B/pt output
    int32_t *count;    // optional record count
    uint8_t *keep;     // -c: the frame's look-back words (uint32: [0] ticket, [1 + t] tile t), zeroed before the launch
    int16_t *dense;    // (unused: the one-pass -c writes compacted records straight into payload)
    int32_t stream;
    int32_t pad;
};

// ---------------------------------------------------------------------------
// Per-warp staging: 32 lanes x 80 B of records are written to shared memory at
// lane*80 (16-byte STS, conflict-free: the eight lanes of a quarter-warp land on
// 8 distinct 4-bank groups) and streamed out as 160 coalesced 16-byte stores.
__device__ __forceinline__ void warp_store_records(uint4 *slab, const uint32_t (&w)[20], int lane,
                                                   uint8_t *dst, int valid_octets, bool aligned) {
    uint4 *mine = slab + lane * 5;
#pragma unroll
    for (int k = 0; k < 5; ++k) mine[k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    __syncwarp();
    const int n16 = valid_octets * 5;  // 16-byte chunks owned by this warp
    if (aligned) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int c = k * 32 + lane;
            if (c < n16) st_global_v4(dst + (size_t)c * 16, slab[c]);
        }
    } else {  // payload not 16-byte aligned: 2-byte stores (always legal for a short*)
        const uint16_t *s16 = reinterpret_cast<const uint16_t *>(slab);
        uint16_t *d16 = reinterpret_cast<uint16_t *>(dst);
        for (int c = lane; c < n16 * 8; c += 32) d16[c] = s16[c];
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------
// K1 direct: grid = (tiles per frame, jobs).
// CUTOFF (the reference's -c, src/pcs-camera-optimized.cpp:499-577): ONE pass.  A block counts the records its 2048
// points keep (point i gated by the test of point i ^ 3 with the reference's reversed lanes), learns how many
// the earlier tiles of its frame keep through a decoupled look-back over one status word per tile (tiles are taken by
// ticket, so every predecessor is running or done), compacts its records in shared memory at the byte alignment of
// their destination and writes them out with 16-byte stores; the last tile writes the frame's count.  No dense
// records, no keep flags, no extra launches (round 1: dense + flags, then count / scan / scatter per frame -- 193
// launches and 38 B of traffic per point for a batch of 64 frames).
// job.keep points at the frame's look-back words, zeroed by the host before the launch: [0] ticket, [1 + t] tile t.
constexpr uint32_t K1C_AGG = 1u << 30, K1C_PREFIX = 2u << 30, K1C_VALUE = (1u << 30) - 1;

// The -c form of a k1_direct block (see the comment above k1_direct): persistent, software-pipelined over the tiles it
// claims by ticket.  A block holds two tiles: A -- counted, its status word published -- and B -- claimed, its depth on the
// way.  One trip of the loop: claim C and request its depth; PHASE 1 of B: depth -> (x, z) of the eight points per
// thread, the cutoff tests, the tile's count, its status word (everything another tile waits for, and nothing of it waits
// for anybody); then A is finished: look-back over the status words (its predecessors published theirs at least one trip
// ago: no waiting in the steady state), PHASE 2: taps, transform and packing ONLY for the points A keeps (with the bounds of
// src/pcs-camera-optimized.cpp:398-401 most warps of a frame keep nothing and skip it), each record written straight to
// its compacted place in shared memory, PHASE 3: the run leaves with 16-byte stores.  Waiting only ever points from a tile
// to tiles claimed before it, and every claimed tile publishes its count without waiting: no deadlock, whatever else runs
// on the device and however few blocks are resident.
template <int MODE, bool FLOATOUT>
__device__ __forceinline__ void k1_direct_cutoff(const DevJob &job, const StreamParams &sp, uint32_t *lb, int octets, int lane,
                                                 int warp, uint4 *slab0, uint32_t &s_base, uint32_t &s_claim, uint32_t *s_warp) {
    const int n_tiles = (octets + K1_THREADS - 1) / K1_THREADS;
    if (threadIdx.x == 0) s_claim = atomicAdd(lb, 1u);
    __syncthreads();
    int tileB = (int)s_claim;
    __syncthreads();
    if (tileB >= n_tiles) return;
    uint4 dB = make_uint4(0, 0, 0, 0);
    if (tileB * K1_THREADS + (int)threadIdx.x < octets)
        dB = ld_global_nc_v4(job.z16 + ((size_t)tileB * K1_THREADS + threadIdx.x) * 8);
    int tileA = -1;
    uint32_t dzA[4] = {0, 0, 0, 0}, keptA = 0, offA = 0, totalA = 0;
    for (;;) {
        const bool haveB = tileB < n_tiles;
        if (threadIdx.x == 0) s_claim = haveB ? atomicAdd(lb, 1u) : (uint32_t)n_tiles;
        // ---- phase 1 of B
        const uint32_t dzB[4] = {dB.x, dB.y, dB.z, dB.w};
        uint32_t keptB = 0;
        {
            const int o = tileB * K1_THREADS + (int)threadIdx.x;
            if (haveB && o < octets) {
                const int p0i = o * 8, y = p0i / sp.W, x0 = p0i - y * sp.W;
                const float ny = __fdiv_rn(__fsub_rn((float)y, sp.ppy), sp.fy);
                uint32_t tests = 0;      // bit k: point k of the octet passes the cutoff test
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t z16 = (k & 1) ? (dzB[k >> 1] >> 16) : (dzB[k >> 1] & 0xFFFFu);
                    const float nx = __fdiv_rn(__fsub_rn((float)(x0 + k), sp.ppx), sp.fx);
                    float p0, p2;
                    deproject_xz<MODE>(sp, z16, nx, ny, p0, p2);
                    if (cutoff_keep(sp, p0, p2)) tests |= 1u << k;
                }
                // bit k of kept: record k of the octet survives (gated by the test of point k ^ 3 with reversed lanes)
                keptB = tests;
                if (sp.lane_rev) {
                    keptB = 0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) keptB |= ((tests >> (k ^ 3)) & 1u) << k;
                }
            }
        }
        const uint32_t cB = __popc(keptB);
        uint32_t inc = cB;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        const int tileC = (int)s_claim;
        uint4 dC = make_uint4(0, 0, 0, 0);
        if (tileC < n_tiles && tileC * K1_THREADS + (int)threadIdx.x < octets)
            dC = ld_global_nc_v4(job.z16 + ((size_t)tileC * K1_THREADS + threadIdx.x) * 8);
        uint32_t before = 0, totalB = 0;
#pragma unroll
        for (int v = 0; v < K1_THREADS / 32; ++v) {
            const uint32_t t = s_warp[v];
            if (v < warp) before += t;
            totalB += t;
        }
        const uint32_t offB = before + inc - cB;          // records kept before this thread's, inside the tile
        if (haveB && tileB > 0 && threadIdx.x == 0) *(volatile uint32_t *)(lb + 1 + tileB) = K1C_AGG | totalB;
        // ---- A: look-back, records, out
        if (tileA >= 0) {
            if (warp == 0) {
                uint32_t psum = 0;
                if (tileA > 0) {
                    int at = tileA - 1;
                    for (;;) {
                        const int t = at - lane;
                        const uint32_t v = t >= 0 ? *(volatile const uint32_t *)(lb + 1 + t) : K1C_PREFIX;
                        const uint32_t f = v >> 30;
                        const uint32_t unready = __ballot_sync(0xffffffffu, f == 0), full = __ballot_sync(0xffffffffu, f == 2);
                        const int stop = full ? __ffs(full) - 1 : 31;
                        if (unready & (0xffffffffu >> (31 - stop))) { __nanosleep(64); continue; }
                        uint32_t a = lane <= stop ? (v & K1C_VALUE) : 0u;
#pragma unroll
                        for (int d = 16; d; d >>= 1) a += __shfl_xor_sync(0xffffffffu, a, d);
                        psum += a;
                        if (full) break;
                        at -= 32;
                    }
                }
                if (lane == 0) {
                    *(volatile uint32_t *)(lb + 1 + tileA) = K1C_PREFIX | (psum + totalA);
                    s_base = psum;
                    if (tileA == n_tiles - 1 && job.count) *job.count = (int32_t)(psum + totalA);
                }
            }
            __syncthreads();
            // the kept records, in shared memory shifted by the destination's offset inside its 16 bytes
            uint8_t *dst0 = reinterpret_cast<uint8_t *>(job.payload) + (size_t)s_base * 10;
            const uint32_t sh = (uint32_t)(reinterpret_cast<uintptr_t>(dst0) & 15);
            uint8_t *stage = reinterpret_cast<uint8_t *>(slab0);
            const int o = tileA * K1_THREADS + (int)threadIdx.x;
            const bool active = o < octets;
            const uint32_t need = (FLOATOUT && job.xyzrgb) ? (active ? 0xFFu : 0u) : keptA;     // float output stays dense
            if (need) {
                const int p0i = o * 8, y = p0i / sp.W, x0 = p0i - y * sp.W;
                const float ny = __fdiv_rn(__fsub_rn((float)y, sp.ppy), sp.fy);
                uint16_t *out = reinterpret_cast<uint16_t *>(stage + sh + (size_t)offA * 10);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if ((need >> k) & 1u) {
                        const uint32_t z16 = (k & 1) ? (dzA[k >> 1] >> 16) : (dzA[k >> 1] & 0xFFFFu);
                        const int x = x0 + k;
                        const float nx = __fdiv_rn(__fsub_rn((float)x, sp.ppx), sp.fx);
                        float p0, p1, p2;
                        int xi, yi;
                        deproject_tap<MODE>(sp, z16, x, y, nx, ny, p0, p1, p2, xi, yi);
                        const uint32_t rgb = load_rgb(job.color, xi * sp.bpp + yi * sp.stride);
                        if (FLOATOUT && job.xyzrgb) {
                            float4 f;
                            f.x = affine_row(sp.tf, 0, p0, p1, p2);
                            f.y = affine_row(sp.tf, 1, p0, p1, p2);
                            f.z = affine_row(sp.tf, 2, p0, p1, p2);
                            f.w = __uint_as_float(0xFF000000u | ((rgb & 0xFF) << 16) | (rgb & 0xFF00) | ((rgb >> 16) & 0xFF));
                            reinterpret_cast<float4 *>(job.xyzrgb)[p0i + k] = f;
                        }
                        if ((keptA >> k) & 1u) {
                            const Rec r = make_record(sp.tf, p0, p1, p2, rgb);
                            out[0] = (uint16_t)r.a; out[1] = (uint16_t)(r.a >> 16);
                            out[2] = (uint16_t)r.b; out[3] = (uint16_t)(r.b >> 16);
                            out[4] = (uint16_t)r.c;
                            out += 5;
                        }
                    }
                }
            }
            __syncthreads();
            // [dst0, dst0 + bytes): 2-byte stores up to the first 16-byte boundary, 16-byte stores, 2-byte stores for the rest
            const uint32_t bytes = totalA * 10;
            const uint32_t head = min(bytes, (16u - sh) & 15u), body = (bytes - head) & ~15u, tail = bytes - head - body;
            for (uint32_t i = threadIdx.x * 2; i < head; i += K1_THREADS * 2)
                *reinterpret_cast<uint16_t *>(dst0 + i) = *reinterpret_cast<const uint16_t *>(stage + sh + i);
            for (uint32_t i = threadIdx.x * 16; i < body; i += K1_THREADS * 16)
                st_global_v4(dst0 + head + i, *reinterpret_cast<const uint4 *>(stage + sh + head + i));
            for (uint32_t i = threadIdx.x * 2; i < tail; i += K1_THREADS * 2)
                *reinterpret_cast<uint16_t *>(dst0 + head + body + i) = *reinterpret_cast<const uint16_t *>(stage + sh + head + body + i);
        }
        __syncthreads();      // s_warp, s_claim and the stage are written again by the next trip
        if (!haveB) break;    // (A was the last tile this block held)
        tileA = tileB; keptA = keptB; offA = offB; totalA = totalB;
#pragma unroll
        for (int q = 0; q < 4; ++q) dzA[q] = dzB[q];
        tileB = tileC; dB = dC;
    }
}

template <int MODE, bool CUTOFF, bool FLOATOUT>
__global__ void __launch_bounds__(K1_THREADS)
k1_direct(const DevJob *__restrict__ jobs, const StreamParams *__restrict__ streams) {
    __shared__ __align__(16) uint4 slabs[K1_THREADS / 32][32 * 5 + (CUTOFF ? 1 : 0)];     // (+16 B per warp: destination alignment)
    __shared__ StreamParams sp;
    // (s_base in a 16-byte slot of its own: the compiler fetches s_claim and s_warp[] with one wide load, which must not
    // touch a word another warp writes between the same two barriers)
    __shared__ __align__(16) uint32_t s_base_slot[4];
    __shared__ __align__(16) uint32_t s_cw[4 + K1_THREADS / 32];
    uint32_t &s_base = s_base_slot[0], &s_claim = s_cw[0];
    uint32_t *s_warp = s_cw + 4;
    const DevJob job = jobs[blockIdx.y];
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(streams + job.stream);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sp);
        for (int i = threadIdx.x; i < (int)(sizeof(StreamParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int octets = sp.N >> 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (CUTOFF) {
        k1_direct_cutoff<MODE, FLOATOUT>(job, sp, reinterpret_cast<uint32_t *>(job.keep), octets, lane, warp, &slabs[0][0], s_base, s_claim, s_warp);
        return;
    }
    const int tile0 = blockIdx.x * K1_THREADS;  // first octet of this block
    if (tile0 >= octets) return;
    const int o = tile0 + threadIdx.x;
    const bool active = o < octets;
    uint32_t w[20];
    if (active) {
        const int p0i = o * 8;
        const int y = p0i / sp.W, x0 = p0i - y * sp.W;
        const uint4 d = ld_global_nc_v4(job.z16 + p0i);
        const uint32_t dz[4] = {d.x, d.y, d.z, d.w};
        const float ny = __fdiv_rn(__fsub_rn((float)y, sp.ppy), sp.fy);
        Rec rec[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t z16 = (k & 1) ? (dz[k >> 1] >> 16) : (dz[k >> 1] & 0xFFFFu);
            const int x = x0 + k;
            const float nx = __fdiv_rn(__fsub_rn((float)x, sp.ppx), sp.fx);
            float p0, p1, p2;
            int xi, yi;
            deproject_tap<MODE>(sp, z16, x, y, nx, ny, p0, p1, p2, xi, yi);
            const uint32_t rgb = load_rgb(job.color, xi * sp.bpp + yi * sp.stride);
            rec[k] = make_record(sp.tf, p0, p1, p2, rgb);
            if (FLOATOUT && job.xyzrgb) {
                float4 f;
                f.x = affine_row(sp.tf, 0, p0, p1, p2);
                f.y = affine_row(sp.tf, 1, p0, p1, p2);
                f.z = affine_row(sp.tf, 2, p0, p1, p2);
                // pcl::PointXYZRGB colour word: b, g, r, a = 255
                f.w = __uint_as_float(0xFF000000u | ((rgb & 0xFF) << 16) | (rgb & 0xFF00) | ((rgb >> 16) & 0xFF));
                reinterpret_cast<float4 *>(job.xyzrgb)[p0i + k] = f;
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) pack_pair(rec[2 * k], rec[2 * k + 1], w + 5 * k);
    }
    const int warp_oct0 = tile0 + warp * 32;
    const int valid = min(32, octets - warp_oct0);
    if (valid <= 0) return;
    uint8_t *dst = reinterpret_cast<uint8_t *>(job.payload) + (size_t)warp_oct0 * 80;
    const bool aligned = (reinterpret_cast<uintptr_t>(dst) & 15) == 0;
    warp_store_records(slabs[warp], w, lane, dst, valid, aligned);
    if (job.count && blockIdx.x == 0 && threadIdx.x == 0) *job.count = sp.N;
}

// ---------------------------------------------------------------------------
// K1a: same inputs as the reference function.  One group of FOUR points per thread (the
// reference's own SIMD group; n % 4 == 0 is its precondition too): the group's vertices are
// 48 contiguous bytes and its tex coords 32, i.e. five 16-byte loads; the four records (40 B)
// go through a per-warp shared-memory slab and leave as coalesced 16-byte stores.
constexpr int K1A_THREADS = 256;

template <bool CUTOFF>
__global__ void __launch_bounds__(K1A_THREADS)
k1a_vertices(const float *__restrict__ xyz, const float *__restrict__ uv, int n,
             const uint8_t *__restrict__ color, const StreamParams *__restrict__ streams, int stream,
             int16_t *__restrict__ out, uint8_t *__restrict__ keep) {
    __shared__ __align__(16) uint2 slabs[K1A_THREADS / 32][32 * 5];   // 32 groups x 40 B per warp
    __shared__ StreamParams sp;
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(streams + stream);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&sp);
        for (int i = threadIdx.x; i < (int)(sizeof(StreamParams) / 4); i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int groups = n >> 2;
    const int gidx = blockIdx.x * blockDim.x + threadIdx.x;     // group index
    const int warp_g0 = gidx - lane;
    if (warp_g0 >= groups) return;
    const bool vec_in = ((reinterpret_cast<uintptr_t>(xyz) | reinterpret_cast<uintptr_t>(uv)) & 15) == 0;
    if (gidx < groups) {
        float p[12], t[8];
        if (vec_in) {
            const uint4 *x4 = reinterpret_cast<const uint4 *>(xyz) + 3 * (size_t)gidx;
            const uint4 *u4 = reinterpret_cast<const uint4 *>(uv) + 2 * (size_t)gidx;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint4 v = ld_global_nc_v4(x4 + k);
                p[4 * k] = __uint_as_float(v.x); p[4 * k + 1] = __uint_as_float(v.y);
                p[4 * k + 2] = __uint_as_float(v.z); p[4 * k + 3] = __uint_as_float(v.w);
            }
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const uint4 v = ld_global_nc_v4(u4 + k);
                t[4 * k] = __uint_as_float(v.x); t[4 * k + 1] = __uint_as_float(v.y);
                t[4 * k + 2] = __uint_as_float(v.z); t[4 * k + 3] = __uint_as_float(v.w);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 12; ++k) p[k] = __ldg(xyz + 12 * (size_t)gidx + k);
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = __ldg(uv + 8 * (size_t)gidx + k);
        }
        Rec r[4];
        uint32_t kp = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int xi = tex_to_pixel(t[2 * k], sp.cwf, sp.CW - 1);
            const int yi = tex_to_pixel(t[2 * k + 1], sp.chf, sp.CH - 1);
            const uint32_t rgb = load_rgb(color, xi * sp.bpp + yi * sp.stride);
            r[k] = make_record(sp.tf, p[3 * k], p[3 * k + 1], p[3 * k + 2], rgb);
            if (CUTOFF) kp |= (cutoff_keep(sp, p[3 * k], p[3 * k + 2]) ? 1u : 0u) << (8 * k);
        }
        if (CUTOFF) reinterpret_cast<uint32_t *>(keep)[gidx] = kp;
        uint32_t w[10];
        pack_pair(r[0], r[1], w);
        pack_pair(r[2], r[3], w + 5);
        uint2 *s2 = slabs[warp] + lane * 5;
#pragma unroll
        for (int k = 0; k < 5; ++k) s2[k] = make_uint2(w[2 * k], w[2 * k + 1]);
    }
    __syncwarp();
    const int valid = min(32, groups - warp_g0);                 // groups in this warp
    uint8_t *dst = reinterpret_cast<uint8_t *>(out) + (size_t)warp_g0 * 40;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (valid * 40) % 16 == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(slabs[warp]);
        const int n16 = valid * 40 / 16;
#pragma unroll
        for (int k = 0; k < 3; ++k)
            if (k * 32 + lane < n16) st_global_v4(dst + (size_t)(k * 32 + lane) * 16, s4[k * 32 + lane]);
    } else {
        const uint16_t *s16 = reinterpret_cast<const uint16_t *>(slabs[warp]);
        uint16_t *d16 = reinterpret_cast<uint16_t *>(dst);
        for (int c = lane; c < valid * 20; c += 32) d16[c] = s16[c];
    }
}

// ---------------------------------------------------------------------------
// -c compaction, order preserving (the reference at -t 1 emits kept points in
// raster order).  Point i is gated by the test of point (i ^ 3) when
// lane_rev != 0 -- the reference's mask is reversed within each group of four
//  -- else by its own test.
constexpr int CMP_THREADS = 256, CMP_PER_THREAD = 8, CMP_TILE = CMP_THREADS * CMP_PER_THREAD;

__device__ __forceinline__ int gate(const uint8_t *keep, int i, int lane_rev) {
    return keep[lane_rev ? (i ^ 3) : i];
}

__global__ void __launch_bounds__(CMP_THREADS)
compact_count(const uint8_t *__restrict__ keep, int n, int lane_rev, int32_t *__restrict__ tile_counts) {
    __shared__ int warp_sums[CMP_THREADS / 32];
    const int base = blockIdx.x * CMP_TILE + threadIdx.x * CMP_PER_THREAD;
    int c = 0;
#pragma unroll
    for (int k = 0; k < CMP_PER_THREAD; ++k)
        if (base + k < n) c += gate(keep, base + k, lane_rev);
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < CMP_THREADS / 32; ++k) t += warp_sums[k];
        tile_counts[blockIdx.x] = t;
    }
}

// single block: exclusive scan of the tile counts, total to *count_out (and the
// [int32] payload-size header when header != nullptr)
__global__ void __launch_bounds__(1024)
compact_scan(int32_t *__restrict__ tile_counts, int n_tiles, int32_t *__restrict__ count_out,
             int32_t *__restrict__ total_slot) {
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_tiles; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < n_tiles ? tile_counts[i] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, inc, d);
            if ((threadIdx.x & 31) >= d) inc += t;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            int t = warp_tot[threadIdx.x], inc2 = t;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int u = __shfl_up_sync(0xffffffffu, inc2, d);
                if (threadIdx.x >= d) inc2 += u;
            }
            warp_tot[threadIdx.x] = inc2 - t;  // exclusive
        }
        __syncthreads();
        const int carry = carry_s;
        if (i < n_tiles) tile_counts[i] = carry + warp_tot[threadIdx.x >> 5] + inc - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_tot[31] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (count_out) *count_out = carry_s;
        if (total_slot) *total_slot = carry_s;
    }
}

__global__ void __launch_bounds__(CMP_THREADS)
compact_scatter(const uint8_t *__restrict__ keep, const int16_t *__restrict__ dense, int n,
                int lane_rev, const int32_t *__restrict__ tile_offsets, int16_t *__restrict__ out) {
    __shared__ int warp_sums[CMP_THREADS / 32];
    const int base = blockIdx.x * CMP_TILE + threadIdx.x * CMP_PER_THREAD;
    int flags = 0, c = 0;
#pragma unroll
    for (int k = 0; k < CMP_PER_THREAD; ++k)
        if (base + k < n && gate(keep, base + k, lane_rev)) { flags |= 1 << k; ++c; }
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, d);
        if ((threadIdx.x & 31) >= d) inc += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = inc;
    __syncthreads();
    int pos = tile_offsets[blockIdx.x] + inc - c;
    for (int k = 0; k < (int)(threadIdx.x >> 5); ++k) pos += warp_sums[k];
#pragma unroll
    for (int k = 0; k < CMP_PER_THREAD; ++k) {
        if (flags & (1 << k)) {
            const int16_t *s = dense + 5 * (size_t)(base + k);
            int16_t *d = out + 5 * (size_t)pos;
            d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = s[3]; d[4] = s[4];
            ++pos;
        }
    }
}

// ---------------------------------------------------------------------------
// Concat of camera payloads whose record counts only exist on the device (-c compaction): what sendStitchToUnity
// (src/pcs-multicamera-client.cpp:385-395) does with the n_shorts readCloud hands it -- every downsample-th record of
// every camera in camera order behind the int32 byte count -- with the counts read from device memory.
struct CountedTable {
    const int16_t *src[MAX_CAMS];
    const int32_t *cnt_dev[MAX_CAMS];     // record count on the device, or NULL: cnt_fixed
    int32_t cnt_fixed[MAX_CAMS];
    int32_t n_cams, downsample;
};
__global__ void __launch_bounds__(256)
stitch_counted(CountedTable t, uint8_t *__restrict__ stitched, int32_t *__restrict__ total_out) {
    __shared__ int32_t off[MAX_CAMS + 1];
    if (threadIdx.x == 0) {
        int32_t run = 0;
        for (int c = 0; c < t.n_cams; ++c) {
            const int32_t n = t.cnt_dev[c] ? *t.cnt_dev[c] : t.cnt_fixed[c];
            off[c] = run;
            run += (n + t.downsample - 1) / t.downsample;     // for (j = 0; j < n_shorts; j += 5 * downsample), :388
        }
        off[t.n_cams] = run;
        if (blockIdx.x == 0) {
            *reinterpret_cast<int32_t *>(stitched) = run * 10;   // :394-395
            *total_out = run * 10;
        }
    }
    __syncthreads();
    int16_t *out = reinterpret_cast<int16_t *>(stitched + 4);
    const int total = off[t.n_cams];
    int c = 0;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < total; j += gridDim.x * blockDim.x) {
        while (c + 1 < t.n_cams && j >= off[c + 1]) ++c;
        while (c > 0 && j < off[c]) --c;
        const int16_t *r = t.src[c] + 5 * (size_t)(j - off[c]) * t.downsample;
#pragma unroll
        for (int q = 0; q < 5; ++q) out[5 * (size_t)j + q] = r[q];
    }
}

// ---------------------------------------------------------------------------
// Stitch side.  One table per launch, passed by value (fits the 4 KB parameter bank).
struct CamTable {
    const int16_t *src[MAX_CAMS];
    int32_t n_in[MAX_CAMS];        // records available per camera
    int32_t out_off[MAX_CAMS + 1]; // exclusive prefix of output records
    int32_t n_cams, downsample;
};
struct TfTable {
    float m[MAX_CAMS][12];
};

__device__ __forceinline__ int find_cam(const CamTable &t, int j) {
    int c = 0;
    while (c + 1 < t.n_cams && j >= t.out_off[c + 1]) ++c;
    return c;
}

// 32 output records per warp, staged in shared memory, 16-byte stores.
// stitched = [int32 bytes][records]; records start at stitched + 4.
template <bool PCL>
__global__ void __launch_bounds__(256)
stitch_kernel(const __grid_constant__ CamTable tab, const __grid_constant__ TfTable tfs,
              uint8_t *__restrict__ stitched, float4 *__restrict__ cloud32) {
    __shared__ __align__(16) uint16_t slabs[8][32 * 5];
    const int total = tab.out_off[tab.n_cams];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int warp_j0 = j - lane;
    if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<int32_t *>(stitched) = total * 10;
    if (warp_j0 >= total) return;
    if (j < total) {
        const int c = find_cam(tab, j);
        const int16_t *s = tab.src[c] + 5 * (size_t)(j - tab.out_off[c]) * tab.downsample;
        uint16_t r[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) r[k] = (uint16_t)__ldg(s + k);
        if (PCL) {
            // src/pcs-multicamera-optimized.cpp:237-242 then SPEC.md s2 then :255-259
            const float x = __fdiv_rn((float)(int16_t)r[0], 1000.0f);
            const float y = __fdiv_rn((float)(int16_t)r[1], 1000.0f);
            const float z = __fdiv_rn((float)(int16_t)r[2], 1000.0f);
            const float *m = tfs.m[c];
            const float ox = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z)), m[3]);
            const float oy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[4], x), __fmul_rn(m[5], y)), __fmul_rn(m[6], z)), m[7]);
            const float oz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m[8], x), __fmul_rn(m[9], y)), __fmul_rn(m[10], z)), m[11]);
            r[0] = (uint16_t)to_mm16(ox);
            r[1] = (uint16_t)to_mm16(oy);
            r[2] = (uint16_t)to_mm16(oz);
            r[4] &= 0xFF;  // b = buffer[4] & 0xFF (:242)
            if (cloud32) {
                const uint32_t bgra = 0xFF000000u | ((uint32_t)(r[3] & 0xFF) << 16) | (r[3] & 0xFF00u) | r[4];
                cloud32[2 * (size_t)j] = make_float4(ox, oy, oz, 1.0f);
                cloud32[2 * (size_t)j + 1] = make_float4(__uint_as_float(bgra), 0.f, 0.f, 0.f);
            }
        }
        uint16_t *d = slabs[warp] + lane * 5;
#pragma unroll
        for (int k = 0; k < 5; ++k) d[k] = r[k];
    }
    __syncwarp();
    const int valid = min(32, total - warp_j0);
    uint8_t *dst = stitched + 4 + (size_t)warp_j0 * 10;
    if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && valid == 32) {
        if (lane < 20) st_global_v4(dst + lane * 16, reinterpret_cast<const uint4 *>(slabs[warp])[lane]);
    } else {
        uint16_t *d16 = reinterpret_cast<uint16_t *>(dst);
        for (int k = lane; k < valid * 5; k += 32) d16[k] = slabs[warp][k];
    }
}

// ---------------------------------------------------------------------------
// Vectorised stitch for the common case (no decimation, whole octets per camera, 16-byte
// aligned source and destination): a warp moves 32 octets = 2560 contiguous bytes with
// coalesced 16-byte loads and stores; for the PCL path the records take a round trip through
// shared memory so that each lane can unpack / transform / repack its own eight records
// with packed fp32x2 arithmetic.  Same bits as stitch_kernel (tests compare both with the oracle).
struct VecTable {
    int32_t tile_off[MAX_CAMS + 1];   // exclusive prefix of warp tiles (32 octets) per camera
    float one;                        // opaque 1.0f (see pcs_k1_pipe.cuh on FFMA2 contraction)
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

template <bool PCL>
__global__ void __launch_bounds__(256)
stitch_vec(const __grid_constant__ CamTable tab, const __grid_constant__ TfTable tfs,
           const __grid_constant__ VecTable vt, uint8_t *__restrict__ stitched) {
    __shared__ uint4 slabs[8][32 * 5];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tile = blockIdx.x * 8 + warp;
    const int total = tab.out_off[tab.n_cams];
    if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<int32_t *>(stitched) = total * 10;
    if (tile >= vt.tile_off[tab.n_cams]) return;
    int c = 0;
    while (c + 1 < tab.n_cams && tile >= vt.tile_off[c + 1]) ++c;
    const int oct0 = (tile - vt.tile_off[c]) * 32;                 // first octet of this tile in camera c
    const int valid = min(32, tab.n_in[c] / 8 - oct0);             // octets in this tile
    const uint8_t *src = reinterpret_cast<const uint8_t *>(tab.src[c]) + (size_t)oct0 * 80;
    uint8_t *dst = stitched + 4 + ((size_t)tab.out_off[c] + (size_t)oct0 * 8) * 10;
    const int n16 = valid * 5;
    uint4 *slab = slabs[warp];
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (k * 32 + lane < n16) slab[k * 32 + lane] = ld_global_nc_v4(src + (size_t)(k * 32 + lane) * 16);
    __syncwarp();
    if (PCL && lane < valid) {
        uint32_t w[20];
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const uint4 v = slab[lane * 5 + k];
            w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
        const float *m = tfs.m[c];
        const float2 one = f2(vt.one, vt.one), k1000 = f2(1000.f, 1000.f), rk = f2(1.0f / 1000.0f, 1.0f / 1000.0f),
                     nk = f2(-1000.f, -1000.f);
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            // records A = 2p, B = 2p+1 in words w[5p .. 5p+4]
            const uint32_t w0 = w[5 * p], w1 = w[5 * p + 1], w2 = w[5 * p + 2], w3 = w[5 * p + 3], w4 = w[5 * p + 4];
            float2 c3[3];
            c3[0] = f2((float)(int16_t)(w0 & 0xFFFF), (float)(int16_t)(w2 >> 16));
            c3[1] = f2((float)(int16_t)(w0 >> 16), (float)(int16_t)(w3 & 0xFFFF));
            c3[2] = f2((float)(int16_t)(w1 & 0xFFFF), (float)(int16_t)(w3 >> 16));
            // src/pcs-multicamera-optimized.cpp:237-239: x = (float)s / 1000.0f -- division by a constant
            // with the correctly rounded reciprocal and two Markstein corrections (see pcs_k1_pipe.cuh)
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                const float2 u0 = __fmul2_rn(c3[a], rk);
                const float2 r0 = __ffma2_rn(nk, u0, c3[a]);
                const float2 u1 = __ffma2_rn(r0, rk, u0);
                const float2 r1 = __ffma2_rn(nk, u1, c3[a]);
                c3[a] = __ffma2_rn(r1, rk, u1);
            }
            uint32_t o[3][2];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                // oracle/SPEC.md s2: ((m0*x + m1*y) + m2*z) + m3, no contraction
                float2 acc = __ffma2_rn(__fmul2_rn(f2(m[4 * r], m[4 * r]), c3[0]), one,
                                        __fmul2_rn(f2(m[4 * r + 1], m[4 * r + 1]), c3[1]));
                acc = __ffma2_rn(acc, one, __fmul2_rn(f2(m[4 * r + 2], m[4 * r + 2]), c3[2]));
                acc = __ffma2_rn(acc, one, f2(m[4 * r + 3], m[4 * r + 3]));
                const float2 mm = __fmul2_rn(acc, k1000);          // :255-257
                o[r][0] = (uint32_t)__float2int_rz(mm.x);
                o[r][1] = (uint32_t)__float2int_rz(mm.y);
            }
            w[5 * p] = __byte_perm(o[0][0], o[1][0], 0x5410);
            w[5 * p + 1] = __byte_perm(o[2][0], w1, 0x7610);                    // keep rg of record A
            w[5 * p + 2] = __byte_perm(w2 & 0xFF, o[0][1], 0x5410);             // b & 0xFF (:242), x of record B
            w[5 * p + 3] = __byte_perm(o[1][1], o[2][1], 0x5410);
            w[5 * p + 4] = w4 & 0x00FFFFFFu;                                    // rg, b & 0xFF of record B
        }
#pragma unroll
        for (int k = 0; k < 5; ++k) slab[lane * 5 + k] = make_uint4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
    __syncwarp();
#pragma unroll
    for (int k = 0; k < 5; ++k)
        if (k * 32 + lane < n16) st_global_v4(dst + (size_t)(k * 32 + lane) * 16, slab[k * 32 + lane]);
}

// ---- PLY vertex rows -----------------------------------------------------------------------
// The reference's viewer path can dump the stitched pcl::PointXYZRGB cloud with
// pcl::io::savePLYFileBinary (src/pcs-multicamera-client.cpp:482-489).  A binary PLY vertex of
// that cloud is 15 bytes: float x, y, z + uchar red, green, blue.  One warp turns 32 of the 32-byte
// PCL points into 480 contiguous bytes through shared memory, so both sides move 16 bytes per lane.
constexpr int PLY_ROW = 15, PLY_THREADS = 256;

__global__ void __launch_bounds__(PLY_THREADS)
ply_rows(const uint4 *__restrict__ cloud32, int n, uint8_t *__restrict__ rows) {
    __shared__ __align__(16) uint8_t stage[PLY_THREADS / 32][32 * PLY_ROW];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_groups = (n + 31) / 32;
    for (int grp = blockIdx.x * (PLY_THREADS / 32) + warp; grp < n_groups; grp += gridDim.x * (PLY_THREADS / 32)) {
        const int i = grp * 32 + lane;
        uint8_t *mine = stage[warp] + lane * PLY_ROW;
        if (i < n) {
            const uint4 xyzw = __ldg(cloud32 + 2 * (size_t)i);
            const uint32_t bgra = __ldg(reinterpret_cast<const uint32_t *>(cloud32 + 2 * (size_t)i + 1));
            const uint32_t w[3] = {xyzw.x, xyzw.y, xyzw.z};
#pragma unroll
            for (int k = 0; k < 12; ++k) mine[k] = (uint8_t)(w[k >> 2] >> (8 * (k & 3)));
            mine[12] = (uint8_t)(bgra >> 16);     // PCL packs b, g, r, a
            mine[13] = (uint8_t)(bgra >> 8);
            mine[14] = (uint8_t)bgra;
        }
        __syncwarp();
        const int valid = min(32, n - grp * 32);
        uint8_t *dst = rows + (size_t)grp * 32 * PLY_ROW;
        if (valid == 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            if (lane < 30) reinterpret_cast<uint4 *>(dst)[lane] = reinterpret_cast<const uint4 *>(stage[warp])[lane];
        } else {
            for (int b = lane; b < valid * PLY_ROW; b += 32) dst[b] = stage[warp][b];
        }
        __syncwarp();
    }
}

}  // namespace pcs
