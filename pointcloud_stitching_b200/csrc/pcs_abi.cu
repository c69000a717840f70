// C ABI (include/pcs_b200.h) over the sm_100a kernels.  Host side in C++ like the
// reference; no CPU compute path: every entry point either launches a kernel or
// fails with a negative status.
#include "../../include/pcs_b200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

#include "pcs_kernels.cuh"
#include "pcs_k1_pipe.cuh"
#include "pcs_voxel.cuh"
#include "pcs_voxel_sweep.cuh"
#include "pcs_voxel_msd.cuh"
#include "pcs_exchange.cuh"

using namespace pcs;

// ---------------------------------------------------------------------------
namespace {

thread_local std::string g_tls_error;

struct StreamState {
    bool configured = false;
    unsigned geom_gen = 0;          // bumped whenever anything but `tf` changes (live batches freeze the rest)
    pcs_stream_desc desc{};
    StreamParams params{};
    cudaStream_t cs = nullptr;
    std::mutex mu;
    // ctx-owned device buffers for the host-pointer entry points
    uint16_t *d_z16 = nullptr;
    uint8_t *d_color = nullptr;
    int16_t *d_payload = nullptr;   // N records (compacted when cutoff)
    int16_t *d_dense = nullptr;     // (unused since the one-pass -c of k1_direct: no dense records any more)
    uint8_t *d_keep = nullptr;
    int32_t *d_tiles = nullptr;
    int32_t *d_count = nullptr;
    int32_t *h_count = nullptr;     // pinned
    DevJob *d_job = nullptr;
    size_t cap_pts = 0, cap_color = 0;
    int32_t *d_rowmap = nullptr;    // TEX_TRANSLATE_X row map (depth row -> colour row), see make_rowmap
    size_t cap_rowmap = 0;
    // vertices path
    float *d_xyz = nullptr, *d_uv = nullptr;
    int16_t *d_vpayload = nullptr;
    size_t cap_vtx = 0;
    // -c scratch of the vertices path
    int16_t *d_cdense = nullptr;
    uint8_t *d_ckeep = nullptr;
    int32_t *d_ctiles = nullptr;
    size_t cap_cut = 0;
    // outstanding begin()
    bool pending = false;
    int16_t *pending_buffer = nullptr;
    int pending_header = 0;
};

// One pipeline of pcs_b200_stitch_frames: device frames of every camera, the device-resident stitched
// buffer all cameras write into, and the cached batch (job table + launch plan) of that camera set.
struct StitchSlot {
    std::mutex mu;
    cudaStream_t cs = nullptr;
    std::vector<int> streams;                 // what the cached plan was built for
    std::vector<unsigned> gens;
    int downsample = 0;
    std::vector<uint16_t *> d_z;
    std::vector<uint8_t *> d_c;
    uint8_t *d_stitched = nullptr;            // [pad 12][int32][records of every camera, full rate]
    uint8_t *d_decimated = nullptr;           // [pad 12][int32][records, every downsample-th]  (downsample > 1)
    pcs_batch *batch = nullptr;               // all cameras in one launch (downsample > 1, or PCS_STITCH_PIPELINE=batch)
    std::vector<pcs_batch *> cam_batch;       // one launch per camera, on that camera's own stream
    std::vector<cudaStream_t> cam_cs;
    std::vector<long long> cam_off;           // first record of every camera in the stitched payload
    // -c streams: a camera compacts inside its own slot, the counts stay on the device until the concat has run
    int32_t *d_counts = nullptr, *d_total = nullptr, *h_total = nullptr;
    bool counted = false;
    uint8_t *pending_host = nullptr;
    long long total_pts = 0, out_bytes = 0;
    bool pending = false, per_camera = false;
};

}  // namespace

struct pcs_ctx {
    int device = 0;
    int max_streams = 0;
    int kernel_variant = 0;
    int voxel_variant = 0;
    int sm_count = 0;
    int plan_slot = 0;      // this context's slot of the voxel sort's constant-memory plan
    StreamState *streams = nullptr;
    StreamParams *d_params = nullptr;
    std::mutex mu;          // guards `error`
    std::mutex scratch_mu;  // guards the stitch / voxel scratch below
    std::string error;
    // stitch-side scratch for the host-pointer entry points
    uint8_t *d_stitch_in = nullptr, *d_stitch_out = nullptr;
    size_t cap_stitch_in = 0, cap_stitch_out = 0;
    VoxelScratch voxel;
    StitchSlot stitch_slots[PCS_B200_STITCH_SLOTS];
    // peer allocations mapped with pcs_b200_ipc_open (an allocation can be opened once per process)
    struct IpcMap { uint8_t handle[64]; void *base; int refs; };
    std::vector<IpcMap> ipc;
    std::mutex ipc_mu;
};

struct pcs_batch {
    std::vector<DevJob> jobs;          // host mirror
    DevJob *d_jobs = nullptr;
    struct Group {                      // one launch per kernel variant present
        int tex_mode, cutoff, floatout;
        int first, count;               // range in the (sorted) job table
        int max_tiles;
        int max_octets;
    };
    std::vector<Group> groups;
    struct CutJob { int job, n, lane_rev, n_tiles; int32_t *d_tiles; };
    std::vector<CutJob> cuts;       // -c jobs (their look-back words live in d_lb)
    void *d_lb = nullptr;           // look-back words of every -c job of the batch: one memset per run
    size_t lb_bytes = 0;
    std::vector<std::pair<int, unsigned>> gens;   // (stream, geom_gen) captured at create
    std::vector<void *> owned;          // scratch to free
    int launches = 0;
    PipeBatch pipe;                     // pipelined-kernel work list (variant 2)
    bool use_pipe = false;
};

namespace {

int fail(pcs_ctx *ctx, int status, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_tls_error = buf;
    if (ctx) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->error = buf;
    }
    return status;
}

// Every entry point runs on the context's device and puts the caller's current device back on
// return (single-process multi-GPU hosts and torch callers keep their own notion of "current").
struct DeviceGuard {
    int prev = -1;
    bool changed = false;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
        if (prev != device) changed = cudaSetDevice(device) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (changed && prev >= 0) cudaSetDevice(prev);
    }
    DeviceGuard(const DeviceGuard &) = delete;
    DeviceGuard &operator=(const DeviceGuard &) = delete;
};

#define CU(ctx, call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(ctx, PCS_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                    \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                          \
    } while (0)

bool is_identity3(const float *r) {
    static const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 9; ++i)
        if (r[i] != I[i]) return false;
    return true;
}

// Choose the cheapest tex-coordinate path that is still bit-exact (pcs_device.cuh).
int pick_tex_mode(const pcs_stream_desc &d) {
    // (librealsense applies inverse Brown-Conrady only when deprojecting and modified Brown-Conrady only when projecting;
    // the other way round the coefficients have no effect)
    if (d.depth.model == PCS_B200_DISTORTION_INVERSE_BROWN_CONRADY || d.color.model == PCS_B200_DISTORTION_MODIFIED_BROWN_CONRADY)
        return TEX_GENERAL;
    if (!is_identity3(d.d2c_rotation)) return TEX_GENERAL;
    const bool t0 = d.d2c_translation[0] == 0.f && d.d2c_translation[1] == 0.f &&
                    d.d2c_translation[2] == 0.f;
    const bool same = d.depth.width == d.color.width && d.depth.height == d.color.height &&
                      d.depth.fx == d.color.fx && d.depth.fy == d.color.fy &&
                      d.depth.ppx == d.color.ppx && d.depth.ppy == d.color.ppy;
    // the error bound behind TEX_ALIGNED (oracle/SPEC.md s1) needs ordinary magnitudes
    const bool sane = d.depth.width <= 4096 && d.depth.height <= 4096 && d.depth.fx >= 16.f &&
                      d.depth.fy >= 16.f && d.depth.fx <= 65536.f && d.depth.fy <= 65536.f &&
                      d.depth.ppx >= 0.f && d.depth.ppx <= (float)d.depth.width &&
                      d.depth.ppy >= 0.f && d.depth.ppy <= (float)d.depth.height &&
                      d.depth_scale >= 1e-6f && d.depth_scale <= 1.0f;
    if (t0 && same && sane) return TEX_ALIGNED;
    // translation along x only, same image size and vertical intrinsics: tap row == pixel row
    const bool same_v = d.depth.width == d.color.width && d.depth.height == d.color.height &&
                        d.depth.fy == d.color.fy && d.depth.ppy == d.color.ppy;
    if (same_v && sane && d.d2c_translation[1] == 0.f && d.d2c_translation[2] == 0.f) return TEX_TRANSLATE_X;
    return TEX_TRANSLATE;
}

// R = I and T = (tx, 0, 0): t1 = p1 and t2 = depth exactly, so the tap ROW of a valid pixel depends on its own row and
// (through rounding only) on the depth value:  py = fl(fl(fl(fl(depth * ny) / depth) * cfy) + cppy), tap =
// trunc(fma(fl(py / CH), CH, .5)).  In real arithmetic that is r(y) = (y - ppy) / fy * cfy + cppy; the float chain strays
// from it by < (|r - cppy| * 5 + |r| * 2) * 2^-24 + ulp(|r|) / 2  (~4.5e-4 px at r ~ 1000: oracle/SPEC.md s1 gives the same
// argument for equal intrinsics).  If r(y) + 0.5 keeps a distance of 1e-3 + |r| * 2e-6 (> 3 x that bound) from every
// integer for EVERY row, the tap row is floor(r(y) + 0.5) clamped, whatever the depth: a table.  Colour frames of another
// size (1280x720 depth + 1920x1080 colour, what the reference records: src/pcs-camera-grab-frames.cpp:69-70) then run
// through the row-exact kernel path instead of the windowed one.  tests/test_k1_gpu.py sweeps every z16 at every row.
bool make_rowmap(const pcs_stream_desc &d, std::vector<int32_t> &tab) {
    const int H = d.depth.height, CH = d.color.height;
    if (H < 1 || H > 8192 || CH < 1 || CH > 8192) return false;
    tab.resize(H);
    for (int y = 0; y < H; ++y) {
        const double r = ((double)y - (double)d.depth.ppy) / (double)d.depth.fy * (double)d.color.fy + (double)d.color.ppy;
        const double t = r + 0.5, k = std::floor(t);
        const double dist = std::min(t - k, k + 1.0 - t);
        if (!(dist >= 1e-3 + std::fabs(r) * 2e-6) || !(std::fabs(r) < 1e6)) return false;
        tab[y] = (int32_t)std::min<double>(std::max<double>(k, 0.0), (double)(CH - 1));
    }
    return true;
}

void digest(const pcs_stream_desc &d, StreamParams &p) {
    memset(&p, 0, sizeof p);
    p.W = d.depth.width; p.H = d.depth.height; p.N = p.W * p.H;
    p.CW = d.color.width; p.CH = d.color.height; p.bpp = d.color_bpp; p.stride = d.color_stride;
    p.tex_mode = pick_tex_mode(d);
    p.ppx = d.depth.ppx; p.ppy = d.depth.ppy; p.fx = d.depth.fx; p.fy = d.depth.fy;
    p.cppx = d.color.ppx; p.cppy = d.color.ppy; p.cfx = d.color.fx; p.cfy = d.color.fy;
    p.cwf = (float)d.color.width; p.chf = (float)d.color.height;
    p.depth_scale = d.depth_scale;
    memcpy(p.R, d.d2c_rotation, sizeof p.R);
    memcpy(p.T, d.d2c_translation, sizeof p.T);
    memcpy(p.tf, d.tf, sizeof p.tf);
    p.dmodel = d.depth.model == PCS_B200_DISTORTION_INVERSE_BROWN_CONRADY ? d.depth.model : 0;
    p.cmodel = d.color.model == PCS_B200_DISTORTION_MODIFIED_BROWN_CONRADY ? d.color.model : 0;
    memcpy(p.dcoef, d.depth.coeffs, sizeof p.dcoef);
    memcpy(p.ccoef, d.color.coeffs, sizeof p.ccoef);
    p.cutoff = d.cutoff != 0; p.lane_rev = d.cutoff_lane_reversed != 0;
    p.z_lo = d.z_lo; p.z_hi = d.z_hi; p.x_lo = d.x_lo; p.x_hi = d.x_hi;
}

int check_stream(pcs_ctx *ctx, int stream, bool need_depth) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    if (stream < 0 || stream >= ctx->max_streams)
        return fail(ctx, PCS_ERR_INVALID, "stream %d out of range [0,%d)", stream, ctx->max_streams);
    StreamState &s = ctx->streams[stream];
    if (!s.configured) return fail(ctx, PCS_ERR_INVALID, "stream %d not configured", stream);
    if (need_depth && (s.params.W % 8 != 0 || s.params.N <= 0))
        return fail(ctx, PCS_ERR_UNSUPPORTED, "depth width must be a positive multiple of 8");
    return PCS_OK;
}

template <class T> int grow(pcs_ctx *ctx, T *&p, size_t count) {
    if (p) cudaFree(p);
    p = nullptr;
    if (cudaMalloc(&p, count * sizeof(T) + 64) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, PCS_ERR_NOMEM, "cudaMalloc of %zu bytes failed", count * sizeof(T));
    }
    return PCS_OK;
}

int ensure_frame_buffers(pcs_ctx *ctx, StreamState &s) {
    const size_t n = (size_t)s.params.N, cb = (size_t)s.params.CH * s.params.stride;
    int rc;
    if (n > s.cap_pts) {
        s.cap_pts = 0;   // a failed grow leaves null pointers: never trust the old capacity afterwards
        if ((rc = grow(ctx, s.d_z16, n))) return rc;
        if ((rc = grow(ctx, s.d_payload, n * 5))) return rc;
        if ((rc = grow(ctx, s.d_keep, n + 64))) return rc;      // (-c: the look-back words of the one-pass compaction)
        if ((rc = grow(ctx, s.d_tiles, n / CMP_TILE + 2))) return rc;
        s.cap_pts = n;
    }
    if (cb > s.cap_color) {
        s.cap_color = 0;
        if ((rc = grow(ctx, s.d_color, cb))) return rc;
        s.cap_color = cb;
    }
    if (!s.d_count) {
        CU(ctx, cudaMalloc(&s.d_count, 64));
        CU(ctx, cudaHostAlloc(&s.h_count, 64, cudaHostAllocDefault));
        CU(ctx, cudaMalloc(&s.d_job, sizeof(DevJob)));
    }
    return PCS_OK;
}

// -c: dense records + keep flags -> compacted payload, count to d_count.
int launch_compaction(pcs_ctx *ctx, const uint8_t *keep, const int16_t *dense, int n, int lane_rev,
                      int32_t *tiles, int16_t *out, int32_t *count, cudaStream_t cs) {
    const int n_tiles = (n + CMP_TILE - 1) / CMP_TILE;
    compact_count<<<n_tiles, CMP_THREADS, 0, cs>>>(keep, n, lane_rev, tiles);
    compact_scan<<<1, 1024, 0, cs>>>(tiles, n_tiles, count, nullptr);
    compact_scatter<<<n_tiles, CMP_THREADS, 0, cs>>>(keep, dense, n, lane_rev, tiles, out);
    CU(ctx, cudaGetLastError());
    return 3;
}

template <int MODE> void launch_k1_mode(bool cutoff, bool floatout, dim3 grid, cudaStream_t cs,
                                        const DevJob *jobs, const StreamParams *params) {
    if (cutoff) {
        if (floatout) k1_direct<MODE, true, true><<<grid, K1_THREADS, 0, cs>>>(jobs, params);
        else k1_direct<MODE, true, false><<<grid, K1_THREADS, 0, cs>>>(jobs, params);
    } else {
        if (floatout) k1_direct<MODE, false, true><<<grid, K1_THREADS, 0, cs>>>(jobs, params);
        else k1_direct<MODE, false, false><<<grid, K1_THREADS, 0, cs>>>(jobs, params);
    }
}

void launch_k1_direct(int tex_mode, bool cutoff, bool floatout, dim3 grid, cudaStream_t cs,
                      const DevJob *jobs, const StreamParams *params) {
    switch (tex_mode) {
        case TEX_ALIGNED: launch_k1_mode<TEX_ALIGNED>(cutoff, floatout, grid, cs, jobs, params); break;
        case TEX_TRANSLATE: launch_k1_mode<TEX_TRANSLATE>(cutoff, floatout, grid, cs, jobs, params); break;
        case TEX_TRANSLATE_X: launch_k1_mode<TEX_TRANSLATE_X>(cutoff, floatout, grid, cs, jobs, params); break;
        default: launch_k1_mode<TEX_GENERAL>(cutoff, floatout, grid, cs, jobs, params); break;
    }
}

int tiles_for(int n_pts) { return (n_pts / 8 + K1_THREADS - 1) / K1_THREADS; }
// -c: persistent blocks per SM over all frames of a launch (tuning knob PCS_CUT_BLOCKS; any number is correct)
int cut_blocks_per_sm() { static const int v = pipe_knob("PCS_CUT_BLOCKS", 16, 1, 64); return v; }


}  // namespace

// ===========================================================================
extern "C" {

int pcs_b200_abi_version(void) { return PCS_B200_ABI_VERSION; }

const char *pcs_b200_status_string(int status) {
    switch (status) {
        case PCS_OK: return "ok";
        case PCS_ERR_INVALID: return "invalid argument";
        case PCS_ERR_CUDA: return "CUDA error";
        case PCS_ERR_NOMEM: return "out of memory";
        case PCS_ERR_UNSUPPORTED: return "unsupported";
        case PCS_ERR_CAPACITY: return "output buffer too small";
        default: return status >= 0 ? "ok" : "unknown error";
    }
}

// The returned pointer is a per-thread copy: other threads failing on the same context rewrite
// ctx->error under its lock, never the buffer handed out here.
const char *pcs_b200_last_error(const pcs_ctx *ctx) {
    thread_local std::string copy;
    if (!ctx) return g_tls_error.c_str();
    {
        std::lock_guard<std::mutex> lk(const_cast<pcs_ctx *>(ctx)->mu);
        copy = ctx->error;
    }
    return copy.c_str();
}

int pcs_b200_create(const pcs_config *cfg, pcs_ctx **out) {
    if (!cfg || !out) return fail(nullptr, PCS_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->max_streams < 1 || cfg->max_streams > 4096)
        return fail(nullptr, PCS_ERR_INVALID, "max_streams must be in [1,4096]");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, PCS_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
    }
    if (cfg->device < 0 || cfg->device >= ndev)
        return fail(nullptr, PCS_ERR_INVALID, "device %d out of range", cfg->device);
    DeviceGuard dg_(cfg->device);
    cudaDeviceProp prop;
    CU(nullptr, cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(nullptr, PCS_ERR_UNSUPPORTED, "built for sm_100a only; device is sm_%d%d",
                    prop.major, prop.minor);
    pcs_ctx *ctx = new pcs_ctx;
    ctx->device = cfg->device;
    ctx->max_streams = cfg->max_streams;
    ctx->kernel_variant = cfg->kernel_variant;
    ctx->voxel_variant = cfg->voxel_variant;
    ctx->sm_count = prop.multiProcessorCount;
    {
        static std::mutex slot_mu;
        static int next_slot = 0;
        std::lock_guard<std::mutex> lk(slot_mu);
        ctx->plan_slot = next_slot++ % SW_PLAN_SLOTS;
    }
    ctx->streams = new StreamState[cfg->max_streams];
    if (cudaMalloc(&ctx->d_params, sizeof(StreamParams) * cfg->max_streams) != cudaSuccess) {
        pcs_b200_destroy(ctx);
        return fail(nullptr, PCS_ERR_NOMEM, "cudaMalloc failed");
    }
    for (int i = 0; i < cfg->max_streams; ++i)
        if (cudaStreamCreateWithFlags(&ctx->streams[i].cs, cudaStreamNonBlocking) != cudaSuccess) {
            pcs_b200_destroy(ctx);
            return fail(nullptr, PCS_ERR_CUDA, "cudaStreamCreate failed");
        }
    int rc = pipe_configure(ctx->device);
    if (rc == PCS_OK && (sweep_configure<8, 256, 16>() != 0 || sweep_configure<10, 256, 16>() != 0 || vm_configure() != 0))
        rc = PCS_ERR_CUDA;
    if (rc != PCS_OK) {
        pcs_b200_destroy(ctx);
        return fail(nullptr, PCS_ERR_CUDA, "kernel attribute setup failed: %s",
                    cudaGetErrorString(cudaGetLastError()));
    }
    *out = ctx;
    return PCS_OK;
}

void pcs_b200_destroy(pcs_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard dg_(ctx->device);
    for (int i = 0; ctx->streams && i < ctx->max_streams; ++i) {
        StreamState &s = ctx->streams[i];
        if (s.cs) { cudaStreamSynchronize(s.cs); cudaStreamDestroy(s.cs); }
        cudaFree(s.d_z16); cudaFree(s.d_color); cudaFree(s.d_payload); cudaFree(s.d_dense);
        cudaFree(s.d_keep); cudaFree(s.d_tiles); cudaFree(s.d_count); cudaFree(s.d_job); cudaFree(s.d_rowmap);
        cudaFree(s.d_xyz); cudaFree(s.d_uv); cudaFree(s.d_vpayload);
        cudaFree(s.d_cdense); cudaFree(s.d_ckeep); cudaFree(s.d_ctiles);
        if (s.h_count) cudaFreeHost(s.h_count);
    }
    delete[] ctx->streams;
    cudaFree(ctx->d_params);
    cudaFree(ctx->d_stitch_in);
    cudaFree(ctx->d_stitch_out);
    voxel_free(ctx->voxel);
    for (auto &m : ctx->ipc) cudaIpcCloseMemHandle(m.base);
    for (StitchSlot &sl : ctx->stitch_slots) {
        if (sl.cs) { cudaStreamSynchronize(sl.cs); cudaStreamDestroy(sl.cs); }
        for (auto *p : sl.d_z) cudaFree(p);
        for (auto *p : sl.d_c) cudaFree(p);
        cudaFree(sl.d_stitched); cudaFree(sl.d_decimated);
        if (sl.batch) pcs_b200_batch_destroy(ctx, sl.batch);
        for (auto *b : sl.cam_batch) pcs_b200_batch_destroy(ctx, b);
        for (auto c : sl.cam_cs) { cudaStreamSynchronize(c); cudaStreamDestroy(c); }
        cudaFree(sl.d_counts); cudaFree(sl.d_total);
        if (sl.h_total) cudaFreeHost(sl.h_total);
    }
    delete ctx;
}

int pcs_b200_set_stream(pcs_ctx *ctx, int stream, const pcs_stream_desc *desc) {
    if (!ctx || !desc) return fail(ctx, PCS_ERR_INVALID, "null argument");
    if (stream < 0 || stream >= ctx->max_streams)
        return fail(ctx, PCS_ERR_INVALID, "stream %d out of range [0,%d)", stream, ctx->max_streams);
    const pcs_stream_desc &d = *desc;
    if (d.color.width < 1 || d.color.height < 1 || d.color_bpp < 3 ||
        d.color_stride < d.color.width * d.color_bpp)
        return fail(ctx, PCS_ERR_INVALID, "bad colour geometry (bpp >= 3, stride >= width*bpp)");
    if (d.depth.width < 0 || d.depth.height < 0 ||
        (long long)d.depth.width * d.depth.height > (1ll << 28))
        return fail(ctx, PCS_ERR_INVALID, "bad depth geometry");
    // rsutil.h (2.16): rs2_deproject_pixel_to_point undistorts for INVERSE_BROWN_CONRADY only, rs2_project_point_to_pixel
    // distorts for MODIFIED_BROWN_CONRADY only; a model on the other side of the chain (a D455 colour stream reports
    // inverse Brown-Conrady) or plain BROWN_CONRADY (a rectified image) is carried and ignored there, as here.  The
    // fisheye models (F-Theta 3, Kannala-Brandt 5) are not implemented: refused, not ignored.
    auto known = [](int m) { return m == PCS_B200_DISTORTION_NONE || m == PCS_B200_DISTORTION_MODIFIED_BROWN_CONRADY ||
                                    m == PCS_B200_DISTORTION_INVERSE_BROWN_CONRADY || m == PCS_B200_DISTORTION_BROWN_CONRADY; };
    if (!known(d.depth.model) || !known(d.color.model))
        return fail(ctx, PCS_ERR_UNSUPPORTED, "distortion models: none, (modified / inverse) Brown-Conrady; not F-Theta / Kannala-Brandt");
    DeviceGuard dg_(ctx->device);
    StreamState &s = ctx->streams[stream];
    std::lock_guard<std::mutex> lk(s.mu);
    if (s.configured) {
        // a live batch froze grid sizes, -c scratch and the pipelined kernel's calibration: only the
        // camera -> world transform may change under it
        pcs_stream_desc a = s.desc, b = d;
        memset(a.tf, 0, sizeof a.tf);
        memset(b.tf, 0, sizeof b.tf);
        if (memcmp(&a, &b, sizeof a) != 0) ++s.geom_gen;
    }
    s.desc = d;
    digest(d, s.params);
    // colour of another size / other vertical intrinsics behind a pure x baseline: row-exact through a proven row map
    if (s.params.tex_mode == TEX_TRANSLATE && d.d2c_translation[1] == 0.f && d.d2c_translation[2] == 0.f &&
        d.depth.width <= 4096 && d.depth.fx >= 16.f && d.depth.fy >= 16.f && d.depth.fx <= 65536.f && d.depth.fy <= 65536.f &&
        d.depth_scale >= 1e-6f && d.depth_scale <= 1.0f) {
        std::vector<int32_t> tab;
        if (make_rowmap(d, tab)) {
            if (tab.size() > s.cap_rowmap) {
                cudaFree(s.d_rowmap);
                s.d_rowmap = nullptr; s.cap_rowmap = 0;
                if (cudaMalloc(&s.d_rowmap, tab.size() * 4 + 64) != cudaSuccess) { cudaGetLastError(); return fail(ctx, PCS_ERR_NOMEM, "cudaMalloc failed"); }
                s.cap_rowmap = tab.size();
            }
            CU(ctx, cudaMemcpy(s.d_rowmap, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
            s.params.tex_mode = TEX_TRANSLATE_X;
            s.params.rowmap = s.d_rowmap;
        }
    }
    s.configured = true;
    CU(ctx, cudaMemcpy(ctx->d_params + stream, &s.params, sizeof(StreamParams), cudaMemcpyHostToDevice));
    return PCS_OK;
}

void *pcs_b200_host_alloc(pcs_ctx *ctx, size_t bytes) {
    void *p = nullptr;
    DeviceGuard dg_(ctx ? ctx->device : 0);
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        fail(ctx, PCS_ERR_NOMEM, "cudaHostAlloc of %zu bytes failed", bytes);
        return nullptr;
    }
    return p;
}

void pcs_b200_host_free(pcs_ctx *ctx, void *p) {
    DeviceGuard dg_(ctx ? ctx->device : 0);
    if (p) cudaFreeHost(p);
}

int pcs_b200_synchronize(pcs_ctx *ctx, void *cuda_stream) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    DeviceGuard dg_(ctx->device);
    CU(ctx, cudaStreamSynchronize((cudaStream_t)cuda_stream));
    return PCS_OK;
}

// ---- camera side, host buffers ---------------------------------------------
int pcs_b200_send_xyzrgb_begin(pcs_ctx *ctx, int stream, const uint16_t *z16_host,
                               const uint8_t *color_host, int16_t *buffer_host, int write_header) {
    int rc = check_stream(ctx, stream, true);
    if (rc) return rc;
    if (!z16_host || !color_host || !buffer_host) return fail(ctx, PCS_ERR_INVALID, "null buffer");
    StreamState &s = ctx->streams[stream];
    std::lock_guard<std::mutex> lk(s.mu);
    if (s.pending) return fail(ctx, PCS_ERR_INVALID, "stream %d already has a frame in flight", stream);
    const StreamParams &p = s.params;
    if ((size_t)p.N * 10 + 4 > (size_t)PCS_B200_CAMERA_BUF_SHORTS * 2)
        return fail(ctx, PCS_ERR_CAPACITY, "%d points do not fit the reference's 10 MB camera buffer", p.N);
    DeviceGuard dg_(ctx->device);
    if ((rc = ensure_frame_buffers(ctx, s))) return rc;
    DevJob job{};
    job.z16 = s.d_z16; job.color = s.d_color; job.payload = s.d_payload; job.xyzrgb = nullptr;
    job.count = s.d_count; job.keep = s.d_keep; job.dense = nullptr; job.stream = stream;
    CU(ctx, cudaMemcpyAsync(s.d_job, &job, sizeof job, cudaMemcpyHostToDevice, s.cs));
    CU(ctx, cudaMemcpyAsync(s.d_z16, z16_host, (size_t)p.N * 2, cudaMemcpyHostToDevice, s.cs));
    CU(ctx, cudaMemcpyAsync(s.d_color, color_host, (size_t)p.CH * p.stride, cudaMemcpyHostToDevice, s.cs));
    // (-c: persistent blocks that claim tiles by ticket -- any number of them is correct; 16 per SM measured best: 4 / 8 / 16 / 32 -> 0.42 / 0.37 / 0.31 / 0.30 ms per 64 frames)
    dim3 grid(p.cutoff ? std::min(tiles_for(p.N), std::max(1, ctx->sm_count * cut_blocks_per_sm())) : tiles_for(p.N), 1);
    if (p.cutoff) CU(ctx, cudaMemsetAsync(s.d_keep, 0, (size_t)(tiles_for(p.N) + 1) * 4, s.cs));    // look-back words
    launch_k1_direct(p.tex_mode, p.cutoff, false, grid, s.cs, s.d_job, ctx->d_params);
    CU(ctx, cudaGetLastError());
    uint8_t *dst = reinterpret_cast<uint8_t *>(buffer_host) + PCS_B200_HEADER_BYTES;
    if (p.cutoff) {
        CU(ctx, cudaMemcpyAsync(s.h_count, s.d_count, 4, cudaMemcpyDeviceToHost, s.cs));
        // the count is only known on the device: the records are copied out in end()
    } else {
        CU(ctx, cudaMemcpyAsync(dst, s.d_payload, (size_t)p.N * 10, cudaMemcpyDeviceToHost, s.cs));
    }
    s.pending = true;
    s.pending_buffer = buffer_host;
    s.pending_header = write_header;
    return PCS_OK;
}

int pcs_b200_send_xyzrgb_end(pcs_ctx *ctx, int stream) {
    int rc = check_stream(ctx, stream, true);
    if (rc) return rc;
    StreamState &s = ctx->streams[stream];
    std::lock_guard<std::mutex> lk(s.mu);
    if (!s.pending) return fail(ctx, PCS_ERR_INVALID, "stream %d has no frame in flight", stream);
    s.pending = false;
    DeviceGuard dg_(ctx->device);
    CU(ctx, cudaStreamSynchronize(s.cs));
    const StreamParams &p = s.params;
    const int count = p.cutoff ? *s.h_count : p.N;
    const int size = 5 * count * (int)sizeof(int16_t);              // :697
    uint8_t *b = reinterpret_cast<uint8_t *>(s.pending_buffer);
    if (p.cutoff && count > 0) {
        CU(ctx, cudaMemcpyAsync(b + PCS_B200_HEADER_BYTES, s.d_payload, (size_t)size, cudaMemcpyDeviceToHost, s.cs));
        CU(ctx, cudaStreamSynchronize(s.cs));
    }
    // :673 memset(buffer, 0, BUF_SIZE) happens before packing: whatever the pack did not
    // overwrite inside the first 5 000 000 bytes is zero afterwards.
    const size_t zero_end = PCS_B200_CAMERA_BUF_SHORTS;             // bytes
    const size_t used_end = PCS_B200_HEADER_BYTES + (size_t)size;
    if (used_end < zero_end) memset(b + used_end, 0, zero_end - used_end);
    memset(b, 0, PCS_B200_HEADER_BYTES);
    if (s.pending_header) memcpy(b, &size, sizeof(int));            // :715-718
    return size;
}

int pcs_b200_send_xyzrgb(pcs_ctx *ctx, int stream, const uint16_t *z16_host,
                         const uint8_t *color_host, int16_t *buffer_host, int write_header) {
    int rc = pcs_b200_send_xyzrgb_begin(ctx, stream, z16_host, color_host, buffer_host, write_header);
    if (rc) return rc;
    return pcs_b200_send_xyzrgb_end(ctx, stream);
}

int pcs_b200_pack_from_vertices_dev(pcs_ctx *ctx, int stream, const float *xyz_dev,
                                    const float *uv_dev, int n, const uint8_t *color_dev,
                                    int16_t *payload_dev, int32_t *count_dev, void *cuda_stream) {
    int rc = check_stream(ctx, stream, false);
    if (rc) return rc;
    if (n < 0 || (n & 3)) return fail(ctx, PCS_ERR_INVALID, "n must be a non-negative multiple of 4");
    if (n == 0) return 0;
    if (!xyz_dev || !uv_dev || !color_dev || !payload_dev) return fail(ctx, PCS_ERR_INVALID, "null buffer");
    if (reinterpret_cast<uintptr_t>(color_dev) & 3)
        return fail(ctx, PCS_ERR_INVALID, "colour frame must be 4-byte aligned");
    StreamState &s = ctx->streams[stream];
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    DeviceGuard dg_(ctx->device);
    const int blocks = (n / 4 + K1A_THREADS - 1) / K1A_THREADS;
    if (!s.params.cutoff) {
        k1a_vertices<false><<<blocks, K1A_THREADS, 0, cs>>>(xyz_dev, uv_dev, n, color_dev, ctx->d_params,
                                                    stream, payload_dev, nullptr);
        CU(ctx, cudaGetLastError());
        return n;
    }
    // -c: the record count only exists on the device after the compaction; the caller must say where
    // it goes (the return value below is the upper bound n, not the count)
    if (!count_dev)
        return fail(ctx, PCS_ERR_INVALID, "stream %d has cutoff set: count_dev must not be NULL", stream);
    // -c needs scratch: ctx-owned, sized for n
    std::lock_guard<std::mutex> lk(s.mu);
    if ((size_t)n > s.cap_cut) {
        s.cap_cut = 0;
        if ((rc = grow(ctx, s.d_cdense, (size_t)n * 5))) return rc;
        if ((rc = grow(ctx, s.d_ckeep, (size_t)n))) return rc;
        if ((rc = grow(ctx, s.d_ctiles, (size_t)n / CMP_TILE + 2))) return rc;
        s.cap_cut = n;
    }
    k1a_vertices<true><<<blocks, K1A_THREADS, 0, cs>>>(xyz_dev, uv_dev, n, color_dev, ctx->d_params, stream,
                                               s.d_cdense, s.d_ckeep);
    CU(ctx, cudaGetLastError());
    rc = launch_compaction(ctx, s.d_ckeep, s.d_cdense, n, s.params.lane_rev, s.d_ctiles, payload_dev,
                           count_dev, cs);
    if (rc < 0) return rc;
    return n;  // upper bound; the exact count is in *count_dev
}

int pcs_b200_pack_from_vertices(pcs_ctx *ctx, int stream, const float *xyz_host,
                                const float *uv_host, int n, const uint8_t *color_host,
                                int16_t *payload_host) {
    int rc = check_stream(ctx, stream, false);
    if (rc) return rc;
    if (n < 0 || (n & 3)) return fail(ctx, PCS_ERR_INVALID, "n must be a non-negative multiple of 4");
    if (n == 0) return 0;
    if (!xyz_host || !uv_host || !color_host || !payload_host) return fail(ctx, PCS_ERR_INVALID, "null buffer");
    StreamState &s = ctx->streams[stream];
    DeviceGuard dg_(ctx->device);
    const size_t cb = (size_t)s.params.CH * s.params.stride;
    {
        std::lock_guard<std::mutex> lk(s.mu);
        if ((size_t)n > s.cap_vtx) {
            s.cap_vtx = 0;
            if ((rc = grow(ctx, s.d_xyz, (size_t)n * 3))) return rc;
            if ((rc = grow(ctx, s.d_uv, (size_t)n * 2))) return rc;
            if ((rc = grow(ctx, s.d_vpayload, (size_t)n * 5))) return rc;
            s.cap_vtx = n;
        }
        if (cb > s.cap_color) {
            s.cap_color = 0;
            if ((rc = grow(ctx, s.d_color, cb))) return rc;
            s.cap_color = cb;
        }
        if (!s.d_count) {
            CU(ctx, cudaMalloc(&s.d_count, 64));
            CU(ctx, cudaHostAlloc(&s.h_count, 64, cudaHostAllocDefault));
            CU(ctx, cudaMalloc(&s.d_job, sizeof(DevJob)));
        }
    }
    CU(ctx, cudaMemcpyAsync(s.d_xyz, xyz_host, (size_t)n * 12, cudaMemcpyHostToDevice, s.cs));
    CU(ctx, cudaMemcpyAsync(s.d_uv, uv_host, (size_t)n * 8, cudaMemcpyHostToDevice, s.cs));
    CU(ctx, cudaMemcpyAsync(s.d_color, color_host, cb, cudaMemcpyHostToDevice, s.cs));
    rc = pcs_b200_pack_from_vertices_dev(ctx, stream, s.d_xyz, s.d_uv, n, s.d_color, s.d_vpayload,
                                         s.d_count, s.cs);
    if (rc < 0) return rc;
    int count = n;
    if (s.params.cutoff) {
        CU(ctx, cudaMemcpyAsync(s.h_count, s.d_count, 4, cudaMemcpyDeviceToHost, s.cs));
        CU(ctx, cudaStreamSynchronize(s.cs));
        count = *s.h_count;
    }
    CU(ctx, cudaMemcpyAsync(payload_host, s.d_vpayload, (size_t)count * 10, cudaMemcpyDeviceToHost, s.cs));
    CU(ctx, cudaStreamSynchronize(s.cs));
    return count;
}

// ---- camera side, batched ----------------------------------------------------
static int batch_create_impl(pcs_ctx *ctx, const pcs_frame_job *jobs, int n_jobs, int n_peers,
                             const long long *peer_delta, pcs_batch **out) {
    if (!ctx || !jobs || !out || n_jobs < 1 || n_jobs > 65535)
        return fail(ctx, PCS_ERR_INVALID, "bad batch arguments (1 <= n_jobs <= 65535)");
    *out = nullptr;
    DeviceGuard dg_(ctx->device);
    pcs_batch *b = new pcs_batch;
    auto bail = [&](int rc) { pcs_b200_batch_destroy(ctx, b); return rc; };
    // order jobs by kernel variant so that each variant is one launch
    std::vector<int> order;
    for (int v = 0; v < 16; ++v) {
        const int tm = v % 4, cut = (v / 4) & 1, fo = v / 8;
        pcs_batch::Group g{tm, cut, fo, (int)order.size(), 0, 0, 0};
        for (int j = 0; j < n_jobs; ++j) {
            int rc = check_stream(ctx, jobs[j].stream, true);
            if (rc) return bail(rc);
            const StreamParams &p = ctx->streams[jobs[j].stream].params;
            if (p.tex_mode == tm && p.cutoff == cut && (jobs[j].xyzrgb_dev != nullptr) == (fo != 0)) {
                order.push_back(j);
                g.count++;
                g.max_tiles = std::max(g.max_tiles, tiles_for(p.N));
                g.max_octets = std::max(g.max_octets, p.N / 8);
            }
        }
        if (g.count) b->groups.push_back(g);
    }
    for (int j : order) {
        const pcs_frame_job &in = jobs[j];
        const StreamParams &p = ctx->streams[in.stream].params;
        if (!in.z16_dev || !in.color_dev || !in.payload_dev) return bail(fail(ctx, PCS_ERR_INVALID, "job %d: null buffer", j));
        if ((reinterpret_cast<uintptr_t>(in.z16_dev) & 15) || (reinterpret_cast<uintptr_t>(in.color_dev) & 15))
            return bail(fail(ctx, PCS_ERR_INVALID, "job %d: frames must be 16-byte aligned", j));
        if (in.xyzrgb_dev && (reinterpret_cast<uintptr_t>(in.xyzrgb_dev) & 15))
            return bail(fail(ctx, PCS_ERR_INVALID, "job %d: xyzrgb must be 16-byte aligned", j));
        DevJob d{};
        d.z16 = in.z16_dev; d.color = in.color_dev; d.payload = in.payload_dev;
        d.xyzrgb = in.xyzrgb_dev; d.count = in.count_dev; d.stream = in.stream;
        if (p.cutoff) {
            // look-back words of the one-pass compaction (k1_direct<CUTOFF>): [ticket][one per tile], 16-byte slices of
            // one allocation for the whole batch (offset kept in d_tiles until the buffer exists)
            const int n_tiles = tiles_for(p.N);
            b->cuts.push_back({(int)b->jobs.size(), p.N, p.lane_rev, n_tiles, reinterpret_cast<int32_t *>(b->lb_bytes)});
            b->lb_bytes += ((size_t)(n_tiles + 1) * 4 + 15) & ~(size_t)15;
        }
        b->jobs.push_back(d);
        b->gens.emplace_back(in.stream, ctx->streams[in.stream].geom_gen);
    }
    if (b->lb_bytes) {
        if (cudaMalloc(&b->d_lb, b->lb_bytes) != cudaSuccess) {
            cudaGetLastError();
            return bail(fail(ctx, PCS_ERR_NOMEM, "cutoff scratch allocation failed"));
        }
        b->owned.push_back(b->d_lb);
        for (auto &c : b->cuts) {
            c.d_tiles = reinterpret_cast<int32_t *>(static_cast<uint8_t *>(b->d_lb) + reinterpret_cast<size_t>(c.d_tiles));
            b->jobs[c.job].keep = reinterpret_cast<uint8_t *>(c.d_tiles);
        }
    }
    if (cudaMalloc(&b->d_jobs, sizeof(DevJob) * b->jobs.size()) != cudaSuccess) {
        cudaGetLastError();
        return bail(fail(ctx, PCS_ERR_NOMEM, "job table allocation failed"));
    }
    if (cudaMemcpy(b->d_jobs, b->jobs.data(), sizeof(DevJob) * b->jobs.size(), cudaMemcpyHostToDevice) != cudaSuccess)
        return bail(fail(ctx, PCS_ERR_CUDA, "job table upload failed: %s", cudaGetErrorString(cudaGetLastError())));
    // pipelined variant: every job must be plain (no cutoff, no float output) and 16-B aligned
    b->use_pipe = false;
    if (ctx->kernel_variant != 1) {
        bool ok = true;
        for (size_t i = 0; i < b->jobs.size(); ++i) {
            const StreamParams &p = ctx->streams[b->jobs[i].stream].params;
            ok = ok && pipe_supports(p) && !b->jobs[i].xyzrgb &&
                 !(reinterpret_cast<uintptr_t>(b->jobs[i].payload) & 15);
        }
        if (ok) {
            std::vector<StreamParams> sp(ctx->max_streams);
            for (int i = 0; i < ctx->max_streams; ++i) sp[i] = ctx->streams[i].params;
            bool remote = false;
            for (int j = 0; j < n_jobs; ++j) remote = remote || (jobs[j].reserved & PCS_B200_JOB_REMOTE_FRAME) != 0;
            int rc = pipe_build(b->pipe, b->jobs, sp, ctx->sm_count, n_peers, peer_delta, remote);
            if (rc == PCS_OK) b->use_pipe = true;
            else if (ctx->kernel_variant == 2 || n_peers)
                return bail(fail(ctx, rc, "pipelined kernel setup failed"));
        } else if (ctx->kernel_variant == 2 || n_peers) {
            return bail(fail(ctx, PCS_ERR_UNSUPPORTED,
                             "kernel_variant=2 needs plain jobs (no cutoff / float output, aligned payloads)"));
        }
    }
    if (n_peers && !b->use_pipe)
        return bail(fail(ctx, PCS_ERR_UNSUPPORTED, "the fused exchange needs the pipelined kernel (kernel_variant != 1)"));
    b->launches = b->use_pipe ? pipe_launches(b->pipe) : (int)b->groups.size();
    *out = b;
    return PCS_OK;
}

int pcs_b200_batch_create(pcs_ctx *ctx, const pcs_frame_job *jobs, int n_jobs, pcs_batch **out) {
    return batch_create_impl(ctx, jobs, n_jobs, 0, nullptr, out);
}

int pcs_b200_batch_create_fanout(pcs_ctx *ctx, const pcs_frame_job *jobs, int n_jobs,
                                 const void *local_base, size_t local_bytes,
                                 const void *const *peer_bases, int n_peers, pcs_batch **out) {
    if (!ctx || !jobs || !out || !local_base || n_peers < 0 || n_peers > PIPE_MAX_PEERS || (n_peers && !peer_bases))
        return fail(ctx, PCS_ERR_INVALID, "bad fan-out arguments (at most %d peers)", PIPE_MAX_PEERS);
    long long delta[PIPE_MAX_PEERS] = {0};
    const uintptr_t lo = reinterpret_cast<uintptr_t>(local_base);
    for (int p = 0; p < n_peers; ++p) {
        if (!peer_bases[p] || (reinterpret_cast<uintptr_t>(peer_bases[p]) & 15) != (lo & 15))
            return fail(ctx, PCS_ERR_INVALID, "peer %d: mirror base must share the local base's 16-byte alignment", p);
        delta[p] = (long long)(reinterpret_cast<uintptr_t>(peer_bases[p]) - lo);
    }
    for (int j = 0; j < n_jobs; ++j) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(jobs[j].payload_dev);
        int rc = check_stream(ctx, jobs[j].stream, true);
        if (rc) return rc;
        const size_t bytes = (size_t)ctx->streams[jobs[j].stream].params.N * 10;
        if (a < lo || a + bytes > lo + local_bytes)
            return fail(ctx, PCS_ERR_INVALID, "job %d: payload lies outside the mirrored buffer", j);
    }
    return batch_create_impl(ctx, jobs, n_jobs, n_peers, delta, out);
}

// Lets this context's kernels read (pull exchange) and write (fan-out) memory that was cudaMalloc'ed
// on another device of the same process.  Across processes the mapping comes from CUDA IPC / VMM /
// symmetric memory instead and this call is not needed.
int pcs_b200_enable_peer(pcs_ctx *ctx, int peer_device) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    if (peer_device == ctx->device) return PCS_OK;
    DeviceGuard dg_(ctx->device);
    int can = 0;
    CU(ctx, cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
    if (!can) return fail(ctx, PCS_ERR_UNSUPPORTED, "device %d cannot access device %d", ctx->device, peer_device);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return PCS_OK;
    }
    CU(ctx, e);
    return PCS_OK;
}

// ---- one process per GPU: CUDA IPC --------------------------------------------------------------
int pcs_b200_ipc_export(pcs_ctx *ctx, const void *dev_ptr, pcs_ipc_handle *out) {
    if (!ctx || !dev_ptr || !out) return fail(ctx, PCS_ERR_INVALID, "null argument");
    DeviceGuard dg_(ctx->device);
    // the handle names the whole allocation: find its base (driver API through the runtime's entry point table)
    typedef int (*get_range_t)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
        cudaGetLastError();
        return fail(ctx, PCS_ERR_CUDA, "cuMemGetAddressRange is not available");
    }
    unsigned long long base = 0;
    size_t bytes = 0;
    if (reinterpret_cast<get_range_t>(fn)(&base, &bytes, (unsigned long long)reinterpret_cast<uintptr_t>(dev_ptr)) != 0)
        return fail(ctx, PCS_ERR_INVALID, "not a device allocation of this process");
    cudaIpcMemHandle_t h;
    CU(ctx, cudaIpcGetMemHandle(&h, reinterpret_cast<void *>(static_cast<uintptr_t>(base))));
    static_assert(sizeof h == sizeof out->reserved, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out->reserved, &h, sizeof h);
    out->offset = (uint64_t)(reinterpret_cast<uintptr_t>(dev_ptr) - (uintptr_t)base);
    out->device = (uint64_t)ctx->device;
    return PCS_OK;
}

int pcs_b200_ipc_open(pcs_ctx *ctx, const pcs_ipc_handle *handle, void **dev_ptr_out) {
    if (!ctx || !handle || !dev_ptr_out) return fail(ctx, PCS_ERR_INVALID, "null argument");
    *dev_ptr_out = nullptr;
    DeviceGuard dg_(ctx->device);
    std::lock_guard<std::mutex> lk(ctx->ipc_mu);
    for (auto &m : ctx->ipc)
        if (memcmp(m.handle, handle->reserved, 64) == 0) {
            ++m.refs;
            *dev_ptr_out = static_cast<uint8_t *>(m.base) + handle->offset;
            return PCS_OK;
        }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle->reserved, sizeof h);
    void *base = nullptr;
    CU(ctx, cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    pcs_ctx::IpcMap m;
    memcpy(m.handle, handle->reserved, 64);
    m.base = base;
    m.refs = 1;
    ctx->ipc.push_back(m);
    *dev_ptr_out = static_cast<uint8_t *>(base) + handle->offset;
    return PCS_OK;
}

int pcs_b200_ipc_close(pcs_ctx *ctx, void *dev_ptr) {
    if (!ctx || !dev_ptr) return fail(ctx, PCS_ERR_INVALID, "null argument");
    DeviceGuard dg_(ctx->device);
    std::lock_guard<std::mutex> lk(ctx->ipc_mu);
    // the mapping that contains dev_ptr: the one with the largest base not above it
    int best = -1;
    for (size_t i = 0; i < ctx->ipc.size(); ++i)
        if (ctx->ipc[i].base <= dev_ptr && (best < 0 || ctx->ipc[i].base > ctx->ipc[best].base)) best = (int)i;
    if (best < 0) return fail(ctx, PCS_ERR_INVALID, "pointer was not returned by pcs_b200_ipc_open");
    if (--ctx->ipc[best].refs == 0) {
        CU(ctx, cudaIpcCloseMemHandle(ctx->ipc[best].base));
        ctx->ipc.erase(ctx->ipc.begin() + best);
    }
    return PCS_OK;
}

int pcs_b200_batch_run(pcs_ctx *ctx, pcs_batch *b, void *cuda_stream) {
    if (!ctx || !b) return fail(ctx, PCS_ERR_INVALID, "null argument");
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    DeviceGuard dg_(ctx->device);
    for (const auto &sg : b->gens)
        if (ctx->streams[sg.first].geom_gen != sg.second)
            return fail(ctx, PCS_ERR_INVALID,
                        "stream %d was reconfigured after this batch was created (only tf may change under a live "
                        "batch): destroy the batch and create it again", sg.first);
    if (b->use_pipe) {
        pipe_launch(b->pipe, b->d_jobs, ctx->d_params, cs,
                    [](void *c, int stream) -> const float * { return static_cast<pcs_ctx *>(c)->streams[stream].params.tf; }, ctx);
        CU(ctx, cudaGetLastError());
        return PCS_OK;
    }
    if (b->d_lb) CU(ctx, cudaMemsetAsync(b->d_lb, 0, b->lb_bytes, cs));      // tickets and tile words of the -c jobs
    for (const auto &g : b->groups) {
        // (-c: persistent blocks that claim their frame's tiles by ticket; any number per frame is correct)
        dim3 grid(g.cutoff ? std::min(g.max_tiles, std::max(1, ctx->sm_count * cut_blocks_per_sm() / std::max(1, g.count))) : g.max_tiles, g.count);
        launch_k1_direct(g.tex_mode, g.cutoff != 0, g.floatout != 0, grid, cs, b->d_jobs + g.first,
                         ctx->d_params);
    }
    CU(ctx, cudaGetLastError());
    return PCS_OK;
}

int pcs_b200_batch_launches(const pcs_batch *b) { return b ? b->launches : 0; }

void pcs_b200_batch_destroy(pcs_ctx *ctx, pcs_batch *b) {
    if (!b) return;
    DeviceGuard dg_(ctx ? ctx->device : 0);
    for (void *p : b->owned) cudaFree(p);
    cudaFree(b->d_jobs);
    pipe_free(b->pipe);
    delete b;
}

// ---- stitch side ---------------------------------------------------------------
static int stitch_common(pcs_ctx *ctx, const int16_t *const *payload_dev, const int32_t *n_shorts,
                         int n_cams, int downsample, const float *transforms, bool pcl,
                         uint8_t *stitched_dev, size_t cap, void *cloud32_dev, cudaStream_t cs) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    if (!payload_dev || !n_shorts || !stitched_dev || n_cams < 0 || downsample < 1 || (pcl && !transforms))
        return fail(ctx, PCS_ERR_INVALID, "bad stitch arguments");
    if (n_cams > MAX_CAMS) return fail(ctx, PCS_ERR_UNSUPPORTED, "at most %d cameras per call", MAX_CAMS);
    if (reinterpret_cast<uintptr_t>(stitched_dev) & 3)
        return fail(ctx, PCS_ERR_INVALID, "stitched buffer must be 4-byte aligned");
    CamTable tab{};
    TfTable tfs{};
    tab.n_cams = n_cams;
    tab.downsample = downsample;
    long long total = 0;
    for (int i = 0; i < n_cams; ++i) {
        if (n_shorts[i] < 0 || (n_shorts[i] && !payload_dev[i])) return fail(ctx, PCS_ERR_INVALID, "camera %d: bad payload", i);
        const int n_in = n_shorts[i] / 5;
        // raw: for (j = 0; j < n_shorts; j += 5*ds)  -> ceil(n_shorts / (5*ds)) records
        //      (src/pcs-multicamera-client.cpp:388); a trailing partial record is read as the
        //      reference would, so n_shorts must be a multiple of 5 here.
        // pcl: width = size / ds (src/pcs-multicamera-optimized.cpp:230)
        if (n_shorts[i] % 5) return fail(ctx, PCS_ERR_INVALID, "camera %d: payload is not whole records", i);
        const int n_out = pcl ? n_in / downsample : (n_in + downsample - 1) / downsample;
        tab.src[i] = payload_dev[i];
        tab.n_in[i] = n_in;
        tab.out_off[i] = (int)total;
        total += n_out;
        if (pcl) memcpy(tfs.m[i], transforms + 16 * i, 12 * sizeof(float));
    }
    if (total * 10 > 0x7fffffffll) return fail(ctx, PCS_ERR_CAPACITY, "stitched payload exceeds int32");
    tab.out_off[n_cams] = (int)total;
    if ((size_t)total * 10 + 4 > cap) return fail(ctx, PCS_ERR_CAPACITY, "stitched buffer too small: need %lld bytes", total * 10 + 4);
    DeviceGuard dg_(ctx->device);
    // vectorised path: no decimation, whole octets per camera, 16-byte aligned everywhere, and
    // (PCL) coordinates that stay far inside int32 so that cvt.rzi == x86 cvttss2si
    bool vec = downsample == 1 && !cloud32_dev && total > 0 &&
               ((reinterpret_cast<uintptr_t>(stitched_dev) + 4) & 15) == 0;
    VecTable vt{};
    vt.one = 1.0f;
    for (int i = 0; i < n_cams && vec; ++i) {
        vec = tab.n_in[i] % 8 == 0 && (reinterpret_cast<uintptr_t>(tab.src[i]) & 15) == 0;
        vt.tile_off[i + 1] = vt.tile_off[i] + (tab.n_in[i] / 8 + 31) / 32;
        if (pcl)
            for (int r = 0; r < 3 && vec; ++r) {
                const float *m = tfs.m[i] + 4 * r;
                vec = (std::fabs(m[0]) + std::fabs(m[1]) + std::fabs(m[2])) * 32.768f + std::fabs(m[3]) < 2.0e6f;
            }
    }
    if (vec) {
        const int blocks = (vt.tile_off[n_cams] + 7) / 8;
        if (pcl) stitch_vec<true><<<blocks, 256, 0, cs>>>(tab, tfs, vt, stitched_dev);
        else stitch_vec<false><<<blocks, 256, 0, cs>>>(tab, tfs, vt, stitched_dev);
    } else {
        const int blocks = std::max(1, (int)((total + 255) / 256));
        if (pcl) stitch_kernel<true><<<blocks, 256, 0, cs>>>(tab, tfs, stitched_dev, (float4 *)cloud32_dev);
        else stitch_kernel<false><<<blocks, 256, 0, cs>>>(tab, tfs, stitched_dev, nullptr);
    }
    CU(ctx, cudaGetLastError());
    return (int)(total * 10);
}

int pcs_b200_stitch_raw_dev(pcs_ctx *ctx, const int16_t *const *payload_dev, const int32_t *n_shorts,
                            int n_cams, int downsample, uint8_t *stitched_dev, size_t stitched_cap,
                            void *cuda_stream) {
    return stitch_common(ctx, payload_dev, n_shorts, n_cams, downsample, nullptr, false, stitched_dev,
                         stitched_cap, nullptr, (cudaStream_t)cuda_stream);
}

int pcs_b200_stitch_pcl_dev(pcs_ctx *ctx, const int16_t *const *payload_dev, const int32_t *n_shorts,
                            int n_cams, int downsample, const float *transforms,
                            uint8_t *stitched_dev, size_t stitched_cap, void *cloud32_dev,
                            void *cuda_stream) {
    if (cloud32_dev && (reinterpret_cast<uintptr_t>(cloud32_dev) & 15))
        return fail(ctx, PCS_ERR_INVALID, "cloud32 must be 16-byte aligned");
    return stitch_common(ctx, payload_dev, n_shorts, n_cams, downsample, transforms, true, stitched_dev,
                         stitched_cap, cloud32_dev, (cudaStream_t)cuda_stream);
}

static int stitch_host(pcs_ctx *ctx, const int16_t *const *payload_host, const int32_t *n_shorts,
                       int n_cams, int downsample, const float *transforms, bool pcl,
                       uint8_t *stitched_host, size_t cap) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    if (!payload_host || !n_shorts || !stitched_host || n_cams < 0 || n_cams > MAX_CAMS)
        return fail(ctx, PCS_ERR_INVALID, "bad stitch arguments");
    DeviceGuard dg_(ctx->device);
    std::lock_guard<std::mutex> lk(ctx->scratch_mu);
    size_t in_bytes = 0;
    std::vector<size_t> off(n_cams);
    for (int i = 0; i < n_cams; ++i) {
        if (n_shorts[i] < 0) return fail(ctx, PCS_ERR_INVALID, "negative payload length");
        off[i] = in_bytes;
        in_bytes += ((size_t)n_shorts[i] * 2 + 15) & ~(size_t)15;
    }
    const size_t out_bytes = in_bytes + 64;
    if (in_bytes + 64 > ctx->cap_stitch_in) {
        cudaFree(ctx->d_stitch_in);
        ctx->d_stitch_in = nullptr; ctx->cap_stitch_in = 0;
        if (cudaMalloc(&ctx->d_stitch_in, in_bytes + 64) != cudaSuccess) { cudaGetLastError(); return fail(ctx, PCS_ERR_NOMEM, "cudaMalloc failed"); }
        ctx->cap_stitch_in = in_bytes + 64;
    }
    if (out_bytes > ctx->cap_stitch_out) {
        cudaFree(ctx->d_stitch_out);
        ctx->d_stitch_out = nullptr; ctx->cap_stitch_out = 0;
        if (cudaMalloc(&ctx->d_stitch_out, out_bytes) != cudaSuccess) { cudaGetLastError(); return fail(ctx, PCS_ERR_NOMEM, "cudaMalloc failed"); }
        ctx->cap_stitch_out = out_bytes;
    }
    cudaStream_t cs = ctx->streams[0].cs;
    std::vector<const int16_t *> dev(n_cams);
    for (int i = 0; i < n_cams; ++i) {
        dev[i] = reinterpret_cast<const int16_t *>(ctx->d_stitch_in + off[i]);
        if (n_shorts[i])
            if (cudaMemcpyAsync(ctx->d_stitch_in + off[i], payload_host[i], (size_t)n_shorts[i] * 2, cudaMemcpyHostToDevice, cs) != cudaSuccess)
                return fail(ctx, PCS_ERR_CUDA, "H2D failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    // records at +16 so that they are 16-byte aligned: the [int32] header sits at +12
    uint8_t *st = ctx->d_stitch_out + 12;
    int size = stitch_common(ctx, dev.data(), n_shorts, n_cams, downsample, transforms, pcl, st,
                             std::min(cap, ctx->cap_stitch_out - 12), nullptr, cs);
    if (size < 0) return size;
    if (cudaMemcpyAsync(stitched_host, st, (size_t)size + 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess ||
        cudaStreamSynchronize(cs) != cudaSuccess)
        return fail(ctx, PCS_ERR_CUDA, "D2H failed: %s", cudaGetErrorString(cudaGetLastError()));
    return size;
}

int pcs_b200_stitch_raw(pcs_ctx *ctx, const int16_t *const *payload_host, const int32_t *n_shorts,
                        int n_cams, int downsample, uint8_t *stitched_host, size_t stitched_cap) {
    return stitch_host(ctx, payload_host, n_shorts, n_cams, downsample, nullptr, false, stitched_host, stitched_cap);
}

int pcs_b200_stitch_pcl(pcs_ctx *ctx, const int16_t *const *payload_host, const int32_t *n_shorts,
                        int n_cams, int downsample, const float *transforms,
                        uint8_t *stitched_host, size_t stitched_cap) {
    if (!transforms) return fail(ctx, PCS_ERR_INVALID, "null transforms");
    return stitch_host(ctx, payload_host, n_shorts, n_cams, downsample, transforms, true, stitched_host, stitched_cap);
}

// ---- camera + stitch side in one call, host buffers ------------------------------------------
// (Re)builds a slot's device buffers and launch plan for a camera set.  Called under the slot lock.
static int stitch_slot_prepare(pcs_ctx *ctx, StitchSlot &sl, int n_cams, const int32_t *streams, int downsample) {
    bool same = (sl.batch || !sl.cam_batch.empty()) && (int)sl.streams.size() == n_cams && sl.downsample == downsample;
    for (int i = 0; same && i < n_cams; ++i)
        same = sl.streams[i] == streams[i] && sl.gens[i] == ctx->streams[streams[i]].geom_gen;
    if (same) return PCS_OK;
    if (!sl.cs) CU(ctx, cudaStreamCreateWithFlags(&sl.cs, cudaStreamNonBlocking));
    CU(ctx, cudaStreamSynchronize(sl.cs));
    for (auto *p : sl.d_z) cudaFree(p);
    for (auto *p : sl.d_c) cudaFree(p);
    cudaFree(sl.d_stitched); cudaFree(sl.d_decimated);
    if (sl.batch) pcs_b200_batch_destroy(ctx, sl.batch);
    for (auto c : sl.cam_cs) cudaStreamSynchronize(c);
    for (auto *b : sl.cam_batch) pcs_b200_batch_destroy(ctx, b);
    sl.cam_batch.clear(); sl.cam_off.clear();
    cudaFree(sl.d_counts); cudaFree(sl.d_total);
    sl.d_counts = sl.d_total = nullptr;
    sl.d_z.clear(); sl.d_c.clear(); sl.d_stitched = sl.d_decimated = nullptr; sl.batch = nullptr;
    sl.streams.clear(); sl.gens.clear();
    long long total = 0, out = 0;
    bool any_cutoff = false;
    for (int i = 0; i < n_cams; ++i) {
        const StreamParams &p = ctx->streams[streams[i]].params;
        total += p.N;
        out += (p.N + downsample - 1) / downsample;
        any_cutoff = any_cutoff || p.cutoff;
    }
    sl.counted = any_cutoff;
    if (total * 10 > 0x7fffffffll) return fail(ctx, PCS_ERR_CAPACITY, "stitched payload exceeds int32");
    std::vector<pcs_frame_job> jobs(n_cams);
    auto oom = [&]() { cudaGetLastError(); return fail(ctx, PCS_ERR_NOMEM, "stitch_frames: device allocation failed"); };
    if (cudaMalloc(&sl.d_stitched, (size_t)total * 10 + 32) != cudaSuccess) return oom();
    if ((downsample > 1 || any_cutoff) && cudaMalloc(&sl.d_decimated, (size_t)out * 10 + 32) != cudaSuccess) return oom();
    if (any_cutoff) {
        if (cudaMalloc(&sl.d_counts, (size_t)n_cams * 4 + 64) != cudaSuccess || cudaMalloc(&sl.d_total, 64) != cudaSuccess) return oom();
        if (!sl.h_total && cudaHostAlloc(&sl.h_total, 64, cudaHostAllocDefault) != cudaSuccess) return oom();
    }
    long long off = 0;
    for (int i = 0; i < n_cams; ++i) {
        const StreamParams &p = ctx->streams[streams[i]].params;
        uint16_t *z = nullptr;
        uint8_t *c = nullptr;
        if (cudaMalloc(&z, (size_t)p.N * 2 + 64) != cudaSuccess) return oom();
        sl.d_z.push_back(z);
        if (cudaMalloc(&c, (size_t)p.CH * p.stride + 64) != cudaSuccess) return oom();
        sl.d_c.push_back(c);
        jobs[i] = pcs_frame_job{};
        jobs[i].stream = streams[i];
        jobs[i].z16_dev = z;
        jobs[i].color_dev = c;
        jobs[i].payload_dev = reinterpret_cast<int16_t *>(sl.d_stitched + 16 + off * 10);
        if (any_cutoff) {          // (a -c camera compacts in place of its dense records: its slot is its payload)
            jobs[i].count_dev = p.cutoff ? sl.d_counts + i : nullptr;
        }
        off += p.N;
    }
    // Two pipelines.  per camera (default when nothing is decimated): every camera has its own stream -- frame up,
    // its kernel, its records down -- so camera 0's records are on their way back while camera 7's frame is still going
    // up (this is what keeps both directions of the link busy: +9 % over one batched launch + one big copy).
    // batch: all cameras in ONE launch, then (decimation and) one copy.
    static const int pipeline = pipe_knob("PCS_STITCH_PIPELINE", 0, 0, 1);     // 0 = per camera, 1 = batch
    sl.per_camera = downsample == 1 && pipeline == 0 && !any_cutoff;
    int rc;
    if (sl.per_camera) {
        long long o = 0;
        for (int i = 0; i < n_cams; ++i) {
            pcs_batch *b = nullptr;
            if ((rc = pcs_b200_batch_create(ctx, &jobs[i], 1, &b))) return rc;
            sl.cam_batch.push_back(b);
            sl.cam_off.push_back(o);
            o += ctx->streams[streams[i]].params.N;
        }
        while ((int)sl.cam_cs.size() < n_cams) {
            cudaStream_t c = nullptr;
            CU(ctx, cudaStreamCreateWithFlags(&c, cudaStreamNonBlocking));
            sl.cam_cs.push_back(c);
        }
    } else if ((rc = pcs_b200_batch_create(ctx, jobs.data(), n_cams, &sl.batch))) {
        return rc;
    }
    // the int32 size header never changes for a camera set (no cutoff here): written once
    const int32_t full = (int32_t)(total * 10), dec = (int32_t)(out * 10);
    CU(ctx, cudaMemcpy(sl.d_stitched + 12, &full, 4, cudaMemcpyHostToDevice));
    if (sl.d_decimated) CU(ctx, cudaMemcpy(sl.d_decimated + 12, &dec, 4, cudaMemcpyHostToDevice));
    sl.total_pts = total;
    sl.out_bytes = out * 10;
    sl.downsample = downsample;
    sl.streams.assign(streams, streams + n_cams);
    for (int i = 0; i < n_cams; ++i) sl.gens.push_back(ctx->streams[streams[i]].geom_gen);
    return PCS_OK;
}

int pcs_b200_stitch_frames_begin(pcs_ctx *ctx, int slot, int n_cams, const int32_t *streams,
                                 const uint16_t *const *z16_host, const uint8_t *const *color_host,
                                 int downsample, uint8_t *stitched_host, size_t stitched_cap) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    if (slot < 0 || slot >= PCS_B200_STITCH_SLOTS) return fail(ctx, PCS_ERR_INVALID, "slot out of range");
    if (n_cams < 1 || n_cams > MAX_CAMS || !streams || !z16_host || !color_host || !stitched_host || downsample < 1)
        return fail(ctx, PCS_ERR_INVALID, "bad stitch_frames arguments (1 <= n_cams <= %d)", MAX_CAMS);
    for (int i = 0; i < n_cams; ++i) {
        int rc = check_stream(ctx, streams[i], true);
        if (rc) return rc;
        if (!z16_host[i] || !color_host[i]) return fail(ctx, PCS_ERR_INVALID, "camera %d: null frame", i);
    }
    StitchSlot &sl = ctx->stitch_slots[slot];
    std::lock_guard<std::mutex> lk(sl.mu);
    if (sl.pending) return fail(ctx, PCS_ERR_INVALID, "slot %d already has a frame in flight", slot);
    DeviceGuard dg_(ctx->device);
    int rc = stitch_slot_prepare(ctx, sl, n_cams, streams, downsample);
    if (rc) return rc;
    if ((size_t)sl.out_bytes + 4 > stitched_cap)
        return fail(ctx, PCS_ERR_CAPACITY, "stitched buffer too small: need %lld bytes", sl.out_bytes + 4);
    if (sl.per_camera) {
        for (int i = 0; i < n_cams; ++i) {
            const StreamParams &p = ctx->streams[streams[i]].params;
            cudaStream_t c = sl.cam_cs[i];
            CU(ctx, cudaMemcpyAsync(sl.d_z[i], z16_host[i], (size_t)p.N * 2, cudaMemcpyHostToDevice, c));
            CU(ctx, cudaMemcpyAsync(sl.d_c[i], color_host[i], (size_t)p.CH * p.stride, cudaMemcpyHostToDevice, c));
            if ((rc = pcs_b200_batch_run(ctx, sl.cam_batch[i], c))) return rc;
            CU(ctx, cudaMemcpyAsync(stitched_host + 4 + sl.cam_off[i] * 10, sl.d_stitched + 16 + sl.cam_off[i] * 10,
                                    (size_t)p.N * 10, cudaMemcpyDeviceToHost, c));
        }
        const int32_t total = (int32_t)sl.out_bytes;
        memcpy(stitched_host, &total, 4);           // :394-395 (nothing on the device writes these four bytes)
        sl.pending = true;
        return PCS_OK;
    }
    for (int i = 0; i < n_cams; ++i) {
        const StreamParams &p = ctx->streams[streams[i]].params;
        CU(ctx, cudaMemcpyAsync(sl.d_z[i], z16_host[i], (size_t)p.N * 2, cudaMemcpyHostToDevice, sl.cs));
        CU(ctx, cudaMemcpyAsync(sl.d_c[i], color_host[i], (size_t)p.CH * p.stride, cudaMemcpyHostToDevice, sl.cs));
    }
    if ((rc = pcs_b200_batch_run(ctx, sl.batch, sl.cs))) return rc;
    const uint8_t *src = sl.d_stitched + 12;
    if (sl.counted) {
        // -c: the cameras' record counts only exist on the device.  The concat reads them there; its total comes back
        // first (4 bytes), the records once end() knows how many there are.
        CountedTable t{};
        t.n_cams = n_cams;
        t.downsample = downsample;
        long long off = 0;
        for (int i = 0; i < n_cams; ++i) {
            const StreamParams &p = ctx->streams[streams[i]].params;
            t.src[i] = reinterpret_cast<const int16_t *>(sl.d_stitched + 16 + off * 10);
            t.cnt_dev[i] = p.cutoff ? sl.d_counts + i : nullptr;
            t.cnt_fixed[i] = p.N;
            off += p.N;
        }
        const int blocks = std::max(1, std::min((int)((sl.out_bytes / 10 + 255) / 256), ctx->sm_count * 8));
        stitch_counted<<<blocks, 256, 0, sl.cs>>>(t, sl.d_decimated + 12, sl.d_total);
        CU(ctx, cudaGetLastError());
        CU(ctx, cudaMemcpyAsync(sl.h_total, sl.d_total, 4, cudaMemcpyDeviceToHost, sl.cs));
        sl.pending_host = stitched_host;
        sl.pending = true;
        return PCS_OK;
    }
    if (downsample > 1) {
        // every downsample-th record of each camera (src/pcs-multicamera-client.cpp:388), device to device
        std::vector<const int16_t *> pay(n_cams);
        std::vector<int32_t> ns(n_cams);
        long long off = 0;
        for (int i = 0; i < n_cams; ++i) {
            const StreamParams &p = ctx->streams[streams[i]].params;
            pay[i] = reinterpret_cast<const int16_t *>(sl.d_stitched + 16 + off * 10);
            ns[i] = p.N * 5;
            off += p.N;
        }
        rc = stitch_common(ctx, pay.data(), ns.data(), n_cams, downsample, nullptr, false, sl.d_decimated + 12,
                           (size_t)sl.out_bytes + 4, nullptr, sl.cs);
        if (rc < 0) return rc;
        src = sl.d_decimated + 12;
    }
    CU(ctx, cudaMemcpyAsync(stitched_host, src, (size_t)sl.out_bytes + 4, cudaMemcpyDeviceToHost, sl.cs));
    sl.pending = true;
    return PCS_OK;
}

int pcs_b200_stitch_frames_end(pcs_ctx *ctx, int slot) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    if (slot < 0 || slot >= PCS_B200_STITCH_SLOTS) return fail(ctx, PCS_ERR_INVALID, "slot out of range");
    StitchSlot &sl = ctx->stitch_slots[slot];
    std::lock_guard<std::mutex> lk(sl.mu);
    if (!sl.pending) return fail(ctx, PCS_ERR_INVALID, "slot %d has no frame in flight", slot);
    sl.pending = false;
    DeviceGuard dg_(ctx->device);
    if (sl.per_camera) {
        for (size_t i = 0; i < sl.streams.size(); ++i) CU(ctx, cudaStreamSynchronize(sl.cam_cs[i]));
    } else {
        CU(ctx, cudaStreamSynchronize(sl.cs));
    }
    if (sl.counted) {
        const int32_t bytes = *sl.h_total;
        if (bytes < 0 || bytes > sl.out_bytes) return fail(ctx, PCS_ERR_CUDA, "stitch_frames: inconsistent record count");
        CU(ctx, cudaMemcpyAsync(sl.pending_host, sl.d_decimated + 12, (size_t)bytes + 4, cudaMemcpyDeviceToHost, sl.cs));
        CU(ctx, cudaStreamSynchronize(sl.cs));
        return bytes;
    }
    return (int)sl.out_bytes;
}

int pcs_b200_stitch_frames(pcs_ctx *ctx, int n_cams, const int32_t *streams, const uint16_t *const *z16_host,
                           const uint8_t *const *color_host, int downsample, uint8_t *stitched_host,
                           size_t stitched_cap) {
    int rc = pcs_b200_stitch_frames_begin(ctx, 0, n_cams, streams, z16_host, color_host, downsample, stitched_host,
                                          stitched_cap);
    if (rc) return rc;
    return pcs_b200_stitch_frames_end(ctx, 0);
}

// ---- voxel merge -----------------------------------------------------------------
static int voxel_args_ok(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, const void *out) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    if (n < 0 || leaf_mm < 1 || leaf_mm > 32767 || (n && (!records_dev || !out)))
        return fail(ctx, PCS_ERR_INVALID, "bad voxel-merge arguments");
    // n * 10 bytes is what the reference's int32 size header can describe (src/pcs-multicamera-client.cpp:394);
    // the sort-based variants have tighter limits of their own (voxel_run)
    if ((long long)n * 10 > 0x7FFFFFFFll)
        return fail(ctx, PCS_ERR_UNSUPPORTED, "voxel merge: n * 10 bytes must fit the int32 size header");
    return PCS_OK;
}

static int voxel_fail(pcs_ctx *ctx, int rc) {
    return fail(ctx, rc == -3 ? PCS_ERR_NOMEM : (rc == -4 ? PCS_ERR_UNSUPPORTED : PCS_ERR_CUDA),
                "voxel merge failed (%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
}

// voxel_variant: 0 = auto: the one-sweep sort when the (key, index) word fits 64 bits and the uint32 sums can
// hold n points, else the slab partition + bitmap ranking (no limit on n), else the pair sort; 1 = pair sort,
// 2 / 3 = one-sweep sort with 8- / 10-bit digits, 4 = slab partition + bitmap ranking only
static int voxel_run_msd(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, int16_t *out_dev, cudaStream_t cs,
                         bool slab, int kz_lo, int kz_hi) {
    int32_t *nv_dev = nullptr;
    int rc = voxel_merge_msd(ctx->voxel, records_dev, n, leaf_mm, out_dev, cs, ctx->sm_count, slab, kz_lo, kz_hi, &nv_dev);
    if (rc != 0) return rc;
    if (cudaMemcpyAsync(ctx->voxel.h_count, nv_dev, 4, cudaMemcpyDeviceToHost, cs) != cudaSuccess ||
        cudaStreamSynchronize(cs) != cudaSuccess || cudaGetLastError() != cudaSuccess)
        return -2;
    const int nv = ctx->voxel.h_count[0];
    return (nv < 0 || nv > n) ? -2 : nv;
}

static int voxel_run(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, int16_t *out_dev, cudaStream_t cs,
                     bool slab, int kz_lo, int kz_hi) {
    const int vv = ctx->voxel_variant;
    if (vv == 4) return voxel_run_msd(ctx, records_dev, n, leaf_mm, out_dev, cs, slab, kz_lo, kz_hi);
    // the sort-based variants keep uint32 per-voxel sums (oracle/SPEC.md s3): 255 * n and (leaf - 1) * n must fit,
    // with room for one more count (the mean's +-1 fix-up works modulo 2^32)
    const bool sums_fit = (long long)n * 256 <= 0xFFFFFFFFll && (long long)n * leaf_mm <= 0xFFFFFFFFll;
    int rc = -4;
    if (sums_fit) {
        if (vv == 3)
            rc = voxel_merge_sweep<10, 256, 16>(ctx->voxel, records_dev, n, leaf_mm, out_dev, cs, ctx->sm_count, ctx->plan_slot, slab, kz_lo, kz_hi);
        else if (vv == 0 || vv == 2 || slab)
            rc = voxel_merge_sweep<8, 256, 16>(ctx->voxel, records_dev, n, leaf_mm, out_dev, cs, ctx->sm_count, ctx->plan_slot, slab, kz_lo, kz_hi);
    }
    if (vv == 0 && rc == -4) rc = voxel_run_msd(ctx, records_dev, n, leaf_mm, out_dev, cs, slab, kz_lo, kz_hi);
    if (sums_fit && !slab && (vv == 1 || (vv == 0 && rc == -4)))
        rc = voxel_merge(ctx->voxel, records_dev, n, leaf_mm, out_dev, cs);
    return rc;
}

int pcs_b200_voxel_merge_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm,
                             int16_t *out_dev, void *cuda_stream) {
    int rc = voxel_args_ok(ctx, records_dev, n, leaf_mm, out_dev);
    if (rc) return rc;
    if (n == 0) return 0;
    DeviceGuard dg_(ctx->device);
    std::lock_guard<std::mutex> lk(ctx->scratch_mu);
    rc = voxel_run(ctx, records_dev, n, leaf_mm, out_dev, (cudaStream_t)cuda_stream, false, 0, 0);
    return rc < 0 ? voxel_fail(ctx, rc) : rc;
}

// Enqueue-only forms: nothing returns to the host; *count_dev = voxel count, or a negative pcs_status.
static int voxel_async(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, int16_t *out_dev, int32_t *count_dev,
                       cudaStream_t cs, bool slab, int kz_lo, int kz_hi, const int32_t *n_dev = nullptr) {
    int rc = voxel_args_ok(ctx, records_dev, n, leaf_mm, out_dev);
    if (rc) return rc;
    if (!count_dev) return fail(ctx, PCS_ERR_INVALID, "count_dev must not be NULL");
    if (ctx->voxel_variant != 0 && ctx->voxel_variant != 2 && ctx->voxel_variant != 3)
        return fail(ctx, PCS_ERR_UNSUPPORTED, "the asynchronous merge is the one-sweep sort (voxel_variant 0, 2 or 3)");
    if ((long long)n * 256 > 0xFFFFFFFFll || (long long)n * leaf_mm > 0xFFFFFFFFll)
        return fail(ctx, PCS_ERR_UNSUPPORTED, "asynchronous voxel merge: n * max(256, leaf_mm) must stay below 2^32");
    DeviceGuard dg_(ctx->device);
    if (n == 0) {
        CU(ctx, cudaMemsetAsync(count_dev, 0, 4, cs));
        return PCS_OK;
    }
    std::lock_guard<std::mutex> lk(ctx->scratch_mu);
    int32_t *nv_dev = nullptr;
    if (ctx->voxel_variant == 3)
        rc = voxel_merge_sweep_enqueue<10, 256, 16>(ctx->voxel, records_dev, n, leaf_mm, out_dev, cs, ctx->sm_count, ctx->plan_slot,
                                                    slab, kz_lo, kz_hi, &nv_dev, n_dev);
    else
        rc = voxel_merge_sweep_enqueue<8, 256, 16>(ctx->voxel, records_dev, n, leaf_mm, out_dev, cs, ctx->sm_count, ctx->plan_slot,
                                                   slab, kz_lo, kz_hi, &nv_dev, n_dev);
    if (rc < 0) return voxel_fail(ctx, rc);
    CU(ctx, cudaMemcpyAsync(count_dev, nv_dev, 4, cudaMemcpyDeviceToDevice, cs));
    return PCS_OK;
}

int pcs_b200_voxel_merge_async_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, int16_t *out_dev,
                                   int32_t *count_dev, void *cuda_stream) {
    return voxel_async(ctx, records_dev, n, leaf_mm, out_dev, count_dev, (cudaStream_t)cuda_stream, false, 0, 0);
}

int pcs_b200_voxel_merge_slab_async_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, int kz_lo, int kz_hi,
                                        int16_t *out_dev, int32_t *count_dev, void *cuda_stream) {
    return voxel_async(ctx, records_dev, n, leaf_mm, out_dev, count_dev, (cudaStream_t)cuda_stream, true, kz_lo, kz_hi);
}

int pcs_b200_voxel_merge_counted_async_dev(pcs_ctx *ctx, const int16_t *records_dev, int n_max, const int32_t *n_dev,
                                           int leaf_mm, int16_t *out_dev, int32_t *count_dev, void *cuda_stream) {
    if (!n_dev) return fail(ctx, PCS_ERR_INVALID, "n_dev must not be NULL");
    return voxel_async(ctx, records_dev, n_max, leaf_mm, out_dev, count_dev, (cudaStream_t)cuda_stream, false, 0, 0, n_dev);
}

// ---- multi-GPU: shard by voxel-key range before the exchange (pcs_exchange.cuh) ---------------------------
int pcs_b200_shard_zbins(int leaf_mm) {
    if (leaf_mm < 1 || leaf_mm > 32767) return PCS_ERR_INVALID;
    return sweep_zbins(sweep_base_geom(leaf_mm));
}

static int shard_peers(pcs_ctx *ctx, const pcs_shard_peers *p, XaPeers &x) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    if (!p || p->n_ranks < 1 || p->n_ranks > XA_MAX_RANKS || p->rank < 0 || p->rank >= p->n_ranks || p->capacity_records < 1)
        return fail(ctx, PCS_ERR_INVALID, "bad peer table (1 <= n_ranks <= %d)", XA_MAX_RANKS);
    memset(&x, 0, sizeof x);
    x.n_ranks = p->n_ranks;
    x.rank = p->rank;
    x.capacity = p->capacity_records;
    for (int r = 0; r < p->n_ranks; ++r) {
        if (!p->inbox_dev[r] || !p->cursor_dev[r] || !p->zhist_dev[r]) return fail(ctx, PCS_ERR_INVALID, "rank %d: null pointer", r);
        x.inbox[r] = static_cast<uint16_t *>(p->inbox_dev[r]);
        x.cursor[r] = p->cursor_dev[r];
        x.zhist[r] = p->zhist_dev[r];
    }
    return PCS_OK;
}

int pcs_b200_shard_hist_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, uint32_t *zhist_dev,
                            uint32_t *cursor_dev, void *cuda_stream) {
    int rc = voxel_args_ok(ctx, records_dev, n, leaf_mm, zhist_dev);
    if (rc) return rc;
    if (!zhist_dev || !cursor_dev) return fail(ctx, PCS_ERR_INVALID, "null buffer");
    DeviceGuard dg_(ctx->device);
    cudaStream_t cs = (cudaStream_t)cuda_stream;
    const SweepGeom g = sweep_base_geom(leaf_mm);
    const int zbins = sweep_zbins(g);
    if ((size_t)zbins * 4 > 40 * 1024) return fail(ctx, PCS_ERR_UNSUPPORTED, "leaf too small for the sharded merge (z histogram > 40 KB)");
    CU(ctx, cudaMemsetAsync(zhist_dev, 0, (size_t)zbins * 4, cs));
    CU(ctx, cudaMemsetAsync(cursor_dev, 0, 4, cs));       // my inbox is empty again (peers add after the barrier)
    if (n > 0) {
        const int tiles = (n + SW_KH_TILE - 1) / SW_KH_TILE;
        const int grid = std::max(1, std::min(tiles, std::max(1, ctx->sm_count) * 4));
        xa_zhist<<<grid, SW_KH_THREADS, (size_t)zbins * 4, cs>>>(records_dev, n, g, zhist_dev, zbins);
        CU(ctx, cudaGetLastError());
    }
    return PCS_OK;
}

int pcs_b200_shard_plan_dev(pcs_ctx *ctx, const pcs_shard_peers *peers, int leaf_mm, int32_t *kz_splits_dev,
                            uint8_t *zslab_dev, void *cuda_stream) {
    XaPeers x;
    int rc = shard_peers(ctx, peers, x);
    if (rc) return rc;
    if (leaf_mm < 1 || leaf_mm > 32767 || !kz_splits_dev || !zslab_dev) return fail(ctx, PCS_ERR_INVALID, "bad arguments");
    DeviceGuard dg_(ctx->device);
    const SweepGeom g = sweep_base_geom(leaf_mm);
    const int zbins = sweep_zbins(g);
    if ((size_t)zbins * 4 > 40 * 1024) return fail(ctx, PCS_ERR_UNSUPPORTED, "leaf too small for the sharded merge");
    xa_plan<<<1, 1024, (size_t)zbins * 4, (cudaStream_t)cuda_stream>>>(x, zbins, g.K, kz_splits_dev, zslab_dev);
    CU(ctx, cudaGetLastError());
    return PCS_OK;
}

int pcs_b200_shard_scatter_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, const uint8_t *zslab_dev,
                               const pcs_shard_peers *peers, uint32_t *err_dev, void *cuda_stream) {
    XaPeers x;
    int rc = shard_peers(ctx, peers, x);
    if (rc) return rc;
    if ((rc = voxel_args_ok(ctx, records_dev, n, leaf_mm, zslab_dev))) return rc;
    if (!zslab_dev || !err_dev) return fail(ctx, PCS_ERR_INVALID, "null buffer");
    if (n == 0) return PCS_OK;
    DeviceGuard dg_(ctx->device);
    const SweepGeom g = sweep_base_geom(leaf_mm);
    const int zbins = sweep_zbins(g);
    const size_t smem = xa_scatter_smem(zbins);
    static bool configured = false;
    if (!configured) {
        CU(ctx, cudaFuncSetAttribute(xa_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured = true;
    }
    if (smem > 100 * 1024) return fail(ctx, PCS_ERR_UNSUPPORTED, "leaf too small for the sharded merge");
    const int tiles = (n + XA_TILE - 1) / XA_TILE;
    const int grid = std::max(1, std::min(tiles, std::max(1, ctx->sm_count) * 4));
    xa_scatter<<<grid, XA_THREADS, smem, (cudaStream_t)cuda_stream>>>(records_dev, n, g, x, zslab_dev, zbins, err_dev);
    CU(ctx, cudaGetLastError());
    return PCS_OK;
}

int pcs_b200_voxel_slab_plan_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, int n_slabs,
                                 int32_t *kz_splits, int32_t *slab_points, void *cuda_stream) {
    int rc = voxel_args_ok(ctx, records_dev, n, leaf_mm, kz_splits);
    if (rc) return rc;
    if (n_slabs < 1 || n_slabs > 1024 || !kz_splits) return fail(ctx, PCS_ERR_INVALID, "1 <= n_slabs <= 1024");
    DeviceGuard dg_(ctx->device);
    std::lock_guard<std::mutex> lk(ctx->scratch_mu);
    rc = voxel_slab_plan(ctx->voxel, records_dev, n, leaf_mm, n_slabs, kz_splits, slab_points,
                         (cudaStream_t)cuda_stream, ctx->sm_count);
    return rc < 0 ? voxel_fail(ctx, rc) : PCS_OK;
}

int pcs_b200_voxel_merge_slab_dev(pcs_ctx *ctx, const int16_t *records_dev, int n, int leaf_mm, int kz_lo,
                                  int kz_hi, int16_t *out_dev, void *cuda_stream) {
    int rc = voxel_args_ok(ctx, records_dev, n, leaf_mm, out_dev);
    if (rc) return rc;
    if (n == 0 || kz_lo >= kz_hi) return 0;
    DeviceGuard dg_(ctx->device);
    std::lock_guard<std::mutex> lk(ctx->scratch_mu);
    rc = voxel_run(ctx, records_dev, n, leaf_mm, out_dev, (cudaStream_t)cuda_stream, true, kz_lo, kz_hi);
    return rc < 0 ? voxel_fail(ctx, rc) : rc;
}

int pcs_b200_voxel_merge(pcs_ctx *ctx, const int16_t *records_host, int n, int leaf_mm,
                         int16_t *out_host) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    if (n < 0 || (n && (!records_host || !out_host))) return fail(ctx, PCS_ERR_INVALID, "bad arguments");
    if (n == 0) return 0;
    DeviceGuard dg_(ctx->device);
    // device staging kept across calls (the stitch scratch: records in, voxels out)
    const size_t bytes = (size_t)n * 10 + 64;
    {
        std::lock_guard<std::mutex> lk(ctx->scratch_mu);
        if (bytes > ctx->cap_stitch_in) {
            cudaFree(ctx->d_stitch_in);
            ctx->d_stitch_in = nullptr; ctx->cap_stitch_in = 0;
            if (cudaMalloc(&ctx->d_stitch_in, bytes) != cudaSuccess) { cudaGetLastError(); return fail(ctx, PCS_ERR_NOMEM, "cudaMalloc failed"); }
            ctx->cap_stitch_in = bytes;
        }
        if (bytes > ctx->cap_stitch_out) {
            cudaFree(ctx->d_stitch_out);
            ctx->d_stitch_out = nullptr; ctx->cap_stitch_out = 0;
            if (cudaMalloc(&ctx->d_stitch_out, bytes) != cudaSuccess) { cudaGetLastError(); return fail(ctx, PCS_ERR_NOMEM, "cudaMalloc failed"); }
            ctx->cap_stitch_out = bytes;
        }
    }
    int16_t *d_in = reinterpret_cast<int16_t *>(ctx->d_stitch_in), *d_out = reinterpret_cast<int16_t *>(ctx->d_stitch_out);
    cudaStream_t cs = ctx->streams[0].cs;
    if (cudaMemcpyAsync(d_in, records_host, (size_t)n * 10, cudaMemcpyHostToDevice, cs) != cudaSuccess)
        return fail(ctx, PCS_ERR_CUDA, "H2D failed: %s", cudaGetErrorString(cudaGetLastError()));
    int rc = pcs_b200_voxel_merge_dev(ctx, d_in, n, leaf_mm, d_out, cs);
    if (rc > 0 && (cudaMemcpyAsync(out_host, d_out, (size_t)rc * 10, cudaMemcpyDeviceToHost, cs) != cudaSuccess ||
                   cudaStreamSynchronize(cs) != cudaSuccess))
        rc = fail(ctx, PCS_ERR_CUDA, "copy back failed: %s", cudaGetErrorString(cudaGetLastError()));
    return rc;
}

// ---- PLY dump ----------------------------------------------------------------------------
int pcs_b200_cloud_to_ply_rows_dev(pcs_ctx *ctx, const void *cloud32_dev, int n, uint8_t *rows_dev, void *cuda_stream) {
    if (!ctx) return fail(nullptr, PCS_ERR_INVALID, "null context");
    if (n < 0 || (n && (!cloud32_dev || !rows_dev))) return fail(ctx, PCS_ERR_INVALID, "bad PLY arguments");
    if (reinterpret_cast<uintptr_t>(cloud32_dev) & 15) return fail(ctx, PCS_ERR_INVALID, "cloud must be 16-byte aligned");
    if (n == 0) return 0;
    DeviceGuard dg_(ctx->device);
    const int groups = (n + 31) / 32;
    const int grid = std::max(1, std::min((groups + PLY_THREADS / 32 - 1) / (PLY_THREADS / 32), ctx->sm_count * 8));
    ply_rows<<<grid, PLY_THREADS, 0, (cudaStream_t)cuda_stream>>>(reinterpret_cast<const uint4 *>(cloud32_dev), n, rows_dev);
    CU(ctx, cudaGetLastError());
    return n;
}

// Header as PCL's PLYWriter emits it for a PointXYZRGB cloud (recalled from PCL 1.8; PCL is not in
// /root/reference, so the exact text is unpinned -- any PLY reader accepts it): vertex element, then a
// one-row camera element (identity pose, focal 0, viewport = cloud width x height = n x 1).
int pcs_b200_save_ply(pcs_ctx *ctx, const void *cloud32_dev, int n, const char *path) {
    if (!ctx || !path) return fail(ctx, PCS_ERR_INVALID, "null argument");
    if (n < 0 || (n && !cloud32_dev)) return fail(ctx, PCS_ERR_INVALID, "bad PLY arguments");
    DeviceGuard dg_(ctx->device);
    uint8_t *d_rows = nullptr;
    std::vector<uint8_t> rows((size_t)n * PLY_ROW);
    if (n) {
        if (cudaMalloc(&d_rows, (size_t)n * PLY_ROW + 16) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, PCS_ERR_NOMEM, "cudaMalloc failed");
        }
        cudaStream_t cs = ctx->streams[0].cs;
        int rc = pcs_b200_cloud_to_ply_rows_dev(ctx, cloud32_dev, n, d_rows, cs);
        if (rc >= 0 && (cudaMemcpyAsync(rows.data(), d_rows, rows.size(), cudaMemcpyDeviceToHost, cs) != cudaSuccess ||
                        cudaStreamSynchronize(cs) != cudaSuccess))
            rc = fail(ctx, PCS_ERR_CUDA, "PLY copy back failed: %s", cudaGetErrorString(cudaGetLastError()));
        cudaFree(d_rows);
        if (rc < 0) return rc;
    }
    FILE *f = fopen(path, "wb");
    if (!f) return fail(ctx, PCS_ERR_INVALID, "cannot open %s", path);
    fprintf(f, "ply\nformat binary_little_endian 1.0\ncomment PCL generated\nelement vertex %d\n"
               "property float x\nproperty float y\nproperty float z\n"
               "property uchar red\nproperty uchar green\nproperty uchar blue\n"
               "element camera 1\n"
               "property float view_px\nproperty float view_py\nproperty float view_pz\n"
               "property float x_axisx\nproperty float x_axisy\nproperty float x_axisz\n"
               "property float y_axisx\nproperty float y_axisy\nproperty float y_axisz\n"
               "property float z_axisx\nproperty float z_axisy\nproperty float z_axisz\n"
               "property float focal\nproperty float scalex\nproperty float scaley\n"
               "property float centerx\nproperty float centery\n"
               "property int viewportx\nproperty int viewporty\n"
               "property float k1\nproperty float k2\nend_header\n", n);
    bool ok = rows.empty() || fwrite(rows.data(), 1, rows.size(), f) == rows.size();
    const float cam_a[17] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0};
    const int32_t viewport[2] = {n, 1};
    const float k12[2] = {0, 0};
    ok = ok && fwrite(cam_a, 4, 17, f) == 17 && fwrite(viewport, 4, 2, f) == 2 && fwrite(k12, 4, 2, f) == 2;
    ok = (fclose(f) == 0) && ok;
    if (!ok) return fail(ctx, PCS_ERR_INVALID, "short write to %s", path);
    return n;
}

}  // extern "C"
