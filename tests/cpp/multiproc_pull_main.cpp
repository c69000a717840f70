// Two PROCESSES, one GPU each -- the reference's deployment shape (one process per camera, fan-in at the
// stitcher, src/pcs-multicamera-client.cpp:363-371) -- doing the pull exchange through the C ABI only:
// each process cudaMalloc's its cameras' frames, exports them with pcs_b200_ipc_export, sends the 80-byte
// handles to the other process over a socketpair, maps the peer's frames with pcs_b200_ipc_open and runs
// ONE batch over ALL cameras: its own frames come from HBM, the peer's over NVLink, read by the fused
// kernel itself.  Both processes end with the whole stitched buffer
// [int32 bytes][cam0 records][cam1 records]... (:385-395), compared here with the oracle restatement.
// No torch, no NCCL, no MPI.  Built and run by tests/test_multigpu_gpu.py.  Exit code 0 = bit-exact in
// both processes (or fewer than two GPUs: prints SKIP).
#include <cuda_runtime.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "pcs_b200.h"
#include "pcs_oracle.h"

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            std::printf("[rank %d] CUDA error %s at %s:%d\n", rank, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)
#define PCS(ctx, call)                                                                        \
    do {                                                                                      \
        int rc_ = (call);                                                                     \
        if (rc_ < 0) {                                                                        \
            std::printf("[rank %d] pcs error %d (%s) at %s:%d\n", rank, rc_, pcs_b200_last_error(ctx), __FILE__, __LINE__); \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

static uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

static bool xfer(int fd, void *buf, size_t n, bool out) {
    uint8_t *p = static_cast<uint8_t *>(buf);
    while (n) {
        const ssize_t k = out ? write(fd, p, n) : read(fd, p, n);
        if (k <= 0) return false;
        p += k;
        n -= (size_t)k;
    }
    return true;
}
static bool barrier(int fd) {      // both sides write a byte, then read one
    char c = 'b';
    return xfer(fd, &c, 1, true) && xfer(fd, &c, 1, false);
}

static const int W = 256, H = 48, N = W * H, CAMS = 5;     // ragged: rank 0 owns cameras 0-2, rank 1 owns 3-4
static const int FIRST[3] = {0, 3, CAMS};
static const float TF[16] = {-0.99977970f, 0.00926272f, 0.01883480f, 0.f, -0.01638983f, 0.21604544f,
                             -0.97624574f, 3.416f, -0.01311186f, -0.97633937f, -0.21584603f, 1.802f,
                             0.f, 0.f, 0.f, 1.f};

static int run(int rank, int fd) {
    // every process can regenerate every frame (the check needs all of them); only its own go to its GPU
    pcs_oracle_calib cal;
    std::memset(&cal, 0, sizeof cal);
    cal.depth.width = cal.color.width = W; cal.depth.height = cal.color.height = H;
    cal.depth.fx = cal.depth.fy = cal.color.fx = cal.color.fy = W / 2.f;
    cal.depth.ppx = cal.color.ppx = (W - 1) / 2.f; cal.depth.ppy = cal.color.ppy = (H - 1) / 2.f;
    cal.rotation[0] = cal.rotation[4] = cal.rotation[8] = 1.f;
    cal.translation[0] = 0.015f;
    cal.depth_scale = 0.001f;
    std::vector<std::vector<uint16_t>> z(CAMS, std::vector<uint16_t>(N));
    std::vector<std::vector<uint8_t>> col(CAMS, std::vector<uint8_t>((size_t)N * 3));
    std::vector<int16_t> want((size_t)CAMS * N * 5);
    uint32_t seed = 4242;
    for (int c = 0; c < CAMS; ++c) {
        float tfc[16];
        std::memcpy(tfc, TF, sizeof tfc);
        tfc[3] += 0.25f * c;
        for (int i = 0; i < N; ++i) z[c][i] = (lcg(seed) % 16 == 0) ? 0 : (uint16_t)(300 + lcg(seed) % 5700);
        for (size_t i = 0; i < col[c].size(); ++i) col[c][i] = (uint8_t)lcg(seed);
        std::vector<float> xyz((size_t)N * 3), uv((size_t)N * 2);
        pcs_oracle_deproject(&cal, z[c].data(), xyz.data(), uv.data(), 1);
        if (pcs_oracle_pack_simd(xyz.data(), uv.data(), N, col[c].data(), W, H, 3, W * 3, tfc, 0,
                                 want.data() + (size_t)c * N * 5) != N) return 2;
    }
    CK(cudaSetDevice(rank));
    pcs_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.device = rank;
    cfg.max_streams = CAMS;
    pcs_ctx *ctx = nullptr;
    PCS(nullptr, pcs_b200_create(&cfg, &ctx));
    // ONE allocation holds all of this rank's frames (depth then colour per camera, 256-byte aligned):
    // the handle of an interior pointer carries its offset
    const size_t zb = ((size_t)N * 2 + 255) & ~(size_t)255, cb = ((size_t)N * 3 + 255) & ~(size_t)255;
    const int mine = FIRST[rank + 1] - FIRST[rank];
    uint8_t *frames = nullptr;
    CK(cudaMalloc(&frames, (zb + cb) * mine));
    const void *d_z[CAMS], *d_col[CAMS];
    pcs_ipc_handle hz[CAMS], hc[CAMS];
    std::memset(hz, 0, sizeof hz);
    std::memset(hc, 0, sizeof hc);
    for (int c = FIRST[rank]; c < FIRST[rank + 1]; ++c) {
        uint8_t *pz = frames + (zb + cb) * (c - FIRST[rank]), *pc = pz + zb;
        CK(cudaMemcpy(pz, z[c].data(), (size_t)N * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(pc, col[c].data(), (size_t)N * 3, cudaMemcpyHostToDevice));
        d_z[c] = pz;
        d_col[c] = pc;
        PCS(ctx, pcs_b200_ipc_export(ctx, pz, &hz[c]));
        PCS(ctx, pcs_b200_ipc_export(ctx, pc, &hc[c]));
    }
    // exchange the handles of every camera (each side sends its own, receives the peer's)
    const int peer = 1 - rank;
    for (int c = FIRST[rank]; c < FIRST[rank + 1]; ++c)
        if (!xfer(fd, &hz[c], sizeof hz[c], true) || !xfer(fd, &hc[c], sizeof hc[c], true)) return 2;
    for (int c = FIRST[peer]; c < FIRST[peer + 1]; ++c) {
        if (!xfer(fd, &hz[c], sizeof hz[c], false) || !xfer(fd, &hc[c], sizeof hc[c], false)) return 2;
        void *pz = nullptr, *pc = nullptr;
        PCS(ctx, pcs_b200_ipc_open(ctx, &hz[c], &pz));
        PCS(ctx, pcs_b200_ipc_open(ctx, &hc[c], &pc));
        d_z[c] = pz;
        d_col[c] = pc;
    }
    const size_t payload = (size_t)CAMS * N * 10;
    uint8_t *d_stitched = nullptr;
    CK(cudaMalloc(&d_stitched, 16 + payload));
    std::vector<pcs_frame_job> jobs(CAMS);
    for (int c = 0; c < CAMS; ++c) {
        pcs_stream_desc d;
        std::memset(&d, 0, sizeof d);
        d.depth.width = d.color.width = W; d.depth.height = d.color.height = H;
        d.depth.fx = d.depth.fy = d.color.fx = d.color.fy = W / 2.f;
        d.depth.ppx = d.color.ppx = (W - 1) / 2.f; d.depth.ppy = d.color.ppy = (H - 1) / 2.f;
        d.d2c_rotation[0] = d.d2c_rotation[4] = d.d2c_rotation[8] = 1.f;
        d.d2c_translation[0] = 0.015f;
        d.depth_scale = 0.001f;
        d.color_bpp = 3;
        d.color_stride = W * 3;
        std::memcpy(d.tf, TF, sizeof d.tf);
        d.tf[3] += 0.25f * c;
        PCS(ctx, pcs_b200_set_stream(ctx, c, &d));
        std::memset(&jobs[c], 0, sizeof jobs[c]);
        jobs[c].stream = c;
        jobs[c].z16_dev = static_cast<const uint16_t *>(d_z[c]);       // the peer's memory for its cameras
        jobs[c].color_dev = static_cast<const uint8_t *>(d_col[c]);
        jobs[c].payload_dev = reinterpret_cast<int16_t *>(d_stitched + 16 + (size_t)c * N * 10);
    }
    if (!barrier(fd)) return 2;                     // the peer's frames are uploaded
    pcs_batch *batch = nullptr;
    PCS(ctx, pcs_b200_batch_create(ctx, jobs.data(), CAMS, &batch));
    for (int rep = 0; rep < 3; ++rep) PCS(ctx, pcs_b200_batch_run(ctx, batch, nullptr));
    PCS(ctx, pcs_b200_synchronize(ctx, nullptr));
    if (!barrier(fd)) return 2;                     // nobody reads my frames any more
    std::vector<int16_t> got((size_t)CAMS * N * 5);
    CK(cudaMemcpy(got.data(), d_stitched + 16, payload, cudaMemcpyDeviceToHost));
    const bool ok = std::memcmp(got.data(), want.data(), payload) == 0;
    std::printf("[rank %d] GPU %d: %d own + %d pulled cameras -> %zu stitched bytes: %s\n", rank, rank, mine, CAMS - mine,
                payload, ok ? "bit-exact" : "MISMATCH");
    pcs_b200_batch_destroy(ctx, batch);
    for (int c = FIRST[peer]; c < FIRST[peer + 1]; ++c) {
        PCS(ctx, pcs_b200_ipc_close(ctx, const_cast<void *>(d_z[c])));
        PCS(ctx, pcs_b200_ipc_close(ctx, const_cast<void *>(d_col[c])));
    }
    if (!barrier(fd)) return 2;                     // the peer has unmapped my frames: safe to free
    cudaFree(frames);
    cudaFree(d_stitched);
    pcs_b200_destroy(ctx);
    return ok ? 0 : 1;
}

int main() {
    // fork BEFORE anything touches CUDA: a CUDA context does not survive fork()
    int sv[2];
    if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) { std::perror("socketpair"); return 2; }
    std::fflush(stdout);
    const pid_t pid = fork();
    if (pid < 0) { std::perror("fork"); return 2; }
    const int rank = pid == 0 ? 1 : 0;
    close(sv[rank == 0 ? 1 : 0]);
    const int fd = sv[rank == 0 ? 0 : 1];
    int n_gpus = 0;
    if (cudaGetDeviceCount(&n_gpus) != cudaSuccess || n_gpus < 2) {
        if (rank == 0) {
            int st = 0;
            waitpid(pid, &st, 0);
            std::printf("SKIP: needs two CUDA devices\nOK\n");
        }
        return 0;
    }
    const int rc = run(rank, fd);
    std::fflush(stdout);
    if (rank == 1) _exit(rc);
    int st = 0;
    waitpid(pid, &st, 0);
    const int child = WIFEXITED(st) ? WEXITSTATUS(st) : 3;
    std::printf((rc == 0 && child == 0) ? "OK\n" : "FAILED (rank 0: %d, rank 1: %d)\n", rc, child);
    return (rc == 0 && child == 0) ? 0 : 1;
}
