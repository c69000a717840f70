// C++ host for the multi-GPU stitch, the way a maintainer of pcs-multicamera-client would write
// it: ONE process, one pcs_ctx per GPU, frames in ordinary cudaMalloc memory.  Cameras are dealt to
// the GPUs in contiguous blocks; after pcs_b200_enable_peer every GPU runs the fused kernel over ALL
// cameras (its own frames from HBM, the others' over NVLink -- the "pull" exchange), so every GPU
// ends with the whole stitched buffer [int32 bytes][cam0 records][cam1 records]...
// (src/pcs-multicamera-client.cpp:385-395) and the voxel merge is then sharded by z-slab.
// Everything is compared with the oracle restatement.  Built and run by tests/test_shim.py.
// Exit code 0 = bit-exact (or fewer than two GPUs: prints SKIP).
#include <cuda_runtime.h>

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
            std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)
#define PCS(ctx, call)                                                                        \
    do {                                                                                      \
        int rc_ = (call);                                                                     \
        if (rc_ < 0) {                                                                        \
            std::printf("pcs error %d (%s) at %s:%d\n", rc_, pcs_b200_last_error(ctx), __FILE__, __LINE__); \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

static uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main() {
    int n_gpus = 0;
    if (cudaGetDeviceCount(&n_gpus) != cudaSuccess || n_gpus < 2) {
        std::printf("SKIP: needs two CUDA devices\nOK\n");
        return 0;
    }
    n_gpus = 2;
    const int W = 256, H = 48, N = W * H, CAMS = 5, LEAF = 10;     // ragged: 3 + 2 cameras
    static const float tf[16] = {-0.99977970f, 0.00926272f, 0.01883480f, 0.f, -0.01638983f, 0.21604544f,
                                 -0.97624574f, 3.416f, -0.01311186f, -0.97633937f, -0.21584603f, 1.802f,
                                 0.f, 0.f, 0.f, 1.f};
    pcs_oracle_calib cal;
    std::memset(&cal, 0, sizeof cal);
    cal.depth.width = cal.color.width = W; cal.depth.height = cal.color.height = H;
    cal.depth.fx = cal.depth.fy = cal.color.fx = cal.color.fy = W / 2.f;
    cal.depth.ppx = cal.color.ppx = (W - 1) / 2.f; cal.depth.ppy = cal.color.ppy = (H - 1) / 2.f;
    cal.rotation[0] = cal.rotation[4] = cal.rotation[8] = 1.f;
    cal.translation[0] = 0.015f;
    cal.depth_scale = 0.001f;

    // frames + the oracle's stitched payload
    std::vector<std::vector<uint16_t>> z(CAMS, std::vector<uint16_t>(N));
    std::vector<std::vector<uint8_t>> col(CAMS, std::vector<uint8_t>((size_t)N * 3));
    std::vector<int16_t> want((size_t)CAMS * N * 5);
    uint32_t seed = 777;
    for (int c = 0; c < CAMS; ++c) {
        float tfc[16];
        std::memcpy(tfc, tf, sizeof tfc);
        tfc[3] += 0.25f * c;      // every camera its own pose
        for (int i = 0; i < N; ++i) z[c][i] = (lcg(seed) % 16 == 0) ? 0 : (uint16_t)(300 + lcg(seed) % 5700);
        for (size_t i = 0; i < col[c].size(); ++i) col[c][i] = (uint8_t)lcg(seed);
        std::vector<float> xyz((size_t)N * 3), uv((size_t)N * 2);
        pcs_oracle_deproject(&cal, z[c].data(), xyz.data(), uv.data(), 1);
        if (pcs_oracle_pack_simd(xyz.data(), uv.data(), N, col[c].data(), W, H, 3, W * 3, tfc, 0,
                                 want.data() + (size_t)c * N * 5) != N) {
            std::printf("oracle pack failed\n");
            return 2;
        }
    }
    std::vector<int16_t> want_vox((size_t)CAMS * N * 5);
    const int want_nv = pcs_oracle_voxel_merge(want.data(), CAMS * N, LEAF, want_vox.data());

    // cameras -> GPUs in contiguous blocks; each GPU holds its own cameras' frames
    const int first_cam[3] = {0, 3, CAMS};
    uint16_t *d_z[CAMS];
    uint8_t *d_col[CAMS];
    for (int g = 0; g < n_gpus; ++g) {
        CK(cudaSetDevice(g));
        for (int c = first_cam[g]; c < first_cam[g + 1]; ++c) {
            CK(cudaMalloc(&d_z[c], (size_t)N * 2));
            CK(cudaMalloc(&d_col[c], (size_t)N * 3));
            CK(cudaMemcpy(d_z[c], z[c].data(), (size_t)N * 2, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(d_col[c], col[c].data(), (size_t)N * 3, cudaMemcpyHostToDevice));
        }
    }
    int failures = 0;
    pcs_ctx *ctx[2] = {nullptr, nullptr};
    uint8_t *d_stitched[2] = {nullptr, nullptr};
    int16_t *d_vox[2] = {nullptr, nullptr};
    const size_t payload = (size_t)CAMS * N * 10;
    for (int g = 0; g < n_gpus; ++g) {
        pcs_config cfg;
        std::memset(&cfg, 0, sizeof cfg);
        cfg.device = g;
        cfg.max_streams = CAMS;
        PCS(nullptr, pcs_b200_create(&cfg, &ctx[g]));
        PCS(ctx[g], pcs_b200_enable_peer(ctx[g], 1 - g));
        CK(cudaSetDevice(g));
        CK(cudaMalloc(&d_stitched[g], 16 + payload));      // records at +16 (16-byte aligned), header at +12
        CK(cudaMalloc(&d_vox[g], payload));
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
            std::memcpy(d.tf, tf, sizeof d.tf);
            d.tf[3] += 0.25f * c;
            PCS(ctx[g], pcs_b200_set_stream(ctx[g], c, &d));
            std::memset(&jobs[c], 0, sizeof jobs[c]);
            jobs[c].stream = c;
            jobs[c].z16_dev = d_z[c];                  // a peer's memory for the other GPU's cameras
            jobs[c].color_dev = d_col[c];
            jobs[c].payload_dev = reinterpret_cast<int16_t *>(d_stitched[g] + 16 + (size_t)c * N * 10);
        }
        pcs_batch *batch = nullptr;
        PCS(ctx[g], pcs_b200_batch_create(ctx[g], jobs.data(), CAMS, &batch));
        PCS(ctx[g], pcs_b200_batch_run(ctx[g], batch, nullptr));
        PCS(ctx[g], pcs_b200_synchronize(ctx[g], nullptr));
        pcs_b200_batch_destroy(ctx[g], batch);
    }
    // every GPU holds the whole stitched payload; each merges its z-slab of the voxel grid
    std::vector<int16_t> vox_all;
    for (int g = 0; g < n_gpus; ++g) {
        CK(cudaSetDevice(g));
        std::vector<int16_t> got((size_t)CAMS * N * 5);
        CK(cudaMemcpy(got.data(), d_stitched[g] + 16, payload, cudaMemcpyDeviceToHost));
        if (std::memcmp(got.data(), want.data(), payload) != 0) {
            std::printf("FAIL stitched payload on GPU %d\n", g);
            ++failures;
        }
        int32_t splits[3], pts[2];
        const int16_t *rec = reinterpret_cast<const int16_t *>(d_stitched[g] + 16);
        PCS(ctx[g], pcs_b200_voxel_slab_plan_dev(ctx[g], rec, CAMS * N, LEAF, n_gpus, splits, pts, nullptr));
        const int nv = pcs_b200_voxel_merge_slab_dev(ctx[g], rec, CAMS * N, LEAF, splits[g], splits[g + 1], d_vox[g], nullptr);
        PCS(ctx[g], nv);
        std::vector<int16_t> slab((size_t)nv * 5);
        CK(cudaMemcpy(slab.data(), d_vox[g], (size_t)nv * 10, cudaMemcpyDeviceToHost));
        vox_all.insert(vox_all.end(), slab.begin(), slab.end());
        std::printf("GPU %d: stitched %zu bytes, slab [%d, %d) -> %d voxels from %d points\n", g, payload, splits[g],
                    splits[g + 1], nv, pts[g]);
    }
    if ((int)vox_all.size() != want_nv * 5 || std::memcmp(vox_all.data(), want_vox.data(), vox_all.size() * 2) != 0) {
        std::printf("FAIL voxel slabs: %zu voxels, want %d\n", vox_all.size() / 5, want_nv);
        ++failures;
    }
    for (int g = 0; g < n_gpus; ++g) pcs_b200_destroy(ctx[g]);
    std::printf(failures ? "FAILED\n" : "OK\n");
    return failures ? 1 : 0;
}
