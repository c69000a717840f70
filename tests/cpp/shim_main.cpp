// C++ host-side parity check of include/pcs_b200_shim.hpp: the reference's call shapes
// (sendXYZRGBPointcloud / copyPointCloudXYZRGBToBufferSIMD) driven with rs2-like frame
// objects (the oracle's stub librealsense types), compared against the oracle restatement.
// Built and run by tests/test_shim.py.  Exit code 0 = bit-exact.
#include <librealsense2/rs.hpp>   // oracle/stubs: rs2::points, rs2::video_frame, rs2::depth_frame

#include <sys/socket.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "pcs_b200_shim.hpp"

static uint32_t lcg(uint32_t &s) { s = s * 1664525u + 1013904223u; return s >> 8; }

int main() {
    const int W = 1280, H = 720, N = W * H;
    static const float tf[16] = {-0.99977970f, 0.00926272f, 0.01883480f, 0.f, -0.01638983f, 0.21604544f,
                                 -0.97624574f, 3.416f, -0.01311186f, -0.97633937f, -0.21584603f, 1.802f,
                                 0.f, 0.f, 0.f, 1.f};
    std::vector<uint16_t> z(N);
    std::vector<uint8_t> col((size_t)N * 3);
    uint32_t seed = 12345;
    for (int i = 0; i < N; ++i) z[i] = (lcg(seed) % 16 == 0) ? 0 : (uint16_t)(300 + lcg(seed) % 5700);
    for (size_t i = 0; i < col.size(); ++i) col[i] = (uint8_t)lcg(seed);

    pcs_oracle_calib cal;
    std::memset(&cal, 0, sizeof cal);
    cal.depth.width = cal.color.width = W; cal.depth.height = cal.color.height = H;
    cal.depth.fx = cal.depth.fy = cal.color.fx = cal.color.fy = 640.f;
    cal.depth.ppx = cal.color.ppx = 639.5f; cal.depth.ppy = cal.color.ppy = 359.5f;
    cal.rotation[0] = cal.rotation[4] = cal.rotation[8] = 1.f;
    cal.translation[0] = 0.015f;
    cal.depth_scale = 0.001f;

    // what rs2::video_stream_profile::get_intrinsics() hands back (rs2_intrinsics' members); a Brown-Conrady model
    // with all-zero coefficients (every D400 depth stream) must come out as "none" and keep the fast kernels
    struct { int width, height; float ppx, ppy, fx, fy; int model; float coeffs[5]; } rs = {
        W, H, 639.5f, 359.5f, 640.f, 640.f, 4, {0.f, 0.f, 0.f, 0.f, 0.f}};
    pcs_intrinsics di = pcs_b200::intrinsics_from_rs2(rs);
    if (di.model != PCS_B200_DISTORTION_NONE || di.fx != 640.f || di.width != W) { std::printf("intrinsics_from_rs2 broken\n"); return 1; }
    rs.model = 1; rs.coeffs[0] = 0.1f;
    if (pcs_b200::intrinsics_from_rs2(rs).model != PCS_B200_DISTORTION_MODIFIED_BROWN_CONRADY ||
        pcs_b200::intrinsics_from_rs2(rs).coeffs[0] != 0.1f) { std::printf("intrinsics_from_rs2 drops the model\n"); return 1; }
    int failures = 0;
    try {
        pcs_b200::Context ctx(2);
        for (int cutoff = 0; cutoff < 2; ++cutoff) {
            ctx.set_stream(cutoff, pcs_b200::make_stream_desc(di, di, cal.rotation, cal.translation, 0.001f, 3,
                                                              W * 3, tf, cutoff != 0));
            // oracle: deproject + send
            std::vector<float> xyz((size_t)N * 3), uv((size_t)N * 2);
            pcs_oracle_deproject(&cal, z.data(), xyz.data(), uv.data(), 4);
            std::vector<short> want(5000000, 0x5A5A), got(5000000, 0x5A5A), got2(5000000, 0x5A5A);
            int want_size = pcs_oracle_send(xyz.data(), uv.data(), N, col.data(), W, H, 3, W * 3, tf, cutoff, 1,
                                            want.data());
            // 1. fused call, with the -s send path against a socketpair
            int sv[2];
            socketpair(AF_UNIX, SOCK_STREAM, 0, sv);
            std::vector<uint8_t> wire;
            std::thread reader([&] {
                uint8_t tmp[65536];
                ssize_t r;
                while ((r = read(sv[1], tmp, sizeof tmp)) > 0) wire.insert(wire.end(), tmp, tmp + r);
            });
            rs2::depth_frame depth(z.data(), W, H);
            rs2::video_frame color(col.data(), W, H, 3, W * 3);
            int size = pcs_b200::sendXYZRGBPointcloudFused(ctx.get(), cutoff, depth, color, got.data(), true, sv[0]);
            close(sv[0]);
            reader.join();
            close(sv[1]);
            if (size != want_size || std::memcmp(got.data(), want.data(), 10000000) != 0) {
                std::printf("FAIL fused cutoff=%d size %d want %d\n", cutoff, size, want_size);
                ++failures;
            }
            if ((int)wire.size() != want_size + 4 || std::memcmp(wire.data(), want.data(), wire.size()) != 0) {
                std::printf("FAIL wire bytes cutoff=%d (%zu bytes)\n", cutoff, wire.size());
                ++failures;
            }
            // 2. the reference seam itself: vertices + tex coords in
            rs2::points pts(reinterpret_cast<const rs2::vertex *>(xyz.data()),
                            reinterpret_cast<const rs2::texture_coordinate *>(uv.data()), (size_t)N);
            int size2 = pcs_b200::sendXYZRGBPointcloud(ctx.get(), cutoff, pts, color, got2.data(), false, -1);
            std::memcpy(want.data(), "\0\0\0\0", 4);   // without -s the header stays zero (:715)
            if (size2 != want_size || std::memcmp(got2.data(), want.data(), 10000000) != 0) {
                std::printf("FAIL from-vertices cutoff=%d size %d want %d\n", cutoff, size2, want_size);
                ++failures;
            }
            std::printf("cutoff=%d: %d bytes, fused + wire + from-vertices compared\n", cutoff, want_size);
        }
    } catch (const std::exception &e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 2;
    }
    std::printf(failures ? "FAILED\n" : "OK\n");
    return failures ? 1 : 0;
}
