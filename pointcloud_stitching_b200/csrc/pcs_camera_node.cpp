// pcs_camera_node -- the camera tier of the reference (pcs-camera-optimized's live loop,
// src/pcs-camera-optimized.cpp:159-215) re-hosted on the C ABI: listen on a TCP port, wait for
// the stitcher's one-byte 'Z' pull (:180-184), turn the next depth + colour frame into the
// reference's camera buffer on the GPU and send [int32 bytes][records] (:715-720).  Frames come
// from raw files (the reference's .bag replay needs librealsense); with --push it sends every
// frame without waiting for pulls, like `pcs-camera-optimized -f X -s` (:264-302).
//
//   pcs_camera_node --depth d.raw --color c.raw --w 1280 --h 720 --frames 4 [--port 8000]
//                   [--tx 0.015] [--tf k | --tf-file transforms.json --camera NAME] [--push] [--loops n]
//
// Host code only: everything per-frame happens inside pcs_b200::sendXYZRGBPointcloudFused.
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/pcs_b200_shim.hpp"

namespace {

struct Frame {   // the accessors the shim needs (what rs2::video_frame / rs2::depth_frame offer)
    const void *d; int w, h, bpp, stride;
    const void *get_data() const { return d; }
    int get_width() const { return w; }
    int get_height() const { return h; }
    int get_bytes_per_pixel() const { return bpp; }
    int get_stride_in_bytes() const { return stride; }
};

// src/pcs-camera-optimized.cpp:64-67 and src/pcs-multicamera-optimized.cpp:417 (k = 0)
const float TF[2][16] = {
    {-0.99977970f, 0.00926272f, 0.01883480f, 0.f, -0.01638983f, 0.21604544f, -0.97624574f, 3.416f,
     -0.01311186f, -0.97633937f, -0.21584603f, 1.802f, 0.f, 0.f, 0.f, 1.f},
    {1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 0.f, 1.f}};

bool read_file(const std::string &path, std::vector<uint8_t> &out, size_t bytes) {
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    out.resize(bytes);
    const size_t got = fread(out.data(), 1, bytes, f);
    fclose(f);
    return got == bytes;
}

}  // namespace

int main(int argc, char **argv) {
    std::string depth_path, color_path, tf_file, camera;
    int w = 1280, h = 720, frames = 1, port = 8000, tf = 0, loops = 1 << 30;
    float tx = 0.f;
    bool push = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? argv[++i] : (char *)""; };
        if (a == "--depth") depth_path = next();
        else if (a == "--color") color_path = next();
        else if (a == "--w") w = atoi(next());
        else if (a == "--h") h = atoi(next());
        else if (a == "--frames") frames = atoi(next());
        else if (a == "--port") port = atoi(next());
        else if (a == "--tx") tx = (float)atof(next());
        else if (a == "--tf") tf = atoi(next());
        else if (a == "--tf-file") tf_file = next();      // written by pointcloud_stitching_b200/calibration.py
        else if (a == "--camera") camera = next();
        else if (a == "--loops") loops = atoi(next());
        else if (a == "--push") push = true;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    signal(SIGPIPE, SIG_IGN);
    const size_t n = (size_t)w * h;
    std::vector<uint8_t> depth, color;
    if (!read_file(depth_path, depth, n * 2 * frames) || !read_file(color_path, color, n * 3 * frames)) {
        fprintf(stderr, "cannot read %zu depth / %zu colour bytes\n", n * 2 * frames, n * 3 * frames);
        return 2;
    }
    try {
        pcs_b200::Context ctx(1);
        pcs_intrinsics in = {w, h, (w - 1) / 2.f, (h - 1) / 2.f, w / 2.f, w / 2.f};
        const float rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, tr[3] = {tx, 0.f, 0.f};
        float tf_loaded[16];
        const float *tf_mat = TF[tf ? 1 : 0];
        if (!tf_file.empty()) {
            if (!pcs_b200::load_transform(tf_file, camera, tf_loaded)) {
                fprintf(stderr, "no transform for camera '%s' in %s\n", camera.c_str(), tf_file.c_str());
                return 2;
            }
            tf_mat = tf_loaded;
        }
        ctx.set_stream(0, pcs_b200::make_stream_desc(in, in, rot, tr, 0.001f, 3, w * 3, tf_mat, false));

        // initSocket (:75-105): bind, listen, accept exactly one client
        int srv = socket(AF_INET, SOCK_STREAM, IPPROTO_TCP), one = 1;
        setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        sockaddr_in addr;
        memset(&addr, 0, sizeof addr);
        addr.sin_family = AF_INET;
        addr.sin_addr.s_addr = INADDR_ANY;
        addr.sin_port = htons((uint16_t)port);
        if (bind(srv, (sockaddr *)&addr, sizeof addr) < 0 || listen(srv, 3) < 0) { perror("bind/listen"); return 1; }
        printf("pcs_camera_node: waiting for client on :%d\n", port);
        fflush(stdout);
        int client = accept(srv, nullptr, nullptr);
        if (client < 0) { perror("accept"); return 1; }

        std::vector<short> buffer(PCS_B200_CAMERA_BUF_SHORTS);   // short[BUF_SIZE] (:157)
        long sent = 0;
        for (int f = 0; f < loops; ++f) {
            if (!push) {
                char z = 0;
                if (recv(client, &z, 1, 0) <= 0) break;          // client gone
                if (z != 'Z') { fprintf(stderr, "Faulty pull request\n"); break; }
            }
            const int k = f % frames;
            Frame d = {depth.data() + (size_t)k * n * 2, w, h, 2, w * 2};
            Frame c = {color.data() + (size_t)k * n * 3, w, h, 3, w * 3};
            int size = pcs_b200::sendXYZRGBPointcloudFused(ctx.get(), 0, d, c, buffer.data(), true, client);
            if (size < 0) { fprintf(stderr, "pcs error: %s\n", pcs_b200_last_error(ctx.get())); return 1; }
            ++sent;
        }
        printf("pcs_camera_node: %ld frames sent\n", sent);
        // Half-close and drain: a push-mode camera never reads the stitcher's pulls, and close() with unread bytes
        // in the receive queue resets the connection -- the peer would lose the frames it has not read yet.
        shutdown(client, SHUT_WR);
        timeval tv = {5, 0};
        setsockopt(client, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
        char drain[256];
        while (recv(client, drain, sizeof drain, 0) > 0) {}
        close(client);
        close(srv);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
