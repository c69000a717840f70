// pcs_stitch_node -- the stitcher tier of the reference (pcs-multicamera-client without -v,
// src/pcs-multicamera-client.cpp:363-432; with --pcl: pcs-multicamera-optimized's loop,
// src/pcs-multicamera-optimized.cpp:321-401) re-hosted on the C ABI: connect to every camera on
// CLIENT_PORT + i (:553-556), accept one viewer on the server port (initServerSocket, :213-243), then
// per frame: one reader thread per camera takes [int32 bytes][payload] off its socket and asks for the
// next frame with a one-byte 'Z' (readCloud, :363-371); the payloads are joined in camera order and
// stitched ON THE GPU -- concat + stride decimation (:385-395, pcs_b200_stitch_raw) or unpack ->
// transform -> append -> repack (pcs_b200_stitch_pcl) -- into [int32 bytes][records]; the node blocks
// on the viewer's 'Z' (:398) and writes the stitched buffer (:403).
//
//   pcs_stitch_node --cameras 2 [--camera-port 8000] [--viewer-port 9000] [--downsample d] [--frames n]
//                   [--prime-pull] [--pcl [--tf-file rig.json --names A,B | --tf k0,k1,...]]
//
// --prime-pull sends the first 'Z' to every camera before reading, as visualize() (:446-449) and
// pcs-multicamera-optimized (src/pcs-multicamera-optimized.cpp:342-345) do; without it the node reads
// first, like the reference's non-visual client -- which therefore only works against a push-mode
// camera.  Host code only.
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <unistd.h>

#include <csignal>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pcs_b200_shim.hpp"

namespace {

const int BUF_SIZE = 5000000;                    // shorts per camera buffer (src/pcs-multicamera-client.cpp:44)

bool read_n(int fd, void *dst, size_t n) {       // readNBytes (:255-268) without the exit()
    uint8_t *p = static_cast<uint8_t *>(dst);
    while (n) {
        const ssize_t got = read(fd, p, n);
        if (got <= 0) return false;
        p += got;
        n -= (size_t)got;
    }
    return true;
}

bool write_n(int fd, const void *src, size_t n) {
    const uint8_t *p = static_cast<const uint8_t *>(src);
    while (n) {
        const ssize_t put = write(fd, p, n);
        if (put <= 0) return false;
        p += put;
        n -= (size_t)put;
    }
    return true;
}

int connect_camera(int port) {                   // initSocket(port, ip) (:188-211), localhost
    for (int attempt = 0; attempt < 600; ++attempt) {
        int fd = socket(AF_INET, SOCK_STREAM, 0);
        sockaddr_in a;
        memset(&a, 0, sizeof a);
        a.sin_family = AF_INET;
        a.sin_port = htons((uint16_t)port);
        inet_pton(AF_INET, "127.0.0.1", &a.sin_addr);
        if (connect(fd, (sockaddr *)&a, sizeof a) == 0) return fd;
        close(fd);
        usleep(100 * 1000);
    }
    return -1;
}

std::vector<std::string> split(const std::string &s) {
    std::vector<std::string> out;
    size_t at = 0;
    while (at <= s.size()) {
        const size_t e = s.find(',', at);
        out.push_back(s.substr(at, e == std::string::npos ? std::string::npos : e - at));
        if (e == std::string::npos) break;
        at = e + 1;
    }
    return out;
}

// the eight hard-coded transforms are the reference's (src/pcs-multicamera-optimized.cpp:417-455); the
// node only knows identity and what --tf-file gives it
const float IDENT[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};

}  // namespace

int main(int argc, char **argv) {
    int cameras = 1, camera_port = 8000, viewer_port = 9000, downsample = 1, frames = 1 << 30;
    bool prime = false, pcl = false;
    std::string tf_file, names;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() { return i + 1 < argc ? argv[++i] : (char *)""; };
        if (a == "--cameras") cameras = atoi(next());
        else if (a == "--camera-port") camera_port = atoi(next());
        else if (a == "--viewer-port") viewer_port = atoi(next());
        else if (a == "--downsample") downsample = atoi(next());      // -d (:87,117-119)
        else if (a == "--frames") frames = atoi(next());
        else if (a == "--prime-pull") prime = true;
        else if (a == "--pcl") pcl = true;
        else if (a == "--tf-file") tf_file = next();
        else if (a == "--names") names = next();
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    signal(SIGPIPE, SIG_IGN);      // a camera that has sent its last frame may be gone when the next pull is written
    if (cameras < 1 || cameras > 32 || downsample < 1) { fprintf(stderr, "1 <= cameras <= 32, downsample >= 1\n"); return 2; }
    std::vector<float> tfs((size_t)cameras * 16);
    for (int c = 0; c < cameras; ++c) memcpy(&tfs[(size_t)c * 16], IDENT, sizeof IDENT);
    if (pcl && !tf_file.empty()) {
        const std::vector<std::string> nm = split(names);
        if ((int)nm.size() != cameras) { fprintf(stderr, "--names needs one name per camera\n"); return 2; }
        for (int c = 0; c < cameras; ++c)
            if (!pcs_b200::load_transform(tf_file, nm[c], &tfs[(size_t)c * 16])) {
                fprintf(stderr, "no transform for camera '%s' in %s\n", nm[c].c_str(), tf_file.c_str());
                return 2;
            }
    }
    try {
        pcs_b200::Context ctx(1);
        // pc_buf[i]: 10 MB per camera (:554); stitched_buf sized for what the cameras can send
        std::vector<std::vector<short>> pc_buf(cameras, std::vector<short>(BUF_SIZE));
        std::vector<uint8_t> stitched((size_t)cameras * BUF_SIZE * 2 + 4);
        std::vector<int> sock(cameras, -1);
        for (int c = 0; c < cameras; ++c) {
            sock[c] = connect_camera(camera_port + c);
            if (sock[c] < 0) { fprintf(stderr, "cannot connect to camera %d on :%d\n", c, camera_port + c); return 1; }
        }
        int srv = socket(AF_INET, SOCK_STREAM, 0), one = 1;
        setsockopt(srv, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        sockaddr_in addr;
        memset(&addr, 0, sizeof addr);
        addr.sin_family = AF_INET;
        addr.sin_addr.s_addr = INADDR_ANY;
        addr.sin_port = htons((uint16_t)viewer_port);
        if (bind(srv, (sockaddr *)&addr, sizeof addr) < 0 || listen(srv, 3) < 0) { perror("bind/listen"); return 1; }
        printf("pcs_stitch_node: %d camera(s) connected, waiting for the viewer on :%d\n", cameras, viewer_port);
        fflush(stdout);
        const int viewer = accept(srv, nullptr, nullptr);
        if (viewer < 0) { perror("accept"); return 1; }
        const char pull = 'Z';
        if (prime)
            for (int c = 0; c < cameras; ++c)
                if (!write_n(sock[c], &pull, 1)) return 1;
        std::vector<int32_t> n_shorts(cameras);
        std::vector<const int16_t *> pay(cameras);
        long served = 0;
        for (int f = 0; f < frames; ++f) {
            std::vector<char> ok(cameras, 0);
            std::vector<std::thread> th;
            for (int c = 0; c < cameras; ++c)
                th.emplace_back([&, c]() {                           // readCloud (:363-371)
                    int32_t bytes = 0;
                    if (!read_n(sock[c], &bytes, 4) || bytes < 0 || bytes > BUF_SIZE * 2) return;
                    if (!read_n(sock[c], pc_buf[c].data(), (size_t)bytes)) return;
                    n_shorts[c] = bytes / 2;
                    ok[c] = write_n(sock[c], &pull, 1) ? 1 : 2;      // next pull; a push-mode camera that has
                });                                                  // finished may already have closed: fine
            bool all = true;
            for (int c = 0; c < cameras; ++c) {                      // join in camera order (:385-386)
                th[c].join();
                all = all && ok[c];
                pay[c] = pc_buf[c].data();
            }
            if (!all) break;                                         // a camera went away
            const int size = pcl ? pcs_b200_stitch_pcl(ctx.get(), pay.data(), n_shorts.data(), cameras, downsample, tfs.data(),
                                                       stitched.data(), stitched.size())
                                 : pcs_b200_stitch_raw(ctx.get(), pay.data(), n_shorts.data(), cameras, downsample,
                                                       stitched.data(), stitched.size());
            if (size < 0) { fprintf(stderr, "pcs error: %s\n", pcs_b200_last_error(ctx.get())); return 1; }
            char z = 0;
            if (read(viewer, &z, 1) <= 0) break;                     // :398
            if (z != 'Z') { fprintf(stderr, "Faulty pull request\n"); break; }
            if (!write_n(viewer, stitched.data(), (size_t)size + 4)) break;   // :403
            ++served;
        }
        printf("pcs_stitch_node: %ld stitched frames served\n", served);
        close(viewer);
        close(srv);
        for (int c = 0; c < cameras; ++c) close(sock[c]);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
