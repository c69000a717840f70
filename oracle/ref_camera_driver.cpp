// TEST INFRASTRUCTURE ONLY.  Linked with the reference's own
// src/pcs-camera-optimized.cpp (compiled UNMODIFIED from /root/reference with
// -Dmain=ref_main, see oracle/Makefile) into oracle/_ref/libpcs_ref_camera.so.
// Exposes the reference's hot functions (src/pcs-camera-optimized.cpp:363,620,669)
// through a C ABI so tests can pin oracle/pcs_oracle.c against them and
// bench.py can time them as the CPU baseline.  Not used by the product.
#include <librealsense2/rs.hpp>

#include <immintrin.h>
#include <sys/socket.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <iostream>
#include <sstream>
#include <thread>
#include <vector>

// ---- symbols defined by the reference TU -----------------------------------
extern bool use_simd, cutoff, send_buffer, initialized, display_updates;
extern int num_of_threads, client_sock;
extern float tf_mat[];
extern __m128 ss_a, ss_b, ss_c, ss_d;
int copyPointCloudXYZRGBToBufferSIMD(rs2::points &pts, const rs2::video_frame &color, short *pc_buffer);
int copyPointCloudXYZRGBToBuffer(rs2::points &pts, const rs2::video_frame &color, short *pc_buffer);
int sendXYZRGBPointcloud(rs2::points pts, rs2::video_frame color, short *buffer);
int ref_main(int argc, char **argv);

// ---- stub librealsense back end ---------------------------------------------
namespace pcs_stub {
session &current() { static session s; return s; }
}

namespace rs2 {
points pointcloud::calculate(const depth_frame &depth) {
    pcs_stub::session &s = pcs_stub::current();
    const size_t n = (size_t)depth.get_width() * depth.get_height();
    v_.resize(n);
    t_.resize(n);
    auto t0 = std::chrono::high_resolution_clock::now();
    pcs_oracle_deproject(&s.calib, static_cast<const uint16_t *>(depth.get_data()),
                         reinterpret_cast<float *>(v_.data()), reinterpret_cast<float *>(t_.data()),
                         num_of_threads);
    auto t1 = std::chrono::high_resolution_clock::now();
    s.calculate_ms += std::chrono::duration<double, std::milli>(t1 - t0).count();
    s.calculate_calls++;
    return points(v_.data(), t_.data(), n);
}

// Frames are numbered 1..n; after the last one the number wraps to 1, which is
// how the reference's replay loop detects end of file (:276).
frameset pipeline::wait_for_frames() {
    pcs_stub::session &s = pcs_stub::current();
    const int dw = s.calib.depth.width, dh = s.calib.depth.height;
    const int cw = s.calib.color.width, ch = s.calib.color.height;
    const int f = s.cursor % s.n_frames;
    s.cursor++;
    depth_frame d(s.depth + (size_t)f * dw * dh, dw, dh);
    video_frame c(s.color + (size_t)f * ch * s.color_stride, cw, ch, s.color_bpp, s.color_stride);
    return frameset(c, d, (unsigned long long)f + 1);
}
}  // namespace rs2

static void set_transform(const float *tf) {
    std::memcpy(tf_mat, tf, 16 * sizeof(float));
    // the reference builds these once at load time (:69-72): column k of the 3x4
    ss_a = _mm_set_ps(0, tf_mat[8], tf_mat[4], tf_mat[0]);
    ss_b = _mm_set_ps(0, tf_mat[9], tf_mat[5], tf_mat[1]);
    ss_c = _mm_set_ps(0, tf_mat[10], tf_mat[6], tf_mat[2]);
    ss_d = _mm_set_ps(0, tf_mat[11], tf_mat[7], tf_mat[3]);
}

extern "C" {

// copyPointCloudXYZRGBToBuffer[SIMD] on caller-provided vertices / tex coords.
int ref_pack(int simd, const float *xyz, const float *uv, int n, const uint8_t *color, int cw,
             int ch, int bpp, int stride, const float *tf, int cut, int threads, int16_t *out) {
    set_transform(tf);
    use_simd = simd != 0;
    cutoff = cut != 0;
    num_of_threads = threads;
    initialized = false;  // geometry is cached behind this flag (:349,369)
    rs2::points pts(reinterpret_cast<const rs2::vertex *>(xyz),
                    reinterpret_cast<const rs2::texture_coordinate *>(uv), (size_t)n);
    rs2::video_frame col(color, cw, ch, bpp, stride);
    return simd ? copyPointCloudXYZRGBToBufferSIMD(pts, col, out)
                : copyPointCloudXYZRGBToBuffer(pts, col, out);
}

// sendXYZRGBPointcloud into `buffer` (short[5 000 000], like :157).  With
// wire != 0 the -s path runs against a socketpair and the bytes that reached the
// peer are copied to wire_out (capacity wire_cap); returns the function's result.
int ref_send(const float *xyz, const float *uv, int n, const uint8_t *color, int cw, int ch,
             int bpp, int stride, const float *tf, int cut, int threads, int16_t *buffer,
             int wire, uint8_t *wire_out, int wire_cap, int *wire_len) {
    set_transform(tf);
    use_simd = true;
    cutoff = cut != 0;
    num_of_threads = threads;
    initialized = false;
    rs2::points pts(reinterpret_cast<const rs2::vertex *>(xyz),
                    reinterpret_cast<const rs2::texture_coordinate *>(uv), (size_t)n);
    rs2::video_frame col(color, cw, ch, bpp, stride);
    if (!wire) {
        send_buffer = false;
        return sendXYZRGBPointcloud(pts, col, buffer);
    }
    int sv[2];
    if (socketpair(AF_UNIX, SOCK_STREAM, 0, sv) != 0) return -1;
    int got = 0;
    std::thread reader([&] {
        for (;;) {
            uint8_t tmp[65536];
            ssize_t r = read(sv[1], tmp, sizeof tmp);
            if (r <= 0) break;
            if (got + r <= wire_cap) std::memcpy(wire_out + got, tmp, (size_t)r);
            got += (int)r;
        }
    });
    client_sock = sv[0];
    send_buffer = true;
    int ret = sendXYZRGBPointcloud(pts, col, buffer);
    send_buffer = false;
    close(sv[0]);
    reader.join();
    close(sv[1]);
    client_sock = 0;
    if (wire_len) *wire_len = got;
    return ret;
}

// Times `iters` calls of sendXYZRGBPointcloud exactly as the reference does
// (:291-293: clock around the call only).  ms_out[iters].
int ref_time_send(const float *xyz, const float *uv, int n, const uint8_t *color, int cw, int ch,
                  int bpp, int stride, const float *tf, int simd, int threads, int16_t *buffer,
                  int iters, double *ms_out) {
    set_transform(tf);
    use_simd = simd != 0;
    cutoff = false;
    send_buffer = false;
    num_of_threads = threads;
    initialized = false;
    rs2::points pts(reinterpret_cast<const rs2::vertex *>(xyz),
                    reinterpret_cast<const rs2::texture_coordinate *>(uv), (size_t)n);
    rs2::video_frame col(color, cw, ch, bpp, stride);
    int ret = 0;
    for (int i = 0; i < iters; ++i) {
        auto t0 = std::chrono::high_resolution_clock::now();
        ret = sendXYZRGBPointcloud(pts, col, buffer);
        auto t1 = std::chrono::high_resolution_clock::now();
        ms_out[i] = std::chrono::duration<double, std::milli>(t1 - t0).count();
    }
    return ret;
}

// Runs the reference's own main() replay loop (:216-343) over n_frames synthetic
// frames: `pcs-camera-optimized -f synthetic -m -t threads`.  Returns the
// "AVG Frame Time" it prints (:322) in *avg_ms, the stub-calculate average in
// *calc_ms, and the captured stdout in log (capacity log_cap).
int ref_replay(const uint16_t *depth, const uint8_t *color, int n_frames,
               const pcs_oracle_calib *calib, int bpp, int stride, const float *tf, int simd,
               int threads, double *avg_ms, double *calc_ms, char *log, int log_cap) {
    pcs_stub::session &s = pcs_stub::current();
    s = pcs_stub::session();
    s.depth = depth;
    s.color = color;
    s.n_frames = n_frames;
    s.calib = *calib;
    s.color_bpp = bpp;
    s.color_stride = stride;
    set_transform(tf);
    initialized = false;
    use_simd = false;
    cutoff = false;
    send_buffer = false;
    std::string t = std::to_string(threads);
    std::vector<std::string> args = {"pcs-camera-optimized", "-f", "synthetic", "-t", t};
    if (simd) args.push_back("-m");
    std::vector<char *> argv;
    for (auto &a : args) argv.push_back(&a[0]);
    argv.push_back(nullptr);
    optind = 1;
    std::ostringstream cap;
    std::streambuf *old = std::cout.rdbuf(cap.rdbuf());
    int rc = ref_main((int)args.size(), argv.data());
    std::cout.rdbuf(old);
    const std::string out = cap.str();
    if (log && log_cap > 0) {
        // keep the tail (the summary block)
        const size_t keep = out.size() < (size_t)log_cap - 1 ? out.size() : (size_t)log_cap - 1;
        std::memcpy(log, out.data() + (out.size() - keep), keep);
        log[keep] = 0;
    }
    const char *key = "### AVG Frame Time: ";
    size_t p = out.find(key);
    if (avg_ms) *avg_ms = p == std::string::npos ? -1.0 : atof(out.c_str() + p + strlen(key));
    // calculate() is also called once after the loop (:315)
    if (calc_ms) *calc_ms = s.calculate_calls ? s.calculate_ms / s.calculate_calls : -1.0;
    return rc;
}

}  // extern "C"
