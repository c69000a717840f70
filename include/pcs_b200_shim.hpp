// pcs_b200_shim.hpp -- the reference's call sites, re-pointed at the C ABI.
//
// Header-only C++11 (the reference builds with -std=c++11, CMakeLists.txt:20).  The
// functions keep the reference's names, argument order and return values
// (src/pcs-camera-optimized.cpp:107,363,669) and are templated on the frame types so
// that this header needs no librealsense: anything with the accessors the reference
// itself uses (get_data / get_width / get_height / get_bytes_per_pixel /
// get_stride_in_bytes / get_vertices / get_texture_coordinates / size) works --
// rs2::video_frame, rs2::depth_frame, rs2::points included.  See INTEGRATION.md.
#pragma once
#include <sys/socket.h>

#include <cstdio>
#include <cstdlib>
#include <cctype>
#include <cstring>
#include <stdexcept>
#include <string>

#include "pcs_b200.h"

namespace pcs_b200 {

// Owns a pcs_ctx (RAII); one per process, shared by all camera streams.
class Context {
public:
    explicit Context(int max_streams = 1, int device = 0, int kernel_variant = 0) {
        pcs_config cfg;
        std::memset(&cfg, 0, sizeof cfg);
        cfg.device = device;
        cfg.max_streams = max_streams;
        cfg.kernel_variant = kernel_variant;
        int rc = pcs_b200_create(&cfg, &ctx_);
        if (rc != PCS_OK) throw std::runtime_error(std::string("pcs_b200_create: ") + pcs_b200_last_error(nullptr));
    }
    ~Context() { pcs_b200_destroy(ctx_); }
    Context(const Context &) = delete;
    Context &operator=(const Context &) = delete;
    pcs_ctx *get() const { return ctx_; }
    void set_stream(int stream, const pcs_stream_desc &d) {
        if (pcs_b200_set_stream(ctx_, stream, &d) != PCS_OK)
            throw std::runtime_error(std::string("pcs_b200_set_stream: ") + pcs_b200_last_error(ctx_));
    }

private:
    pcs_ctx *ctx_ = nullptr;
};

// rs2_intrinsics -> pcs_intrinsics, distortion included (templated so that this header does not need librealsense2:
// any struct with rs2_intrinsics' members works).  librealsense's rs2_distortion numbers NONE = 0, MODIFIED_BROWN_CONRADY
// = 1, INVERSE_BROWN_CONRADY = 2, BROWN_CONRADY = 4 -- the values of PCS_B200_DISTORTION_*; a model with all-zero
// coefficients maps to none, everything else passes through: the library applies a model where rsutil.h does (inverse
// Brown-Conrady when deprojecting, modified Brown-Conrady when projecting), ignores it where rsutil.h does, and
// pcs_b200_set_stream refuses the fisheye models.
template <class Rs2Intrinsics>
inline pcs_intrinsics intrinsics_from_rs2(const Rs2Intrinsics &in) {
    pcs_intrinsics o;
    std::memset(&o, 0, sizeof o);
    o.width = in.width; o.height = in.height;
    o.ppx = in.ppx; o.ppy = in.ppy; o.fx = in.fx; o.fy = in.fy;
    const int model = (int)in.model;
    bool any = false;
    for (int i = 0; i < 5; ++i) any = any || in.coeffs[i] != 0.f;
    o.model = any ? model : PCS_B200_DISTORTION_NONE;
    if (o.model != PCS_B200_DISTORTION_NONE)
        for (int i = 0; i < 5; ++i) o.coeffs[i] = in.coeffs[i];
    return o;
}

// A stream descriptor with the reference's constants filled in: tf_mat
// (src/pcs-camera-optimized.cpp:64-67 layout: row-major 4x4) and the -c bounds (:398-401).
inline pcs_stream_desc make_stream_desc(const pcs_intrinsics &depth, const pcs_intrinsics &color,
                                        const float rotation_colmajor[9], const float translation[3],
                                        float depth_scale, int color_bpp, int color_stride,
                                        const float tf_mat[16], bool cutoff) {
    pcs_stream_desc d;
    std::memset(&d, 0, sizeof d);
    d.depth = depth;
    d.color = color;
    std::memcpy(d.d2c_rotation, rotation_colmajor, sizeof d.d2c_rotation);
    std::memcpy(d.d2c_translation, translation, sizeof d.d2c_translation);
    d.depth_scale = depth_scale;
    d.color_bpp = color_bpp;
    d.color_stride = color_stride;
    std::memcpy(d.tf, tf_mat, sizeof d.tf);
    d.cutoff = cutoff ? 1 : 0;
    d.z_lo = 0.f; d.z_hi = 1.5f; d.x_lo = -2.f; d.x_hi = 2.f;
    d.cutoff_lane_reversed = 1;   // what the reference's -m -c actually does
    return d;
}

// The camera -> world transform of one camera from a calibration file instead of a constant pasted
// into the source (the reference: calibration/camera_alignment.py prints `transform[k] << ...` for
// src/pcs-multicamera-optimized.cpp:417-455 and tf_mat, src/pcs-camera-optimized.cpp:64-67 -- a
// recompile per rig).  The file is what pointcloud_stitching_b200/calibration.py writes:
//     { "CAMERA": [[r00, r01, r02, tx], [r10, ...], [r20, ...], [0, 0, 0, 1]], ... }
// Returns false when the file cannot be read or does not hold 16 numbers under that name.
inline bool load_transform(const std::string &path, const std::string &camera, float tf[16]) {
    std::FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) return false;
    std::string text;
    char buf[4096];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
    std::fclose(f);
    const std::string key = "\"" + camera + "\"";
    size_t at = 0;
    while ((at = text.find(key, at)) != std::string::npos) {
        size_t p = at + key.size();
        while (p < text.size() && std::isspace((unsigned char)text[p])) ++p;
        if (p >= text.size() || text[p] != ':') { at = p; continue; }      // a value that happens to match, not a key
        ++p;
        // the value: numbers inside brackets (4 rows of 4, 3 rows of 4, or flat), nothing else; it ends where the
        // brackets close, so a short matrix never borrows numbers from the next entry
        int n = 0, depth = 0;
        bool opened = false, bad = false;
        float v16[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1};
        while (p < text.size() && !(opened && depth == 0)) {
            const char c = text[p];
            if (std::isspace((unsigned char)c) || c == ',') { ++p; continue; }
            if (c == '[') { ++depth; opened = true; ++p; continue; }
            if (c == ']') { --depth; ++p; continue; }
            char *end = nullptr;
            const double v = std::strtod(text.c_str() + p, &end);
            if (!opened || end == text.c_str() + p || n == 16 || !(v == v)) { bad = true; break; }
            v16[n++] = (float)v;
            p = (size_t)(end - text.c_str());
        }
        if (!bad && opened && depth == 0 && (n == 16 || n == 12)) {     // 12: the last row (0 0 0 1) is implied
            std::memcpy(tf, v16, sizeof v16);
            return true;
        }
        return false;
    }
    return false;
}

// int copyPointCloudXYZRGBToBufferSIMD(rs2::points&, const rs2::video_frame&, short*)
// (src/pcs-camera-optimized.cpp:363): same inputs, same record bytes, returns the point count.
template <class Points, class VideoFrame>
int copyPointCloudXYZRGBToBufferSIMD(pcs_ctx *ctx, int stream, Points &pts, const VideoFrame &color,
                                     short *pc_buffer) {
    return pcs_b200_pack_from_vertices(ctx, stream, reinterpret_cast<const float *>(pts.get_vertices()),
                                       reinterpret_cast<const float *>(pts.get_texture_coordinates()),
                                       (int)pts.size(), static_cast<const uint8_t *>(color.get_data()),
                                       pc_buffer);
}

// int sendXYZRGBPointcloud(rs2::points pts, rs2::video_frame color, short *buffer)
// (src/pcs-camera-optimized.cpp:669-723), vertices in: memset, pack at byte 4, optional
// header + send().  Returns the payload size in bytes.
template <class Points, class VideoFrame>
int sendXYZRGBPointcloud(pcs_ctx *ctx, int stream, Points pts, VideoFrame color, short *buffer,
                         bool send_buffer, int client_sock) {
    std::memset(buffer, 0, PCS_B200_CAMERA_BUF_SHORTS);                       // :673 (BUF_SIZE bytes)
    int size = copyPointCloudXYZRGBToBufferSIMD(ctx, stream, pts, color, &buffer[0] + sizeof(short));  // :690
    if (size < 0) return size;
    size = 5 * size * (int)sizeof(short);                                     // :697
    if (send_buffer) {
        std::memcpy(buffer, &size, sizeof(int));                              // :718
        send(client_sock, (char *)buffer, size + sizeof(int), 0);             // :719
    }
    return size;
}

// The fused form: the depth frame goes in instead of rs2::pointcloud::calculate()'s output
// (replaces :288-292 in one call).  Same buffer image, same return value.
template <class DepthFrame, class VideoFrame>
int sendXYZRGBPointcloudFused(pcs_ctx *ctx, int stream, const DepthFrame &depth, const VideoFrame &color,
                              short *buffer, bool send_buffer, int client_sock) {
    int size = pcs_b200_send_xyzrgb(ctx, stream, static_cast<const uint16_t *>(depth.get_data()),
                                    static_cast<const uint8_t *>(color.get_data()), buffer, send_buffer ? 1 : 0);
    if (size < 0) return size;
    if (send_buffer) send(client_sock, (char *)buffer, size + sizeof(int), 0);
    return size;
}

}  // namespace pcs_b200
