// TEST INFRASTRUCTURE ONLY -- a stand-in for <librealsense2/rs.hpp>.
//
// librealsense2 is an un-vendored, un-pinned apt dependency of the reference
// (/root/reference/Dockerfile:20-23) and is absent from this image.  This stub
// declares just the names the reference translation units touch, so that
// src/pcs-camera-optimized.cpp and src/pcs-multicamera-*.cpp compile UNMODIFIED
// (oracle/Makefile).  Frames are served from memory handed in by
// oracle/ref_camera_driver.cpp; rs2::pointcloud::calculate() forwards to the
// written-down deprojection spec in oracle/pcs_oracle.c (SPEC.md section 1).
// Nothing under pointcloud_stitching_b200/ includes this file.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

extern "C" {
#include "pcs_oracle.h"
}

enum rs2_option { RS2_OPTION_EMITTER_ENABLED = 0 };
enum rs2_camera_info { RS2_CAMERA_INFO_NAME = 0, RS2_CAMERA_INFO_FIRMWARE_VERSION = 1 };

namespace pcs_stub {
// One synthetic capture session; filled by the driver before ref_main() runs.
struct session {
    const uint16_t *depth = nullptr;   // [n_frames][dh][dw]
    const uint8_t *color = nullptr;    // [n_frames][ch][cstride]
    int n_frames = 0;
    pcs_oracle_calib calib{};          // intrinsics / extrinsics / depth scale
    int color_bpp = 3, color_stride = 0;
    int cursor = 0;                    // next frame to hand out
    double calculate_ms = 0.0;         // time spent inside pointcloud::calculate
    int calculate_calls = 0;
};
session &current();
}  // namespace pcs_stub

namespace rs2 {

struct vertex { float x, y, z; };
struct texture_coordinate { float u, v; };

class video_frame {
public:
    video_frame() {}
    video_frame(const void *d, int w, int h, int bpp, int stride)
        : data_(d), w_(w), h_(h), bpp_(bpp), stride_(stride) {}
    const void *get_data() const { return data_; }
    int get_width() const { return w_; }
    int get_height() const { return h_; }
    int get_bytes_per_pixel() const { return bpp_; }
    int get_stride_in_bytes() const { return stride_; }
private:
    const void *data_ = nullptr;
    int w_ = 0, h_ = 0, bpp_ = 0, stride_ = 0;
};

class depth_frame : public video_frame {
public:
    depth_frame() {}
    depth_frame(const void *d, int w, int h) : video_frame(d, w, h, 2, 2 * w) {}
};

class points {
public:
    points() {}
    points(const vertex *v, const texture_coordinate *t, size_t n) : v_(v), t_(t), n_(n) {}
    const vertex *get_vertices() const { return v_; }
    const texture_coordinate *get_texture_coordinates() const { return t_; }
    size_t size() const { return n_; }
private:
    const vertex *v_ = nullptr;
    const texture_coordinate *t_ = nullptr;
    size_t n_ = 0;
};

class frameset {
public:
    frameset() {}
    frameset(video_frame c, depth_frame d, unsigned long long n) : c_(c), d_(d), n_(n) {}
    video_frame get_color_frame() const { return c_; }
    depth_frame get_depth_frame() const { return d_; }
    unsigned long long get_frame_number() const { return n_; }
private:
    video_frame c_;
    depth_frame d_;
    unsigned long long n_ = 0;
};

class pointcloud {
public:
    points calculate(const depth_frame &depth);
    void map_to(const video_frame &) {}
private:
    std::vector<vertex> v_;
    std::vector<texture_coordinate> t_;
};

class depth_sensor {
public:
    bool supports(rs2_option) const { return false; }
    void set_option(rs2_option, float) const {}
};

class device {
public:
    template <class T> T first() const { return T(); }
    const char *get_info(rs2_camera_info i) const {
        return i == RS2_CAMERA_INFO_NAME ? "pcs-oracle synthetic camera" : "0.0";
    }
};

class pipeline_profile {
public:
    device get_device() const { return device(); }
};

class config {
public:
    void enable_device_from_file(const char *) {}
};

class pipeline {
public:
    pipeline_profile start() { return pipeline_profile(); }
    pipeline_profile start(const config &) { return pipeline_profile(); }
    pipeline_profile get_active_profile() const { return pipeline_profile(); }
    void stop() {}
    frameset wait_for_frames();
    bool poll_for_frames(frameset *out) { *out = wait_for_frames(); return true; }
};

}  // namespace rs2
