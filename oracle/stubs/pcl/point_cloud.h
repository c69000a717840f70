// TEST INFRASTRUCTURE ONLY -- a stand-in for the PCL / Eigen headers the
// reference stitcher TUs include (PCL is an un-pinned apt dependency,
// /root/reference/Dockerfile:25, absent here).  Declares only what
// src/pcs-multicamera-client.cpp and src/pcs-multicamera-optimized.cpp touch so
// they compile UNMODIFIED.  pcl::transformPointCloud forwards to the written
// spec in oracle/pcs_oracle.c (SPEC.md section 2); the viewer is a no-op.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

extern "C" {
#include "pcs_oracle.h"
}

namespace Eigen {
struct Matrix4f {
    float m[16];  // row-major
    struct CommaInit {
        Matrix4f *self; int k;
        CommaInit operator,(double v) { self->m[k] = (float)v; return CommaInit{self, k + 1}; }
    };
    CommaInit operator<<(double v) { m[0] = (float)v; return CommaInit{this, 1}; }
};
template <class T> using aligned_allocator = std::allocator<T>;
}  // namespace Eigen

namespace pcl {

struct PointXYZ {
    float x = 0, y = 0, z = 0, w = 1.0f;
};

struct PointXYZRGB {  // 32 bytes, same layout as pcs_oracle_pclpoint
    float x = 0, y = 0, z = 0, w = 1.0f;
    uint8_t b = 0, g = 0, r = 0, a = 255;
    uint32_t pad[3] = {0, 0, 0};
};
static_assert(sizeof(PointXYZRGB) == sizeof(pcs_oracle_pclpoint), "layout");

template <class PointT> class PointCloud {
public:
    typedef std::shared_ptr<PointCloud<PointT>> Ptr;
    std::vector<PointT> points;
    uint32_t width = 0, height = 0;
    bool is_dense = true;
    void clear() { points.clear(); width = 0; height = 0; }
    PointCloud &operator+=(const PointCloud &rhs) {
        points.insert(points.end(), rhs.points.begin(), rhs.points.end());
        width = (uint32_t)points.size();
        height = 1;
        is_dense = is_dense && rhs.is_dense;
        return *this;
    }
};

inline void transformPointCloud(const PointCloud<PointXYZRGB> &in, PointCloud<PointXYZRGB> &out,
                                const Eigen::Matrix4f &t) {
    if (&in != &out) out = in;
    pcs_oracle_transform_cloud(reinterpret_cast<pcs_oracle_pclpoint *>(out.points.data()),
                               (int)out.points.size(), t.m);
}

namespace visualization {
enum RenderingProperties { PCL_VISUALIZER_POINT_SIZE = 0 };
template <class PointT> struct PointCloudColorHandlerRGBField {
    explicit PointCloudColorHandlerRGBField(const typename PointCloud<PointT>::Ptr &) {}
};
class PCLVisualizer {
public:
    explicit PCLVisualizer(const std::string &) {}
    void setBackgroundColor(double, double, double, int = 0) {}
    template <class C, class H> void addPointCloud(const C &, const H &, const std::string &) {}
    void setPointCloudRenderingProperties(int, double, const std::string &) {}
    template <class C> void updatePointCloud(const C &, const std::string &) {}
    void spinOnce() {}
    bool wasStopped() const { return true; }
};
}  // namespace visualization

namespace io {
template <class C> int savePLYFileBinary(const std::string &, const C &) { return 0; }
}
}  // namespace pcl
