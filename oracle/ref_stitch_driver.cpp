// TEST INFRASTRUCTURE ONLY.  Linked with ONE of the reference's stitcher TUs,
// compiled UNMODIFIED from /root/reference with -Dmain=ref_main (oracle/Makefile):
//   -DPCS_REF_CLIENT    src/pcs-multicamera-client.cpp    -> _ref/libpcs_ref_client.so
//   (default)           src/pcs-multicamera-optimized.cpp -> _ref/libpcs_ref_optimized.so
// The reference functions talk to sockets; the driver feeds them through
// socketpairs, one camera only (NUM_CAMERAS is `const int 1` in both TUs,
// SURVEY F8).  Not used by the product.
#include <librealsense2/rs.hpp>
#include <pcl/point_cloud.h>

#include <sys/socket.h>
#include <unistd.h>

#include <cstring>
#include <thread>
#include <vector>

typedef pcl::PointCloud<pcl::PointXYZRGB> pointCloudXYZRGB;

// ---- symbols defined by the reference TU -----------------------------------
extern int downsample, client_sockfd;
extern int sockfd_array[];
extern short *stitched_buf;
extern Eigen::Matrix4f transform[];
pointCloudXYZRGB::Ptr convertBufferToPointCloudXYZRGB(short *buffer, int size);
int convertPointCloudXYZRGBToBuffer(pointCloudXYZRGB::Ptr cloud, short *buffer);
void updateCloudXYZRGB(int thread_num, int sockfd, pointCloudXYZRGB::Ptr cloud);
void send_stitchedXYZRGB(pointCloudXYZRGB::Ptr stitched_cloud);
#ifdef PCS_REF_CLIENT
extern short *pc_buf[];
void sendStitchToUnity();
#endif

namespace {

// camera side of the wire: send [i32 bytes][payload], then swallow the 'Z' pull
struct fake_camera {
    int sv[2];
    std::thread th;
    fake_camera(const int16_t *payload, int n_shorts) {
        socketpair(AF_UNIX, SOCK_STREAM, 0, sv);
        th = std::thread([this, payload, n_shorts] {
            int bytes = n_shorts * (int)sizeof(int16_t);
            const char *p = reinterpret_cast<const char *>(&bytes);
            for (int off = 0; off < 4;) off += (int)write(sv[1], p + off, 4 - off);
            p = reinterpret_cast<const char *>(payload);
            for (int off = 0; off < bytes;) {
                ssize_t w = write(sv[1], p + off, (size_t)(bytes - off));
                if (w <= 0) break;
                off += (int)w;
            }
            char z;
            (void)!read(sv[1], &z, 1);
        });
    }
    ~fake_camera() { th.join(); close(sv[0]); close(sv[1]); }
};

// viewer side: send the 'Z' pull, read [i32 bytes][payload]
struct fake_viewer {
    int sv[2];
    std::thread th;
    std::vector<uint8_t> got;
    fake_viewer() {
        socketpair(AF_UNIX, SOCK_STREAM, 0, sv);
        th = std::thread([this] {
            char z = 'Z';
            (void)!write(sv[1], &z, 1);
            int bytes = 0;
            uint8_t *h = reinterpret_cast<uint8_t *>(&bytes);
            for (int off = 0; off < 4;) {
                ssize_t r = read(sv[1], h + off, 4 - off);
                if (r <= 0) return;
                off += (int)r;
            }
            got.resize((size_t)bytes + 4);
            std::memcpy(got.data(), &bytes, 4);
            for (int off = 0; off < bytes;) {
                ssize_t r = read(sv[1], got.data() + 4 + off, (size_t)(bytes - off));
                if (r <= 0) return;
                off += (int)r;
            }
        });
    }
    int finish(uint8_t *out, int cap) {
        th.join();
        close(sv[0]);
        close(sv[1]);
        if ((int)got.size() > cap) return -1;
        std::memcpy(out, got.data(), got.size());
        return (int)got.size();
    }
};

void ensure_stitched_buf() {
    if (!stitched_buf) stitched_buf = (short *)malloc(sizeof(short) * 32000000);  // :407 / :499
}

}  // namespace

extern "C" {

// convertBufferToPointCloudXYZRGB; size % ds must be 0 (the reference writes one
// element past its vector otherwise).  Returns the cloud width.
int ref_unpack(const int16_t *buffer, int size, int ds, pcs_oracle_pclpoint *out) {
    if (ds < 1 || size % ds) return -1;
    downsample = ds;
    pointCloudXYZRGB::Ptr c = convertBufferToPointCloudXYZRGB(const_cast<short *>(buffer), size);
    std::memcpy(out, c->points.data(), c->points.size() * sizeof(pcs_oracle_pclpoint));
    downsample = 1;
    return (int)c->width;
}

int ref_repack(const pcs_oracle_pclpoint *pts, int n, int16_t *out) {
    pointCloudXYZRGB::Ptr c(new pointCloudXYZRGB);
    c->points.resize((size_t)n);
    std::memcpy(c->points.data(), pts, (size_t)n * sizeof(pcs_oracle_pclpoint));
    c->width = (uint32_t)n;
    c->height = 1;
    return convertPointCloudXYZRGBToBuffer(c, out);
}

// One frame of the PCL path for one camera (runStitching body,
// src/pcs-multicamera-optimized.cpp:354-382): updateCloudXYZRGB -> += ->
// send_stitchedXYZRGB.  Returns the bytes the viewer received, copied to out.
int ref_pcl_stitch_1cam(const int16_t *payload, int n_shorts, int ds, const float *tf16,
                        uint8_t *out, int cap) {
    if (ds < 1 || (n_shorts / 5) % ds) return -1;
    ensure_stitched_buf();
    downsample = ds;
    std::memcpy(transform[0].data(), tf16, 16 * sizeof(float));
#ifdef PCS_REF_CLIENT
    if (!pc_buf[0]) pc_buf[0] = (short *)malloc(sizeof(short) * 5000000);  // :554
#endif
    pointCloudXYZRGB::Ptr cloud(new pointCloudXYZRGB), stitched(new pointCloudXYZRGB);
    {
        fake_camera cam(payload, n_shorts);
        updateCloudXYZRGB(0, cam.sv[0], cloud);
    }
    *stitched += *cloud;
    fake_viewer v;
    client_sockfd = v.sv[0];
    send_stitchedXYZRGB(stitched);
    downsample = 1;
    return v.finish(out, cap);
}

#ifdef PCS_REF_CLIENT
// One frame of the raw path (src/pcs-multicamera-client.cpp:373-409), one camera.
int ref_raw_stitch_1cam(const int16_t *payload, int n_shorts, int ds, uint8_t *out, int cap) {
    if (ds < 1) return -1;
    ensure_stitched_buf();
    if (!pc_buf[0]) pc_buf[0] = (short *)malloc(sizeof(short) * 5000000);
    downsample = ds;
    fake_camera cam(payload, n_shorts);
    sockfd_array[0] = cam.sv[0];
    fake_viewer v;
    client_sockfd = v.sv[0];
    sendStitchToUnity();
    downsample = 1;
    return v.finish(out, cap);
}
#endif

}  // extern "C"
