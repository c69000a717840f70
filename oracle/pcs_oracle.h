/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product
 * (pointcloud_stitching_b200/) never does and has no CPU fallback.
 *
 * Every function cites the reference file:line (relative to /root/reference)
 * whose arithmetic it restates.  Parity status:
 *   - pack / send / concat / unpack / repack: PINNED against the reference's
 *     own compiled functions (oracle/_ref, built by oracle/Makefile from the
 *     sources under /root/reference) and against tests/golden/ fixtures that
 *     were generated from them (tests/golden/make_golden.py).
 *   - deproject (librealsense2 rs2::pointcloud::calculate), PCL
 *     transformPointCloud / operator+=, voxel merge: PARITY UNPINNED -- the
 *     arithmetic lives in third-party libraries that are not under
 *     /root/reference (or, for the voxel grid, is never called by the
 *     reference at all); oracle/SPEC.md is the specification.
 */
#ifndef PCS_ORACLE_H
#define PCS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* model / coeffs: rs2_distortion (0 none, 1 modified Brown-Conrady: applied when projecting into this sensor,
 * 2 inverse Brown-Conrady: applied when deprojecting from it); coeffs = k1, k2, p1, p2, k3. */
typedef struct pcs_oracle_intrinsics {
    int width, height;
    float ppx, ppy, fx, fy;
    int model;
    float coeffs[5];
} pcs_oracle_intrinsics;

/* depth -> colour calibration, rs2_intrinsics / rs2_extrinsics conventions:
 * rotation is column-major 3x3, translation in metres. */
typedef struct pcs_oracle_calib {
    pcs_oracle_intrinsics depth, color;
    float rotation[9];
    float translation[3];
    float depth_scale;
} pcs_oracle_calib;

/* pcl::PointXYZRGB memory layout (32 bytes): data[4] = {x,y,z,1.0f}; then b,g,r,a; 12 B pad. */
typedef struct pcs_oracle_pclpoint {
    float x, y, z, w;
    uint8_t b, g, r, a;
    uint32_t pad[3];
} pcs_oracle_pclpoint;

/* SPEC.md s1: z16 -> vertex[N] (xyz, 3 floats/pt) + texture_coordinate[N] (uv). */
void pcs_oracle_deproject(const pcs_oracle_calib *c, const uint16_t *z16, float *xyz, float *uv,
                          int num_threads);

/* src/pcs-camera-optimized.cpp:363-616 (SIMD path, -m).  n must be a multiple of 4.
 * tf = 16 floats row-major.  cutoff != 0 reproduces -c at -t 1 (raster order,
 * lane-reversed mask).  Returns the number of records written, <0 on bad args. */
int pcs_oracle_pack_simd(const float *xyz, const float *uv, int n, const uint8_t *color, int cw,
                         int ch, int bpp, int stride, const float *tf, int cutoff, int16_t *out);

/* Same loop, also returning the float xyz' (metres, before the *1000) for the 1e-5 check. */
void pcs_oracle_transform_points(const float *xyz, int n, const float *tf, float *xyz_out);

/* src/pcs-camera-optimized.cpp:669-723: memset 5 000 000 B, pack at byte 4, optional header. */
int pcs_oracle_send(const float *xyz, const float *uv, int n, const uint8_t *color, int cw, int ch,
                    int bpp, int stride, const float *tf, int cutoff, int write_header,
                    int16_t *buffer);

/* src/pcs-multicamera-client.cpp:373-395: ordered concat with stride decimation.
 * n_shorts[i] = payload length of camera i in shorts.  Returns stitched payload bytes. */
int pcs_oracle_concat(const int16_t *const *pc_buf, const int *n_shorts, int n_cams, int downsample,
                      int16_t *stitched_buf);

/* src/pcs-multicamera-optimized.cpp:226-248.  Writes size/downsample points; returns that count. */
int pcs_oracle_unpack(const int16_t *buffer, int size, int downsample, pcs_oracle_pclpoint *out);

/* SPEC.md s2 (pcl::transformPointCloud restated), in place; m = 16 floats row-major. */
void pcs_oracle_transform_cloud(pcs_oracle_pclpoint *pts, int n, const float *m);

/* src/pcs-multicamera-optimized.cpp:251-265. */
int pcs_oracle_repack(const pcs_oracle_pclpoint *pts, int n, int16_t *buffer);

/* SPEC.md s3: integer voxel merge of 10-byte records; returns voxel count. */
int pcs_oracle_voxel_merge(const int16_t *records, int n, int leaf_mm, int16_t *out);

#ifdef __cplusplus
}
#endif
#endif
