/* TEST INFRASTRUCTURE ONLY -- see pcs_oracle.h for who may load this.
 *
 * Plain-C restatement of the reference hot path.  Build (oracle/Makefile):
 *   gcc -O2 -fPIC -shared -fopenmp -mfma -ffp-contract=off pcs_oracle.c
 * -ffp-contract=off so that a*b+c is two roundings unless written as fmaf();
 * -mfma so that fmaf() is the single-rounding hardware instruction the
 * reference's _mm_fmadd_ps uses.
 */
#include "pcs_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* x86 CVTTSS2SI / CVTTPS2DQ: truncate toward zero; NaN and out-of-range give the
 * "integer indefinite" value 0x80000000 (used by _mm_cvttps_epi32 at
 * src/pcs-camera-optimized.cpp:438-439 and by gcc for short(float) at :581-583). */
static int32_t x86_cvtt(float f)
{
    if (!(f < 2147483648.0f) || f < -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* ---- SPEC.md s1: rs2::pointcloud::calculate (third party; call sites
 * src/pcs-camera-optimized.cpp:198-199,288-289).  Restates librealsense2's
 * rsutil.h rs2_deproject_pixel_to_point / rs2_transform_point_to_point /
 * rs2_project_point_to_pixel with zero distortion, and proc/pointcloud.cpp's
 * tex = pixel / (width, height), z == 0 -> tex (0,0).  PARITY UNPINNED. */
/* librealsense2 rsutil.h (2.16): the radial factor and the two Brown-Conrady forms, written as that header writes
 * them (this file is compiled with -ffp-contract=off, so every operator below rounds once, left to right). */
static float bc_radial(const float *k, float r2)
{
    return 1 + k[0] * r2 + k[1] * r2 * r2 + k[4] * r2 * r2 * r2;
}
/* rs2_project_point_to_pixel, RS2_DISTORTION_MODIFIED_BROWN_CONRADY */
static void bc_project(const float *k, float *x, float *y)
{
    const float r2 = *x * *x + *y * *y;
    const float f = bc_radial(k, r2);
    *x *= f;
    *y *= f;
    {
        const float dx = *x + 2 * k[2] * *x * *y + k[3] * (r2 + 2 * *x * *x);
        const float dy = *y + 2 * k[3] * *x * *y + k[2] * (r2 + 2 * *y * *y);
        *x = dx;
        *y = dy;
    }
}
/* rs2_deproject_pixel_to_point, RS2_DISTORTION_INVERSE_BROWN_CONRADY */
static void bc_deproject(const float *k, float *x, float *y)
{
    const float r2 = *x * *x + *y * *y;
    const float f = bc_radial(k, r2);
    const float ux = *x * f + 2 * k[2] * *x * *y + k[3] * (r2 + 2 * *x * *x);
    const float uy = *y * f + 2 * k[3] * *x * *y + k[2] * (r2 + 2 * *y * *y);
    *x = ux;
    *y = uy;
}

void pcs_oracle_deproject(const pcs_oracle_calib *c, const uint16_t *z16, float *xyz, float *uv,
                          int num_threads)
{
    const int W = c->depth.width, H = c->depth.height;
    const float *R = c->rotation, *T = c->translation;
    const float cw = (float)c->color.width, chh = (float)c->color.height;
    if (num_threads < 1) num_threads = 1;
#pragma omp parallel for schedule(static) num_threads(num_threads)
    for (int y = 0; y < H; ++y) {
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            const float depth = c->depth_scale * (float)z16[i];
            float nx = ((float)x - c->depth.ppx) / c->depth.fx;
            float ny = ((float)y - c->depth.ppy) / c->depth.fy;
            if (c->depth.model == 2) bc_deproject(c->depth.coeffs, &nx, &ny);
            const float p0 = depth * nx, p1 = depth * ny, p2 = depth;
            xyz[3 * i + 0] = p0;
            xyz[3 * i + 1] = p1;
            xyz[3 * i + 2] = p2;
            if (p2 != 0.0f) {
                const float t0 = R[0] * p0 + R[3] * p1 + R[6] * p2 + T[0];
                const float t1 = R[1] * p0 + R[4] * p1 + R[7] * p2 + T[1];
                const float t2 = R[2] * p0 + R[5] * p1 + R[8] * p2 + T[2];
                float qx = t0 / t2, qy = t1 / t2;
                if (c->color.model == 1) bc_project(c->color.coeffs, &qx, &qy);
                const float px = qx * c->color.fx + c->color.ppx;
                const float py = qy * c->color.fy + c->color.ppy;
                uv[2 * i + 0] = px / cw;
                uv[2 * i + 1] = py / chh;
            } else {
                uv[2 * i + 0] = 0.0f;
                uv[2 * i + 1] = 0.0f;
            }
        }
    }
}

/* One point of the SIMD loop: tex lookup (:431-452) + affine (:471-491). */
static void pack_one(const float *p, const float *t, const uint8_t *color, int cw, int ch, int bpp,
                     int stride, const float *tf, int16_t *rec)
{
    /* :434-444  fma(u, w, .5) -> cvtt -> max 0 -> min w-1 */
    int xi = x86_cvtt(fmaf(t[0], (float)cw, 0.5f));
    int yi = x86_cvtt(fmaf(t[1], (float)ch, 0.5f));
    xi = imin(imax(xi, 0), cw - 1);
    yi = imin(imax(yi, 0), ch - 1);
    const int idx = xi * bpp + yi * stride; /* :449-452 */
    for (int r = 0; r < 3; ++r) {
        /* :471-473  v = fma(z, c, fma(y, b, fma(x, a, d))), then a separate *1000.0f (:488) */
        float v = fmaf(p[0], tf[4 * r + 0], tf[4 * r + 3]);
        v = fmaf(p[1], tf[4 * r + 1], v);
        v = fmaf(p[2], tf[4 * r + 2], v);
        v = v * 1000.0f;
        rec[r] = (int16_t)x86_cvtt(v); /* :581-583 */
    }
    rec[3] = (int16_t)(color[idx] + (color[idx + 1] << 8)); /* :584 */
    rec[4] = (int16_t)color[idx + 2];                       /* :585 */
}

/* :499-511, on the PRE-transform vertex. */
static int cutoff_keep(const float *p)
{
    return p[2] > 0.0f && p[2] <= 1.5f && p[0] > -2.0f && p[0] <= 2.0f;
}

int pcs_oracle_pack_simd(const float *xyz, const float *uv, int n, const uint8_t *color, int cw,
                         int ch, int bpp, int stride, const float *tf, int cutoff, int16_t *out)
{
    if (n < 0 || (n & 3)) return -1;
    if (!cutoff) {
        for (int i = 0; i < n; ++i)
            pack_one(xyz + 3 * i, uv + 2 * i, color, cw, ch, bpp, stride, tf, out + 5 * (size_t)i);
        return n; /* :615 */
    }
    /* :501-577.  _mm_set_ps(vert[i].z, .., vert[i+3].z) puts point i in lane 3, but
     * pt_mask_f[0] (lane 0 = point i+3) gates point i: the mask is lane-reversed
     * within each group of four.  -t 1 order = raster order. */
    int count = 0;
    for (int i = 0; i < n; i += 4) {
        for (int k = 0; k < 4; ++k) {
            if (cutoff_keep(xyz + 3 * (i + 3 - k))) {
                pack_one(xyz + 3 * (i + k), uv + 2 * (i + k), color, cw, ch, bpp, stride, tf,
                         out + 5 * (size_t)count);
                ++count;
            }
        }
    }
    return count; /* :612-613 */
}

void pcs_oracle_transform_points(const float *xyz, int n, const float *tf, float *xyz_out)
{
    for (int i = 0; i < n; ++i) {
        const float *p = xyz + 3 * (size_t)i;
        for (int r = 0; r < 3; ++r) {
            float v = fmaf(p[0], tf[4 * r + 0], tf[4 * r + 3]);
            v = fmaf(p[1], tf[4 * r + 1], v);
            v = fmaf(p[2], tf[4 * r + 2], v);
            xyz_out[3 * (size_t)i + r] = v;
        }
    }
}

int pcs_oracle_send(const float *xyz, const float *uv, int n, const uint8_t *color, int cw, int ch,
                    int bpp, int stride, const float *tf, int cutoff, int write_header,
                    int16_t *buffer)
{
    memset(buffer, 0, 5000000); /* :673 -- BUF_SIZE bytes, i.e. half of the short[BUF_SIZE] buffer */
    /* :690 -- &buffer[0] + sizeof(short) on a short* = byte offset 4 */
    int size = pcs_oracle_pack_simd(xyz, uv, n, color, cw, ch, bpp, stride, tf, cutoff, buffer + 2);
    if (size < 0) return size;
    size = 5 * size * (int)sizeof(int16_t); /* :697 */
    if (write_header) memcpy(buffer, &size, sizeof(int)); /* :715-718 */
    return size;
}

int pcs_oracle_concat(const int16_t *const *pc_buf, const int *n_shorts, int n_cams, int downsample,
                      int16_t *stitched_buf)
{
    if (downsample < 1) return -1;
    int stitch_size = 0;
    const int increment = 5 * downsample;      /* src/pcs-multicamera-client.cpp:375 */
    int16_t *pcs_buf = stitched_buf + 2;       /* :378 */
    for (int i = 0; i < n_cams; ++i) {
        for (int j = 0; j < n_shorts[i]; j += increment) { /* :388-391 */
            memcpy(pcs_buf + stitch_size, pc_buf[i] + j, 5 * sizeof(int16_t));
            stitch_size += 5;
        }
    }
    stitch_size *= (int)sizeof(int16_t);       /* :394 */
    memcpy(stitched_buf, &stitch_size, sizeof(int)); /* :395 */
    return stitch_size;
}

int pcs_oracle_unpack(const int16_t *buffer, int size, int downsample, pcs_oracle_pclpoint *out)
{
    if (downsample < 1) return -1;
    /* src/pcs-multicamera-optimized.cpp:230-233: width = size / downsample, resize(width).
     * The reference loop (:235-245) keeps every i % downsample == 0, which is
     * ceil(size/downsample) points -- one past the vector when size % downsample != 0 (UB).
     * The restatement stops at `width`, the cloud's declared size. */
    const int width = size / downsample;
    int count = 0;
    for (int i = 0; i < size && count < width; ++i) {
        if (i % downsample == 0) {
            pcs_oracle_pclpoint q;
            memset(&q, 0, sizeof q);
            q.x = (float)buffer[i * 5 + 0] / 1000.0f; /* :237, CONV_RATE is const float (:46) */
            q.y = (float)buffer[i * 5 + 1] / 1000.0f;
            q.z = (float)buffer[i * 5 + 2] / 1000.0f;
            q.w = 1.0f;                                /* PointXYZRGB() sets data[3] = 1 */
            q.r = (uint8_t)(buffer[i * 5 + 3] & 0xFF); /* :240 */
            q.g = (uint8_t)(buffer[i * 5 + 3] >> 8);   /* :241 */
            q.b = (uint8_t)(buffer[i * 5 + 4] & 0xFF); /* :242 */
            q.a = 255;                                 /* PointXYZRGB() sets a = 255 */
            out[count++] = q;
        }
    }
    return count;
}

/* SPEC.md s2.  pcl::transformPointCloud (third party, PCL 1.8; call site
 * src/pcs-multicamera-optimized.cpp:289) on a non-dense cloud: points with a
 * non-finite coordinate are left untouched; otherwise
 * out = ((m0*x + m1*y) + m2*z) + m3 per row, fp32, no contraction.  PARITY UNPINNED. */
void pcs_oracle_transform_cloud(pcs_oracle_pclpoint *pts, int n, const float *m)
{
    for (int i = 0; i < n; ++i) {
        const float x = pts[i].x, y = pts[i].y, z = pts[i].z;
        if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;
        pts[i].x = m[0] * x + m[1] * y + m[2] * z + m[3];
        pts[i].y = m[4] * x + m[5] * y + m[6] * z + m[7];
        pts[i].z = m[8] * x + m[9] * y + m[10] * z + m[11];
    }
}

int pcs_oracle_repack(const pcs_oracle_pclpoint *pts, int n, int16_t *buffer)
{
    int size = 0;
    for (int i = 0; i < n; ++i) { /* src/pcs-multicamera-optimized.cpp:254-262 */
        buffer[size * 5 + 0] = (int16_t)x86_cvtt(pts[i].x * 1000.0f);
        buffer[size * 5 + 1] = (int16_t)x86_cvtt(pts[i].y * 1000.0f);
        buffer[size * 5 + 2] = (int16_t)x86_cvtt(pts[i].z * 1000.0f);
        buffer[size * 5 + 3] = (int16_t)((int16_t)pts[i].r + (int16_t)(pts[i].g << 8));
        buffer[size * 5 + 4] = (int16_t)pts[i].b;
        ++size;
    }
    return size;
}

/* ---- SPEC.md s3: voxel merge (own spec; the reference only #includes
 * pcl/filters/voxel_grid.h, src/pcs-multicamera-optimized.cpp:17). */
static int floordiv(int a, int b) { int q = a / b; return (a % b != 0 && (a < 0)) ? q - 1 : q; }

typedef struct { uint64_t key; int32_t idx; } vox_item;

static int vox_cmp(const void *a, const void *b)
{
    const vox_item *p = (const vox_item *)a, *q = (const vox_item *)b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return (p->idx > q->idx) - (p->idx < q->idx);
}

int pcs_oracle_voxel_merge(const int16_t *records, int n, int leaf_mm, int16_t *out)
{
    if (leaf_mm < 1 || n < 0) return -1;
    if (n == 0) return 0;
    vox_item *it = (vox_item *)malloc((size_t)n * sizeof *it);
    if (!it) return -2;
    for (int i = 0; i < n; ++i) {
        const int16_t *r = records + 5 * (size_t)i;
        const uint64_t kx = (uint64_t)(floordiv(r[0], leaf_mm) + 32768);
        const uint64_t ky = (uint64_t)(floordiv(r[1], leaf_mm) + 32768);
        const uint64_t kz = (uint64_t)(floordiv(r[2], leaf_mm) + 32768);
        it[i].key = (kz << 34) | (ky << 17) | kx; /* ascending (kz, ky, kx) */
        it[i].idx = i;
    }
    qsort(it, (size_t)n, sizeof *it, vox_cmp);
    int nv = 0;
    for (int i = 0; i < n;) {
        int j = i;
        /* 64-bit sums: exact integer means for any n (SPEC.md s3 said uint32 when n * 255 < 2^32 was an ABI limit) */
        uint64_t sx = 0, sy = 0, sz = 0, sr = 0, sg = 0, sb = 0, cnt = 0;
        const int kx = (int)(it[i].key & 0x1FFFF) - 32768;
        const int ky = (int)((it[i].key >> 17) & 0x1FFFF) - 32768;
        const int kz = (int)((it[i].key >> 34) & 0x1FFFF) - 32768;
        for (; j < n && it[j].key == it[i].key; ++j) {
            const int16_t *r = records + 5 * (size_t)it[j].idx;
            sx += (uint64_t)(r[0] - leaf_mm * kx);
            sy += (uint64_t)(r[1] - leaf_mm * ky);
            sz += (uint64_t)(r[2] - leaf_mm * kz);
            sr += (uint64_t)((uint16_t)r[3] & 0xFF);
            sg += (uint64_t)((uint16_t)r[3] >> 8);
            sb += (uint64_t)((uint16_t)r[4] & 0xFF);
            ++cnt;
        }
        int16_t *o = out + 5 * (size_t)nv;
        o[0] = (int16_t)(leaf_mm * kx + (int)(sx / cnt));
        o[1] = (int16_t)(leaf_mm * ky + (int)(sy / cnt));
        o[2] = (int16_t)(leaf_mm * kz + (int)(sz / cnt));
        o[3] = (int16_t)((sr / cnt) | ((sg / cnt) << 8));
        o[4] = (int16_t)(sb / cnt);
        ++nv;
        i = j;
    }
    free(it);
    return nv;
}
