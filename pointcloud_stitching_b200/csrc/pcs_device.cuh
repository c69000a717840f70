// Device-side building blocks shared by every kernel of the hot path.
//
// All floating-point arithmetic is written with explicit round-to-nearest
// intrinsics: the results must be bit-identical to the reference's SSE/FMA3 CPU
// code (src/pcs-camera-optimized.cpp:431-491) and to the deprojection spec
// (oracle/SPEC.md s1), so nvcc must neither contract a*b+c into an FMA where the
// CPU does two roundings nor split an FMA where the CPU fuses.  The file is also
// compiled with -fmad=false as a second line of defence.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcs {

enum TexMode : int {
    TEX_GENERAL = 0,    // full rotation + translation, then projection
    TEX_TRANSLATE = 1,  // rotation is exactly the identity: R*p + T == p + T bit for bit
    TEX_ALIGNED = 2,    // identity extrinsics and equal intrinsics: the tap is the pixel itself
    TEX_TRANSLATE_X = 3 // translation along x only, equal vertical intrinsics: the tap ROW is the pixel's own
};

// Per-stream constants (pcs_stream_desc, pre-digested on the host).
struct StreamParams {
    int W, H, N;            // depth geometry, N = W*H
    int CW, CH, bpp, stride;
    int tex_mode;
    float ppx, ppy, fx, fy;         // depth intrinsics
    float cppx, cppy, cfx, cfy;     // colour intrinsics
    float cwf, chf;                 // float(CW), float(CH)
    float depth_scale;
    float R[9], T[3];               // depth -> colour, column-major R
    float tf[12];                   // rows 0..2 of the camera -> world 4x4
    int cutoff, lane_rev;
    float z_lo, z_hi, x_lo, x_hi;
    // TEX_TRANSLATE_X with a colour frame of another size / other vertical intrinsics: the colour row every depth
    // row taps (device memory, H entries), proven on the host to be what the exact chain yields for every depth
    // value (pcs_abi.cu make_rowmap); NULL when the tap row is the pixel's own row
    const int32_t *rowmap;
    // lens distortion (rs2_distortion: 0 none, 1 modified Brown-Conrady on projection, 2 inverse Brown-Conrady on
    // deprojection); coefficients k1, k2, p1, p2, k3.  Any model forces TEX_GENERAL.
    int dmodel, cmodel;
    float dcoef[5], ccoef[5];
};

// librealsense rsutil.h, evaluated left to right without contraction (oracle/SPEC.md s1):
//   r2 = x*x + y*y;  f = 1 + k1*r2 + k2*r2*r2 + k3*r2*r2*r2
__device__ __forceinline__ float bc_radial(const float *c, float r2) {
    float f = __fadd_rn(1.0f, __fmul_rn(c[0], r2));
    f = __fadd_rn(f, __fmul_rn(__fmul_rn(c[1], r2), r2));
    f = __fadd_rn(f, __fmul_rn(__fmul_rn(__fmul_rn(c[4], r2), r2), r2));
    return f;
}
// rs2_project_point_to_pixel, RS2_DISTORTION_MODIFIED_BROWN_CONRADY:
//   x *= f; y *= f; dx = x + 2*p1*x*y + p2*(r2 + 2*x*x); dy = y + 2*p2*x*y + p1*(r2 + 2*y*y)
__device__ __forceinline__ void bc_project(const float *c, float &x, float &y) {
    const float r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    const float f = bc_radial(c, r2);
    x = __fmul_rn(x, f);
    y = __fmul_rn(y, f);
    const float dx = __fadd_rn(__fadd_rn(x, __fmul_rn(__fmul_rn(__fmul_rn(2.0f, c[2]), x), y)),
                               __fmul_rn(c[3], __fadd_rn(r2, __fmul_rn(__fmul_rn(2.0f, x), x))));
    const float dy = __fadd_rn(__fadd_rn(y, __fmul_rn(__fmul_rn(__fmul_rn(2.0f, c[3]), x), y)),
                               __fmul_rn(c[2], __fadd_rn(r2, __fmul_rn(__fmul_rn(2.0f, y), y))));
    x = dx;
    y = dy;
}
// rs2_deproject_pixel_to_point, RS2_DISTORTION_INVERSE_BROWN_CONRADY:
//   ux = x*f + 2*p1*x*y + p2*(r2 + 2*x*x); uy = y*f + 2*p2*x*y + p1*(r2 + 2*y*y)
__device__ __forceinline__ void bc_deproject(const float *c, float &x, float &y) {
    const float r2 = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    const float f = bc_radial(c, r2);
    const float ux = __fadd_rn(__fadd_rn(__fmul_rn(x, f), __fmul_rn(__fmul_rn(__fmul_rn(2.0f, c[2]), x), y)),
                               __fmul_rn(c[3], __fadd_rn(r2, __fmul_rn(__fmul_rn(2.0f, x), x))));
    const float uy = __fadd_rn(__fadd_rn(__fmul_rn(y, f), __fmul_rn(__fmul_rn(__fmul_rn(2.0f, c[3]), x), y)),
                               __fmul_rn(c[2], __fadd_rn(r2, __fmul_rn(__fmul_rn(2.0f, y), y))));
    x = ux;
    y = uy;
}

// x86 CVTTSS2SI semantics: truncate; NaN / out of range -> 0x80000000.
// (CUDA's cvt.rzi saturates +overflow to INT_MAX and maps NaN to 0 instead.)
__device__ __forceinline__ int x86_cvtt(float f) {
    return (f < 2147483648.0f) ? __float2int_rz(f) : (int)0x80000000;
}

// src/pcs-camera-optimized.cpp:434-444: trunc(fma(u, w, .5)) clamped to [0, w-1].
__device__ __forceinline__ int tex_to_pixel(float u, float wf, int wmax) {
    int xi = x86_cvtt(__fmaf_rn(u, wf, 0.5f));
    return min(max(xi, 0), wmax);
}

// src/pcs-camera-optimized.cpp:471-491,581-583: one output coordinate in int16 mm.
__device__ __forceinline__ float affine_row(const float *tf, int r, float x, float y, float z) {
    float v = __fmaf_rn(x, tf[4 * r + 0], tf[4 * r + 3]);
    v = __fmaf_rn(y, tf[4 * r + 1], v);
    v = __fmaf_rn(z, tf[4 * r + 2], v);
    return v;
}
__device__ __forceinline__ uint32_t to_mm16(float metres) {
    return (uint32_t)x86_cvtt(__fmul_rn(metres, 1000.0f)) & 0xFFFFu;
}

// Three colour bytes at byte offset o of a 4-byte aligned image, as 0x00BBGGRR.
// Two aligned word loads instead of three byte loads; the second word is only
// touched when it holds one of the three bytes, so nothing past the last valid
// word of the image is read.
__device__ __forceinline__ uint32_t load_rgb(const uint8_t *__restrict__ base, int o) {
    const uint32_t *w = reinterpret_cast<const uint32_t *>(base + (o & ~3));
    const int sh = o & 3;
    uint32_t lo = __ldg(w);
    uint32_t hi = (sh >= 2) ? __ldg(w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh * 8) & 0x00FFFFFFu;
}

// oracle/SPEC.md s1 for one pixel: vertex p and the colour tap (xi, yi).
// nx, ny are ((x - ppx) / fx, (y - ppy) / fy), computed by the caller (they are
// shared along rows / columns).
// S: StreamParams, or any struct with its calibration members (the pipelined kernel passes its parameter-bank copy).
template <int MODE, class S>
__device__ __forceinline__ void deproject_tap(const S &s, uint32_t z16, int x, int y,
                                              float nx, float ny, float &p0, float &p1, float &p2,
                                              int &xi, int &yi) {
    const float depth = __fmul_rn(s.depth_scale, (float)z16);
    if (MODE == TEX_GENERAL && s.dmodel == 2) bc_deproject(s.dcoef, nx, ny);
    p0 = __fmul_rn(depth, nx);
    p1 = __fmul_rn(depth, ny);
    p2 = depth;
    if (MODE == TEX_ALIGNED) {
        // SPEC.md s1: the chain's rounding error is < 1e-3 px, so a valid pixel taps itself
        const bool valid = depth != 0.0f;
        xi = valid ? x : 0;
        yi = valid ? y : 0;
        return;
    }
    if (MODE == TEX_TRANSLATE_X) {
        // t1 == p1 and t2 == p2 exactly, so the row follows the TEX_ALIGNED argument; only the
        // column needs the projection chain
        float u = 0.0f;
        const bool valid = depth != 0.0f;
        if (valid) {
            const float t0 = __fadd_rn(p0, s.T[0]);
            const float px = __fadd_rn(__fmul_rn(__fdiv_rn(t0, p2), s.cfx), s.cppx);
            u = __fdiv_rn(px, s.cwf);
        }
        xi = tex_to_pixel(u, s.cwf, s.CW - 1);
        yi = valid ? (s.rowmap ? __ldg(s.rowmap + y) : y) : 0;
        return;
    }
    float u = 0.0f, v = 0.0f;
    if (depth != 0.0f) {
        float t0, t1, t2;
        if (MODE == TEX_TRANSLATE) {
            t0 = __fadd_rn(p0, s.T[0]);
            t1 = __fadd_rn(p1, s.T[1]);
            t2 = __fadd_rn(p2, s.T[2]);
        } else {
            t0 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s.R[0], p0), __fmul_rn(s.R[3], p1)),
                                     __fmul_rn(s.R[6], p2)), s.T[0]);
            t1 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s.R[1], p0), __fmul_rn(s.R[4], p1)),
                                     __fmul_rn(s.R[7], p2)), s.T[1]);
            t2 = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(s.R[2], p0), __fmul_rn(s.R[5], p1)),
                                     __fmul_rn(s.R[8], p2)), s.T[2]);
        }
        float qx = __fdiv_rn(t0, t2), qy = __fdiv_rn(t1, t2);
        if (MODE == TEX_GENERAL && s.cmodel == 1) bc_project(s.ccoef, qx, qy);
        const float px = __fadd_rn(__fmul_rn(qx, s.cfx), s.cppx);
        const float py = __fadd_rn(__fmul_rn(qy, s.cfy), s.cppy);
        u = __fdiv_rn(px, s.cwf);
        v = __fdiv_rn(py, s.chf);
    }
    xi = tex_to_pixel(u, s.cwf, s.CW - 1);
    yi = tex_to_pixel(v, s.chf, s.CH - 1);
}

// The vertex's x and z alone (what the -c test looks at), exactly as deproject_tap computes them.
template <int MODE, class S>
__device__ __forceinline__ void deproject_xz(const S &s, uint32_t z16, float nx, float ny, float &p0, float &p2) {
    const float depth = __fmul_rn(s.depth_scale, (float)z16);
    if (MODE == TEX_GENERAL && s.dmodel == 2) bc_deproject(s.dcoef, nx, ny);
    p0 = __fmul_rn(depth, nx);
    p2 = depth;
}

// src/pcs-camera-optimized.cpp:499-511 on the pre-transform vertex.
__device__ __forceinline__ bool cutoff_keep(const StreamParams &s, float x, float z) {
    return z > s.z_lo && z <= s.z_hi && x > s.x_lo && x <= s.x_hi;
}

// A 10-byte record as (xy, z|rg, b): the two words and the trailing half-word.
struct Rec {
    uint32_t a, b, c;
};
__device__ __forceinline__ Rec make_record(const float *tf, float p0, float p1, float p2,
                                           uint32_t rgb) {
    Rec r;
    const uint32_t x = to_mm16(affine_row(tf, 0, p0, p1, p2));
    const uint32_t y = to_mm16(affine_row(tf, 1, p0, p1, p2));
    const uint32_t z = to_mm16(affine_row(tf, 2, p0, p1, p2));
    r.a = x | (y << 16);
    r.b = z | ((rgb & 0xFFFFu) << 16);   // R + (G << 8)   (:584)
    r.c = (rgb >> 16) & 0xFFu;           // B, high byte 0 (:585)
    return r;
}

// Two consecutive records (20 bytes) as five little-endian words.
__device__ __forceinline__ void pack_pair(const Rec &r0, const Rec &r1, uint32_t *w) {
    w[0] = r0.a;
    w[1] = r0.b;
    w[2] = r0.c | (r1.a << 16);
    w[3] = __funnelshift_r(r1.a, r1.b, 16);
    w[4] = (r1.b >> 16) | (r1.c << 16);
}

__device__ __forceinline__ void st_global_v4(void *p, uint4 v) {
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x),
                 "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_global_nc_v4(const void *p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

}  // namespace pcs
