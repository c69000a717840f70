// CPU check of the guard behind the pipelined kernel's colour taps under a rotated depth->colour calibration
// (pointcloud_stitching_b200/csrc/pcs_guard.h): the bound eps that pipe_guard() derives for a calibration must
// dominate the distance between
//   * the exact chain (oracle/SPEC.md s1, every operator rounded once: what the reference's call into librealsense
//     computes, src/pcs-camera-optimized.cpp:198-199 with :434-444 for the index), argument of trunc(), and
//   * the cheap chain the kernel evaluates (three FMAs per component, an approximate reciprocal, one FMA to pixels),
// for every pixel (the bound is linear in the pixel's normalised source coordinate) and every depth >= the guard depth.  Sampled here on random rigs (rotations up to ~1.5 degrees,
// translations up to 6 cm, depth and colour sensors of different sizes) with the reciprocal perturbed by up to one
// unit in the last place in either direction (PTX rcp.approx.ftz.f32: at most 1 ulp).  Prints the largest observed
// fraction of the bound (must stay below 1 / PIPE_GUARD_SAFETY) and exits 0 when it does.
// Build: g++ -O2 -ffp-contract=off -fopenmp -I pointcloud_stitching_b200/csrc tests/cpp/guard_check.cpp
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "pcs_guard.h"

struct Params {
    int W, H, CW, CH;
    float ppx, ppy, fx, fy, cppx, cppy, cfx, cfy, cwf, chf, depth_scale;
    float R[9], T[3];
};

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static double urand() {
    rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull;
    return (double)(rng_state >> 11) / 9007199254740992.0;
}
static float nudge(float v, int ulps) {
    uint32_t b;
    memcpy(&b, &v, 4);
    b += (uint32_t)ulps;     // positive normal floats: +-1 on the bit pattern = +-1 ulp
    memcpy(&v, &b, 4);
    return v;
}

int main(int argc, char **argv) {
    const int rigs = argc > 1 ? atoi(argv[1]) : 40;
    const long long samples = argc > 2 ? atoll(argv[2]) : 400000;
    double worst = 0;
    int refused = 0;
    for (int rig = 0; rig < rigs; ++rig) {
        Params p;
        static const int sizes[][4] = {{1280, 720, 1280, 720}, {1280, 720, 1920, 1080}, {848, 480, 1280, 720}, {640, 480, 640, 480}};
        const int *sz = sizes[rig % 4];
        p.W = sz[0]; p.H = sz[1]; p.CW = sz[2]; p.CH = sz[3];
        p.fx = p.W * (float)(0.45 + 0.3 * urand()); p.fy = p.fx * (float)(0.98 + 0.04 * urand());
        p.ppx = (float)(p.W * (0.48 + 0.04 * urand())); p.ppy = (float)(p.H * (0.48 + 0.04 * urand()));
        p.cfx = p.CW * (float)(0.45 + 0.3 * urand()); p.cfy = p.cfx * (float)(0.98 + 0.04 * urand());
        p.cppx = (float)(p.CW * (0.48 + 0.04 * urand())); p.cppy = (float)(p.CH * (0.48 + 0.04 * urand()));
        p.cwf = (float)p.CW; p.chf = (float)p.CH;
        p.depth_scale = 0.001f;
        const double rx = 0.025 * (urand() - 0.5), ry = 0.025 * (urand() - 0.5), rz = 0.025 * (urand() - 0.5);
        const double cx = cos(rx), sx = sin(rx), cy = cos(ry), sy = sin(ry), cz = cos(rz), sz_ = sin(rz);
        const double Rm[3][3] = {{cz * cy, cz * sy * sx - sz_ * cx, cz * sy * cx + sz_ * sx},
                                 {sz_ * cy, sz_ * sy * sx + cz * cx, sz_ * sy * cx - cz * sx},
                                 {-sy, cy * sx, cy * cx}};
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) p.R[3 * c + r] = (float)Rm[r][c];      // column-major
        p.T[0] = (float)(0.06 * (urand() - 0.5)); p.T[1] = (float)(0.01 * (urand() - 0.5)); p.T[2] = (float)(0.01 * (urand() - 0.5));
        const pcs::PipeGuard g = pcs::pipe_guard(p);
        if (!g.ok) { ++refused; continue; }
        double worst_rig = 0;
        for (long long i = 0; i < samples; ++i) {
            const int x = (int)(urand() * p.W), y = (int)(urand() * p.H);
            // depths: a third near the guard depth, where the bound is tightest
            const int z = i % 3 == 0 ? pcs::PIPE_GUARD_Z16 + (int)(urand() * 200) : pcs::PIPE_GUARD_Z16 + (int)(urand() * (65535 - pcs::PIPE_GUARD_Z16));
            const float nx = ((float)x - p.ppx) / p.fx, ny = ((float)y - p.ppy) / p.fy;
            const float d = p.depth_scale * (float)z, p0 = d * nx, p1 = d * ny;
            // exact chain
            const float t0 = p.R[0] * p0 + p.R[3] * p1 + p.R[6] * d + p.T[0];
            const float t1 = p.R[1] * p0 + p.R[4] * p1 + p.R[7] * d + p.T[1];
            const float t2 = p.R[2] * p0 + p.R[5] * p1 + p.R[8] * d + p.T[2];
            const float px = (t0 / t2) * p.cfx + p.cppx, py = (t1 / t2) * p.cfy + p.cppy;
            const float tx = fmaf(px / p.cwf, p.cwf, 0.5f), ty = fmaf(py / p.chf, p.chf, 0.5f);
            // cheap chain
            const float a0 = fmaf(p.R[0], p0, fmaf(p.R[3], p1, fmaf(p.R[6], d, p.T[0])));
            const float a1 = fmaf(p.R[1], p0, fmaf(p.R[4], p1, fmaf(p.R[7], d, p.T[1])));
            const float a2 = fmaf(p.R[2], p0, fmaf(p.R[5], p1, fmaf(p.R[8], d, p.T[2])));
            const float y0 = nudge((float)(1.0 / (double)a2), (int)(i % 3) - 1);
            // (the kernel never adds the 1/2: it rounds fx - 1/2 to nearest where the guard passes)
            const double fx = (double)fmaf(a0 * y0, p.cfx, p.cppx) + 0.5, fy = (double)fmaf(a1 * y0, p.cfy, p.cppy) + 0.5;
            // taps far outside the frame are clamped by both chains
            const double eps_x = g.ax + g.bx * std::fabs(nx), eps_y = g.ay + g.by * std::fabs(ny);
            if (tx > -4 && tx < p.CW + 4) worst_rig = std::max(worst_rig, std::fabs((double)fx - (double)tx) / eps_x);
            if (ty > -4 && ty < p.CH + 4) worst_rig = std::max(worst_rig, std::fabs((double)fy - (double)ty) / eps_y);
        }
        worst = std::max(worst, worst_rig);
        if (rig < 8 || worst_rig > 0.5)
            printf("rig %2d  %dx%d -> %dx%d  eps_x = %.2e + %.2e |nx|, eps_y = %.2e + %.2e |ny| px   largest |cheap - exact| / eps = %.3f\n",
                   rig, p.W, p.H, p.CW, p.CH, g.ax, g.bx, g.ay, g.by, worst_rig);
    }
    // rigs the bound cannot cover must be refused (the kernel then never runs its cheap chain on them): a colour sensor
    // 20 cm IN FRONT of the depth sensor (t2 <= 0 for near points), a 60-degree yaw (t2 changes sign inside the frame), a
    // focal length that makes the bound wider than the guard could usefully be
    {
        Params p;
        p.W = p.CW = 1280; p.H = p.CH = 720;
        p.fx = p.fy = p.cfx = p.cfy = 640.f; p.ppx = p.cppx = 639.5f; p.ppy = p.cppy = 359.5f;
        p.cwf = 1280.f; p.chf = 720.f; p.depth_scale = 0.001f;
        const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        memcpy(p.R, I, sizeof I);
        p.T[0] = 0.015f; p.T[1] = 0.f; p.T[2] = -0.2f;
        if (pcs::pipe_guard(p).ok) { printf("accepted a rig with T.z = -0.2 m\n"); return 1; }
        p.T[2] = 0.f;
        const float c = 0.5f, s_ = 0.8660254f;      // yaw of 60 degrees, column-major
        const float Y[9] = {c, 0, -s_, 0, 1, 0, s_, 0, c};
        memcpy(p.R, Y, sizeof Y);
        if (pcs::pipe_guard(p).ok) { printf("accepted a 60-degree yaw\n"); return 1; }
        memcpy(p.R, I, sizeof I);
        p.cfx = p.cfy = 4.0e5f;
        if (pcs::pipe_guard(p).ok) { printf("accepted a bound wider than 0.05 px\n"); return 1; }
        p.cfx = p.cfy = 640.f;
        if (!pcs::pipe_guard(p).ok) { printf("refused the plain x-baseline rig\n"); return 1; }
    }
    printf("%d rigs (%d refused by pipe_guard), %lld samples each: largest fraction of the bound %.3f (limit %.3f)\n", rigs, refused,
           samples, worst, 1.0 / pcs::PIPE_GUARD_SAFETY);
    return worst < 1.0 / pcs::PIPE_GUARD_SAFETY && refused < rigs ? 0 : 1;
}
