// Host-side analysis behind the guarded colour taps of the pipelined kernel (pcs_k1_pipe.cuh, windowed modes).
// Plain C++ (no CUDA): also compiled into tests/cpp/guard_check.cpp.
//
// The tap of a valid pixel is, by oracle/SPEC.md s1 (the reference's call into librealsense,
// src/pcs-camera-optimized.cpp:198-199,288-289, and its own index arithmetic :434-444),
//     xi = trunc(fma(u, CW, .5)) clamped,  u = fl(px / CW),  px = fl(fl(q * cfx) + cppx),  q = fl(t0 / t2),
//     t = R p + T evaluated left to right, one rounding per operator
// -- fifteen roundings whose only product is an INTEGER.  The kernel therefore evaluates a cheaper chain
//     t_a = fma(R.0, p0, fma(R.1, p1, fma(R.2, p2, T)));  fx_a = fma(t0_a * rcp.approx(t2_a), cfx, cppx) + 1/2
// (the 1/2 is never added: the kernel takes rint(fx_a - 1/2), which equals trunc(fx_a) away from the integers, out of the
// guard's own magic-number add) and accepts it only where fx_a keeps a distance eps from every integer, with eps >= E1 + E2:
//     E1 = |exact float chain - value in real arithmetic|,  E2 = |cheap chain - value in real arithmetic|
// (both chains start from the SAME floats p0, p1, p2).  Then the exact chain's argument of trunc lies in the same unit
// interval and the integers agree.  Pixels that fail the test, pixels nearer than the guard depth and taps outside the
// staged window are re-evaluated with the exact chain (deproject_tap<TEX_GENERAL>) -- the guard only decides WHO pays.
//
// Error model (u = 2^-24; gamma_k = k u / (1 - k u); standard results for recursive summation and FMA chains), for the x
// coordinate of a pixel whose normalised source coordinate is nx (y alike, with the second row of R):
//     |t_f - t*| <= gamma_4 S,  |t_a - t*| <= gamma_3 S,   S = |R.0 p0| + |R.1 p1| + |R.2 p2| + |T|
//     t2 >= k' * depth for depths >= z_g,  k' = R22 - |R20| nx_max - |R21| ny_max - max(0, -T2) / z_g
//     |q| <= S0 / |t2| <= qb(nx) = (|R00| |nx| + |R01| ny_max + |R02| + |T0| / z_g) / k'
//     S2 / |t2| <= s2 = (|R20| nx_max + |R21| ny_max + |R22| + |T2| / z_g) / k'
//     rcp.approx: relative error <= 2^-22 (PTX ISA: at most 1 ulp; doubled here)
//     E1 <= qb f (gamma_4 (1 + s2) / (1 - gamma_4 s2) + 5 u) + 4 u |pp| + u            (divide, scale, add, /W, fma .5)
//     E2 <= qb f ((gamma_3 (1 + s2) + 2^-22) / (1 - gamma_3 s2 - 2^-22) + 3 u) + 2 u (|pp| + .5)
// (times 1.01 for the second-order terms).  The bound is worst-case (every rounding error aligned), is multiplied by
// PIPE_GUARD_SAFETY on top, and is LINEAR in |nx|: eps_x(nx) = ax + bx |nx|, so every thread of the kernel carries the
// bound of its own columns and rows instead of the frame's corner (half the pixels at the guard, on average).
// tests/cpp/guard_check.cpp samples both chains on random calibrations and pixels (largest observed |fx_a - tx_f|:
// about 0.4 of the bound); the GPU tests compare every byte with the oracle anyway.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <vector>

namespace pcs {

constexpr double PIPE_GUARD_SAFETY = 1.1;
constexpr int PIPE_GUARD_Z16 = 160;      // z16 below this (and above 0) always takes the exact chain
#ifndef PIPE_SEG_SHIFT
#define PIPE_SEG_SHIFT 7
#endif
constexpr int PIPE_SEG_PX = 1 << PIPE_SEG_SHIFT;   // colour window segment: 128 px = 384 B (a multiple of 16 for the bulk copies)
constexpr int PIPE_SEG_BYTES = PIPE_SEG_PX * 3;
constexpr int PIPE_MAX_SEGS = 2048 / PIPE_SEG_PX;  // colour frames up to 2048 px wide

struct PipeGuard {
    bool ok;
    float ax, bx, ay, by;    // eps_x = ax + bx |nx|, eps_y = ay + by |ny|  (pixels)
};

template <class P>
inline PipeGuard pipe_guard(const P &p) {
    PipeGuard g{false, 0.f, 0.f, 0.f, 0.f};
    const double u = std::ldexp(1.0, -24), g4 = 4 * u / (1 - 4 * u), g3 = 3 * u / (1 - 3 * u), er = std::ldexp(1.0, -22);
    const double nxmax = std::max(std::fabs((0.0 - p.ppx) / p.fx), std::fabs(((double)p.W - p.ppx) / p.fx));
    const double nymax = std::max(std::fabs((0.0 - p.ppy) / p.fy), std::fabs(((double)p.H - p.ppy) / p.fy));
    const double nmax[2] = {nxmax, nymax};
    const double zg = (double)p.depth_scale * PIPE_GUARD_Z16;
    if (!(zg > 0)) return g;
    const double k = p.R[8] - std::fabs(p.R[2]) * nxmax - std::fabs(p.R[5]) * nymax - std::max(0.0, -(double)p.T[2]) / zg;
    if (!(k > 0.04)) return g;
    const double s2 = (std::fabs(p.R[2]) * nxmax + std::fabs(p.R[5]) * nymax + std::fabs(p.R[8]) + std::fabs(p.T[2]) / zg) / k;
    if (!(g4 * s2 < 0.01)) return g;
    const double C = 1.01 * (g4 * (1 + s2) / (1 - g4 * s2) + (g3 * (1 + s2) + er) / (1 - g3 * s2 - er) + 8 * u);
    const double f[2] = {std::fabs(p.cfx), std::fabs(p.cfy)}, pp[2] = {std::fabs(p.cppx), std::fabs(p.cppy)};
    double a[2], b[2];
    for (int i = 0; i < 2; ++i) {
        // row i of R (column-major storage): own-axis coefficient R[4 i + ... ] -> R[i + 3 i] = R[4 i]; the other one R[i + 3 (1 - i)]
        const double own = std::fabs(p.R[i + 3 * i]), other = std::fabs(p.R[i + 3 * (1 - i)]);
        const double alpha = own / k, beta = (other * nmax[1 - i] + std::fabs(p.R[6 + i]) + std::fabs(p.T[i]) / zg) / k;
        const double d0 = 1.01 * (4 * u * pp[i] + u + 2 * u * (pp[i] + 0.5));
        a[i] = PIPE_GUARD_SAFETY * (beta * f[i] * C + d0) + 1e-5;
        b[i] = PIPE_GUARD_SAFETY * alpha * f[i] * C;
        if (!(a[i] + b[i] * nmax[i] < 0.05)) return g;     // a guard that wide would send every pixel to the exact chain
    }
    g.ok = true;
    g.ax = (float)a[0]; g.bx = (float)b[0];
    g.ay = (float)a[1]; g.by = (float)b[1];
    return g;
}

// Where the taps land vertically, per depth row and per 128-pixel segment of the colour row.  A rotation about the
// optical axis shears the tap rows along x (5 mrad: 6 rows across 1280 px) and the other two axes bend them (terms in
// x*y and y*y, about a row across a 720p frame), so one window of whole colour rows per tile would have to be several rows
// taller than the tile.  Instead every (depth row, segment) gets the exact range of colour rows its taps can reach for
// depths >= z_near -- sampled on the host in double, every 4th column, eight depths, 0.05 rows of slack; a tap outside
// costs an exact re-evaluation with a global load, never a wrong byte -- and a tile stages, per segment, the union
// over its rows.  z_near is 0.2 m unless a y offset between the sensors makes the windows of such near points tall (the
// parallax grows with 1 / z): then 0.3 or 0.45 m, whichever first brings a depth row's window down to 4 colour rows.
struct PipeSegWindow {
    bool ok = false;
    int n_segs = 0, H = 0;
    std::vector<int> lo, hi;      // [H][n_segs] first / last colour row (not clamped to the frame)
};

template <class P>
inline PipeSegWindow pipe_seg_window(const P &p, double z_near = 0.2) {
    PipeSegWindow w;
    w.n_segs = (p.CW + PIPE_SEG_PX - 1) / PIPE_SEG_PX;
    w.H = p.H;
    if (w.n_segs > PIPE_MAX_SEGS || w.n_segs < 1 || p.H < 1) return w;
    const int NONE = 1 << 30;
    w.lo.assign((size_t)p.H * w.n_segs, NONE);
    w.hi.assign((size_t)p.H * w.n_segs, -NONE);
    const double zs[] = {z_near, z_near * 1.25, z_near * 1.75, z_near * 2.5, std::max(1.0, z_near * 4), std::max(2.0, z_near * 8), 8.0, 65.0};
    for (int y = 0; y < p.H; ++y) {
        int *lo = &w.lo[(size_t)y * w.n_segs], *hi = &w.hi[(size_t)y * w.n_segs];
        const double ny = ((double)y - p.ppy) / p.fy;
        for (int ix = 0;; ix += 4) {
            const int x = std::min(ix, p.W - 1);
            const double nx = ((double)x - p.ppx) / p.fx;
            for (double z : zs) {
                const double X = nx * z, Y = ny * z;
                const double t0 = p.R[0] * X + p.R[3] * Y + p.R[6] * z + p.T[0];
                const double t1 = p.R[1] * X + p.R[4] * Y + p.R[7] * z + p.T[1];
                const double t2 = p.R[2] * X + p.R[5] * Y + p.R[8] * z + p.T[2];
                if (!(t2 > 1e-6)) return w;
                const double xa = p.cfx * t0 / t2 + p.cppx + 0.5, ya = p.cfy * t1 / t2 + p.cppy + 0.5;
                if (!(std::fabs(ya) < 1e6)) return w;
                const int r0 = (int)std::floor(ya - 0.05), r1 = (int)std::floor(ya + 0.05);
                // columns between the samples, and a tap next to a segment border, may land on either side
                for (double dx : {-2.5, 2.5}) {
                    const int col = (int)std::floor(std::min(std::max(xa + dx, 0.0), p.CW - 1.0));
                    const int s = col / PIPE_SEG_PX;
                    lo[s] = std::min(lo[s], r0);
                    hi[s] = std::max(hi[s], r1);
                }
            }
            if (x == p.W - 1) break;
        }
        // segments no tap of this row reaches (colour wider than the depth field of view): the nearest one's rows
        for (int s = 0; s < w.n_segs; ++s) {
            if (lo[s] != NONE) continue;
            int best = -1;
            for (int t = 0; t < w.n_segs; ++t)
                if (hi[t] != -NONE && lo[t] != NONE && (t < s ? true : t > s) && (best < 0 || std::abs(t - s) < std::abs(best - s)) &&
                    w.lo[(size_t)y * w.n_segs + t] != NONE)
                    best = t;
            if (best < 0) return w;
            // (mark as borrowed by writing hi first: lo[s] stays NONE until both are set)
            hi[s] = hi[best];
            lo[s] = lo[best];
        }
    }
    w.ok = true;
    return w;
}

// first colour row and row count of segment s for the tile of rt depth rows starting at row0 (not clamped)
inline void pipe_seg_tile(const PipeSegWindow &w, int row0, int rt, int s, int &lo, int &n) {
    int a = 1 << 30, b = -(1 << 30);
    for (int y = row0; y < row0 + rt && y < w.H; ++y) {
        a = std::min(a, w.lo[(size_t)y * w.n_segs + s]);
        b = std::max(b, w.hi[(size_t)y * w.n_segs + s]);
    }
    lo = a;
    n = b - a + 1;
}

// rows a segment window needs with rt depth rows per tile (max over tiles and segments)
inline int pipe_seg_rows(const PipeSegWindow &w, int H, int rt) {
    int rows = 1;
    for (int row0 = 0; row0 + rt <= H; row0 += rt)
        for (int s = 0; s < w.n_segs; ++s) {
            int lo, n;
            pipe_seg_tile(w, row0, rt, s, lo, n);
            rows = std::max(rows, n);
        }
    return rows;
}

}  // namespace pcs
