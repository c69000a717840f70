// CPU check of the segment windows the pipelined kernel stages for a rotated depth->colour calibration
// (pointcloud_stitching_b200/csrc/pcs_guard.h, pipe_seg_window): per depth row and 128-pixel segment of the colour row the
// host samples the colour rows a tap can reach (every 4th column, eight depths).  A tap outside its window is not an error
// -- the kernel re-evaluates it from global memory -- but it costs time, so: random pixels and depths >= the near limit
// must land inside the window of their row and segment, and the windows must stay as short as DESIGN.md says.
// Prints "rows <rt=1> <rt=2> misses <n> of <N>" per rig; exit code 0 when no tap misses.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "pcs_guard.h"

struct Params {
    int W, H, CW, CH;
    float ppx, ppy, fx, fy, cppx, cppy, cfx, cfy, cwf, chf, depth_scale;
    float R[9], T[3];
};

static void rotation(double rx, double ry, double rz, float *R) {     // Rz Ry Rx, column-major (rs2_extrinsics)
    const double cx = cos(rx), sx = sin(rx), cy = cos(ry), sy = sin(ry), cz = cos(rz), sz = sin(rz);
    const double m[3][3] = {{cz * cy, cz * sy * sx - sz * cx, cz * sy * cx + sz * sx},
                            {sz * cy, sz * sy * sx + cz * cx, sz * sy * cx - cz * sx},
                            {-sy, cy * sx, cy * cx}};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) R[3 * c + r] = (float)m[r][c];
}

static uint64_t state = 88172645463325252ull;
static double urand() {
    state ^= state << 13; state ^= state >> 7; state ^= state << 17;
    return (double)(state >> 11) / 9007199254740992.0;
}

int main() {
    const int rigs[][4] = {{1280, 720, 1280, 720}, {1280, 720, 1920, 1080}, {848, 480, 848, 480}, {848, 480, 1280, 720}};
    long long total_miss = 0;
    for (const auto &g : rigs) {
        Params p;
        p.W = g[0]; p.H = g[1]; p.CW = g[2]; p.CH = g[3];
        p.fx = p.fy = p.W / 2.f; p.ppx = (p.W - 1) / 2.f; p.ppy = (p.H - 1) / 2.f;
        p.cfx = p.cfy = p.CW / 2.f; p.cppx = (p.CW - 1) / 2.f; p.cppy = (p.CH - 1) / 2.f;
        p.cwf = (float)p.CW; p.chf = (float)p.CH; p.depth_scale = 0.001f;
        rotation(0.004, -0.003, 0.005, p.R);                 // synth.D2C_ROTATION_SMALL: what bench.py --tex rotated uses
        p.T[0] = 0.015f; p.T[1] = 0.0003f; p.T[2] = -0.0002f;
        const double z_near = 0.2;
        const pcs::PipeSegWindow w = pcs::pipe_seg_window(p, z_near);
        if (!w.ok) { std::printf("%dx%d -> %dx%d: no window\n", p.W, p.H, p.CW, p.CH); return 1; }
        long long miss = 0, n = 0;
        for (int i = 0; i < 2000000; ++i) {
            const int x = (int)(urand() * p.W), y = (int)(urand() * p.H);
            const double z = z_near + urand() * urand() * 8.0;
            const double nx = (x - p.ppx) / p.fx, ny = (y - p.ppy) / p.fy, X = nx * z, Y = ny * z;
            const double t0 = p.R[0] * X + p.R[3] * Y + p.R[6] * z + p.T[0], t1 = p.R[1] * X + p.R[4] * Y + p.R[7] * z + p.T[1],
                         t2 = p.R[2] * X + p.R[5] * Y + p.R[8] * z + p.T[2];
            const double xa = p.cfx * t0 / t2 + p.cppx + 0.5, ya = p.cfy * t1 / t2 + p.cppy + 0.5;
            if (xa < 0 || xa >= p.CW || ya < 0 || ya >= p.CH) continue;      // clamped taps sit on the frame's edge rows
            const int seg = (int)xa / pcs::PIPE_SEG_PX, row = (int)std::floor(ya);
            ++n;
            if (row < w.lo[(size_t)y * w.n_segs + seg] || row > w.hi[(size_t)y * w.n_segs + seg]) ++miss;
        }
        const int r1 = pcs::pipe_seg_rows(w, p.H, 1), r2 = pcs::pipe_seg_rows(w, p.H, 2);
        std::printf("%dx%d -> %dx%d: rows %d %d misses %lld of %lld\n", p.W, p.H, p.CW, p.CH, r1, r2, miss, n);
        total_miss += miss;
        if (r1 > 4 || r2 > 6) { std::printf("windows taller than documented\n"); return 1; }
    }
    return total_miss ? 1 : 0;
}
