/* CPU check of the division-by-constant used by the pipelined kernel
 * (pointcloud_stitching_b200/csrc/pcs_k1_pipe.cuh): with y = RN(1/b),
 *     q0 = RN(a*y); r0 = fma(-b,q0,a); q1 = fma(r0,y,q0); r1 = fma(-b,q1,a); q2 = fma(r1,y,q1)
 * must equal the IEEE quotient a/b for every float a in [2^-12, 2^24) and every listed b.
 * fmaf here is the same single-rounding operation as the GPU's FFMA.  Exit code 0 = all equal. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

int main(void)
{
    static const float widths[] = {1000 /* CONV_RATE of the stitcher's unpack */, 8, 24, 64, 96, 128, 256, 320, 424, 480, 640, 720, 848, 1024, 1080, 1280, 1920, 2048, 3840};
    long long bad = 0, total = 0;
    for (unsigned wi = 0; wi < sizeof widths / sizeof *widths; ++wi) {
        const float b = widths[wi], y = 1.0f / b, nb = -b;
        long long bad_w = 0;
        const uint32_t lo = 0x39800000u /* 2^-12 */, hi = 0x4B800000u /* 2^24 */;
#pragma omp parallel for reduction(+ : bad_w) schedule(static)
        for (uint32_t bits = lo; bits < hi; ++bits) {
            float a;
            memcpy(&a, &bits, 4);
            const float q0 = a * y;
            const float r0 = fmaf(nb, q0, a);
            const float q1 = fmaf(r0, y, q0);
            const float r1 = fmaf(nb, q1, a);
            const float q2 = fmaf(r1, y, q1);
            const float want = a / b;
            if (q2 != want) ++bad_w;
        }
        total += (long long)(hi - lo);
        bad += bad_w;
        if (bad_w) printf("width %g: %lld mismatches\n", b, bad_w);
    }
    printf("%lld quotients checked, %lld mismatches\n", total, bad);
    return bad ? 1 : 0;
}
