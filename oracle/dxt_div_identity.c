/* TEST INFRASTRUCTURE (never part of the product).  The DXT5-YCoCg decode kernel (ultragrid_b200/csrc/dxt_decode_kernels.cu, div_small) replaces the two
 * divisions of the alpha palette of cuda_dxt/dxt62tga.c:38-62 -  ((8-k) a0 + (k-1) a1) / 7.0  and  ((6-k) a0 + (k-1) a1) / 5.0  with a0, a1 = code / 255.0 -
 * by q = RN(x * r), q' = fma(fma(-d, q, x), r, q) with r = RN(1 / d).  This file compares q' with the IEEE division for EVERY value x the palette can
 * take (256 x 256 endpoint pairs, all entries: 327 424 quotients) and returns the number of differences (0).  Built without FP contraction. */
#include <math.h>

__attribute__((visibility("default"))) long orc_dxt_div_identity(long *count)
{
        const double r7 = 1.0 / 7.0, r5 = 1.0 / 5.0;
        long bad = 0, n = 0;
        for (int c0 = 0; c0 < 256; ++c0) {
                for (int c1 = 0; c1 < 256; ++c1) {
                        volatile double a0 = c0 / 255.0, a1 = c1 / 255.0;
                        const int seven = a0 > a1;
                        const double d = seven ? 7.0 : 5.0, r = seven ? r7 : r5;
                        for (int k = 2; k < (seven ? 8 : 6); ++k) {
                                volatile double m0 = (double) ((seven ? 8 : 6) - k) * a0, m1 = (double) (k - 1) * a1;
                                volatile double x = m0 + m1;
                                const double q = x * r;
                                const double q2 = fma(fma(-d, q, x), r, q);
                                bad += q2 != x / d;
                                ++n;
                        }
                }
        }
        if (count) {
                *count = n;
        }
        return bad;
}
