/* TEST INFRASTRUCTURE (never part of the product path).  Exhaustive check of the FP64 re-association used by
 * ultragrid_b200/csrc/dxt6_device.cuh for ConvertRGBToYCoCg (reference: cuda_dxt/cuda_dxt.cu:141-148 as compiled):
 * the reference's 11 double operations per pixel and the kernel's 8 give bit-identical floats for every pixel both loaders can produce -
 * all 2^24 (Y, U, V) byte triples through the YUV -> RGB loader (cuda_dxt.cu:444-451) and all 2^24 (R, G, B) byte triples through the RGB loader.
 * Build: gcc -O2 -ffp-contract=off ycocg_identity.c -lm ; exit status 0 = identical. */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static float kInv255 = 0.00392156862745f;
int main(void)
{
        const float off = 0.50196081399917602539f; const double offd = (double) off;
        long bad = 0, n = 0;
        for (int mode = 0; mode < 2; ++mode)
        for (int a = 0; a < 256; ++a) for (int b_ = 0; b_ < 256; ++b_) for (int c = 0; c < 256; ++c) {
                float r, g, b;
                if (mode == 0) {  // YUV source: a = Y, b_ = U, c = V
                        const float y = fmaf((float) a, kInv255, -0.0625f) * 1.1643f;
                        const float u = fmaf((float) b_, kInv255, -0.5f), v = fmaf((float) c, kInv255, -0.5f);
                        r = fmaf(v, 1.7926f, y); g = fmaf(v, -0.5328f, fmaf(u, -0.2132f, y)); b = fmaf(u, 2.1124f, y);
                } else {
                        r = (float) a * kInv255; g = (float) b_ * kInv255; b = (float) c * kInv255;
                }
                const double dr = r, dg = g, db = b;
                const double g2 = dg + dg;
                const float Y0 = (float) (((dr + g2) + db) * 0.25);
                const float Co0 = (float) fma(((dr + dr) - (db + db)), 0.25, offd);
                const float Cg0 = (float) fma(((-dr + g2) - db), 0.25, offd);
                const float Y1 = (float) ((fma(dg, 2.0, dr) + db) * 0.25);
                const float Co1 = (float) fma(dr - db, 0.5, offd);
                const float Cg1 = (float) fma(fma(dg, 2.0, -dr) - db, 0.25, offd);
                ++n;
                if (memcmp(&Y0, &Y1, 4) || memcmp(&Co0, &Co1, 4) || memcmp(&Cg0, &Cg1, 4)) ++bad;
        }
        printf("%ld triples, %ld differ\n", n, bad);
        return bad != 0;
}
