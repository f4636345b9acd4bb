/* TEST INFRASTRUCTURE — CPU restatement of UltraGrid's CUDA DXT encoders (cuda_dxt/cuda_dxt.cu).
 *
 * The reference arithmetic is single-precision on the GPU and its result depends on how ptxas contracted
 * the expressions into FMAs.  That tree was read from the SASS of the reference built with nvcc 12.9 for
 * sm_100a (default --fmad=true) and is restated here with fmaf(); compile with -ffp-contract=off so that
 * the C compiler adds no contraction of its own.  Two things a CPU cannot reproduce bit-for-bit:
 *   - __fdividef(1.0f, x) is MUFU.RCP (approximate, <= 1 ulp); 1.0f / x (correctly rounded) is used here;
 * so palettes are exact and a very small fraction of DXT1 index words may differ from the GPU.  The
 * BIT-EXACT oracle for the GPU tests is the unmodified reference kernel, oracle/_ref/libcuda_dxt_ref.so;
 * this file is pinned against it on the GPU box (tests/test_dxt_gpu.py::test_cpu_oracle_vs_reference) and is
 * what the CPU-only tests and smoke() use.
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <string.h>

#define API __attribute__((visibility("default")))

static const float K255 = 0.00392156862745f; /* cuda_dxt.cu:666 */

static float sat(float v) { return v < 0.f ? 0.f : v > 1.f ? 1.f : (v != v ? 0.f : v); } /* __saturatef */

/* yuv_to_rgb, cuda_dxt.cu:444-451 as contracted: see dxt_device.cuh */
static void yuv_px(uint8_t Y, uint8_t U, uint8_t V, float *r, float *g, float *b)
{
        const float y = fmaf((float) Y, K255, -0.0625f) * 1.1643f;
        const float u = fmaf((float) U, K255, -0.5f), v = fmaf((float) V, K255, -0.5f);
        *r = fmaf(v, 1.7926f, y);
        *g = fmaf(v, -0.5328f, fmaf(u, -0.2132f, y));
        *b = fmaf(u, 2.1124f, y);
}

/* encode_endpoint, cuda_dxt.cu:424-431: rintf(saturate(c) * levels) — round half to even */
static float quant(float v, float levels) { return rintf(sat(v) * levels); }

/* dxt_encode<1>, cuda_dxt.cu:512-617 */
static void dxt1_block(const float *r, const float *g, const float *b, uint32_t out[2])
{
        float mnr = r[0], mng = g[0], mnb = b[0], mxr = r[0], mxg = g[0], mxb = b[0];
        for (int i = 1; i < 16; ++i) {
                mnr = fminf(mnr, r[i]), mng = fminf(mng, g[i]), mnb = fminf(mnb, b[i]);
                mxr = fmaxf(mxr, r[i]), mxg = fmaxf(mxg, g[i]), mxb = fmaxf(mxb, b[i]);
        }
        const float dr = mxr - mnr, dg = mxg - mng, db = mxb - mnb;
        const float lor = fmaf(dr, 0.0625f, mnr), hir = fmaf(dr, -0.0625f, mxr);
        const float log_ = fmaf(dg, 0.0625f, mng), hig = fmaf(dg, -0.0625f, mxg);
        const float lob = fmaf(db, 0.0625f, mnb), hib = fmaf(db, -0.0625f, mxb);
        const float sr = lor + hir, sg = log_ + hig, sb = lob + hib;
        float covx = 0.f, covy = 0.f;
        for (int i = 0; i < 16; ++i) {
                const float er = fmaf(sr, -0.5f, r[i]), eg = fmaf(sg, -0.5f, g[i]), eb = fmaf(sb, -0.5f, b[i]);
                covx = fmaf(er, eb, covx);
                covy = fmaf(eg, eb, covy);
        }
        const int swr = covx < 0.f, swg = covy < 0.f;
        const float maxr = swr ? lor : hir, minr = swr ? hir : lor;
        const float maxg = swg ? log_ : hig, ming = swg ? hig : log_;
        const float qxr = quant(maxr, 31.f), qxg = quant(maxg, 63.f), qxb = quant(hib, 31.f);
        const float qnr = quant(minr, 31.f), qng = quant(ming, 63.f), qnb = quant(lob, 31.f);
        const uint32_t max_code = ((uint32_t) qxr << 11) + ((uint32_t) qxg << 5) + (uint32_t) qxb;
        const uint32_t min_code = ((uint32_t) qnr << 11) + ((uint32_t) qng << 5) + (uint32_t) qnb;
        uint32_t indices = 0;
        if (max_code != min_code) {
                const float k31 = 0.0322580645161f, k63 = 0.015873015873f;
                const float ex_r = qxr * k31, ex_g = qxg * k63, ex_b = qxb * k31;
                const float dir_r = fmaf(qnr, k31, -ex_r), dir_g = fmaf(qng, k63, -ex_g), dir_b = fmaf(qnb, k31, -ex_b);
                const float len2 = fmaf(dir_b, dir_b, fmaf(dir_r, dir_r, dir_g * dir_g));
                const float inv = 1.0f / len2; /* GPU: MUFU.RCP */
                const float tr = dir_r * inv, tg = dir_g * inv, tb = dir_b * inv;
                const float bias = fmaf(ex_b, tb, fmaf(ex_r, tr, ex_g * tg));
                for (int i = 0; i < 16; ++i) {
                        const float t = fmaf(b[i], tb, fmaf(r[i], tr, g[i] * tg));
                        const float x = fmaf(sat(t - bias), 3.0f, 0.5f);
                        indices += (uint32_t) x << (2 * i);
                }
        }
        const int swap_end = max_code < min_code;
        if (swap_end) {
                indices = ~indices;
        }
        const uint32_t lsbs = indices & 0x55555555u, msbs = indices & 0xaaaaaaaau;
        indices = msbs ^ (2 * lsbs + (msbs >> 1));
        out[0] = swap_end ? min_code + (max_code << 16) : max_code + (min_code << 16);
        out[1] = indices;
}

/* dxt_encode<6> (DXT5-YCoCg), cuda_dxt.cu:471-509 with helpers :141-410, as compiled (see dxt6_device.cuh).
 * Everything here is IEEE single/double with explicit fma, so this restatement is bit-exact with the GPU. */
static uint32_t roundu(float x) { return (uint32_t) roundf(x); } /* CUDA roundf + (u32): trunc(add.rz(x, .5)), x >= 0 */
static float satadd(float a, float b) { return sat(a + b); }
static void dxt6_block(const float *r, const float *g, const float *b, uint32_t out[4])
{
        const float off = 0.50196081399917602539f; /* (float)(128.0/255.0), :139 */
        const double offd = (double) off;
        float Y[16], Co[16], Cg[16];
        for (int i = 0; i < 16; ++i) { /* :141-148, double sub-expressions */
                const double dr = r[i], dg = g[i], db = b[i], g2 = dg + dg;
                Y[i] = (float) (((dr + g2) + db) * 0.25);
                Co[i] = (float) fma((dr + dr) - (db + db), 0.25, offd);
                Cg[i] = (float) fma((-dr + g2) - db, 0.25, offd);
        }
        float mnY = Y[0], mxY = Y[0], mnCo = Co[0], mxCo = Co[0], mnCg = Cg[0], mxCg = Cg[0];
        for (int i = 1; i < 16; ++i) {
                mnY = fminf(mnY, Y[i]), mxY = fmaxf(mxY, Y[i]);
                mnCo = fminf(mnCo, Co[i]), mxCo = fmaxf(mxCo, Co[i]);
                mnCg = fminf(mnCg, Cg[i]), mxCg = fmaxf(mxCg, Cg[i]);
        }
        const float sCo = mnCo + mxCo, sCg = mnCg + mxCg; /* :260-270 */
        float cov = 0.f;
        for (int i = 0; i < 16; ++i) {
                cov = fmaf(fmaf(sCo, -0.5f, Co[i]), fmaf(sCg, -0.5f, Cg[i]), cov);
        }
        if (cov < 0.f) {
                const float t = mxCg;
                mxCg = mnCg, mnCg = t;
        }
        const float eXo = mxCo - off, eXg = mxCg - off, eNo = mnCo - off, eNg = mnCg - off; /* :241-258 */
        const float m = fmaxf(fmaxf(fabsf(eNo), fabsf(eNg)), fmaxf(fabsf(eXo), fabsf(eXg)));
        uint32_t scale = 1;
        if (m < (float) (64.0 / 255.0)) {
                scale = 2;
        }
        if (m < (float) (32.0 / 255.0)) {
                scale = 4;
        }
        const float fs = (float) scale, inv_s = 1.0f / fs;
        const float sXo = fmaf(eXo, fs, off), sXg = fmaf(eXg, fs, off), sNo = fmaf(eNo, fs, off), sNg = fmaf(eNg, fs, off); /* :279-280 */
        const float kIns = (float) ((8.0 / 255.0) / 16.0); /* :184 */
        const float insO = fmaf(sXo - sNo, 0.0625f, -kIns), insG = fmaf(sXg - sNg, 0.0625f, -kIns);
        const float cXo = satadd(sXo, -insO), cXg = satadd(sXg, -insG), cNo = satadd(sNo, insO), cNg = satadd(sNg, insG);
        const uint32_t qXo = roundu(cXo * 31.f), qXg = roundu(cXg * 63.f), qNo = roundu(cNo * 31.f), qNg = roundu(cNg * 63.f); /* :284-288 */
        out[2] = ((qXo << 11) | (qXg << 5) | (scale - 1)) | (((qNo << 11) | (qNg << 5) | (scale - 1)) << 16); /* :291-312 */
        const float k255 = (float) (1.0 / 255.0);
#define EXPAND(e) fmaf(fmaf((float) (e), k255, -off), inv_s, off) /* :294-304 */
        const float pXo = EXPAND((qXo << 3) | (qXo >> 2)), pXg = EXPAND((qXg << 2) | (qXg >> 4));
        const float pNo = EXPAND((qNo << 3) | (qNo >> 2)), pNg = EXPAND((qNg << 2) | (qNg >> 4));
#undef EXPAND
        const float q13 = (float) (1.0 / 3.0), q23 = (float) (2.0 / 3.0); /* :321-322 */
        const float c2o = fmaf(pNo, q13, pXo * (1.0f - q13)), c2g = fmaf(pXg, 1.0f - q13, pNg * q13);
        const float c3o = fmaf(pXo, 1.0f - q23, pNo * q23), c3g = fmaf(pXg, 1.0f - q23, pNg * q23);
        uint32_t cidx = 0;
        for (int i = 0; i < 16; ++i) { /* :326-344 */
#define DIST(co, cg) fmaf(Co[i] - (co), Co[i] - (co), (Cg[i] - (cg)) * (Cg[i] - (cg)))
                const float d0 = DIST(pXo, pXg), d1 = DIST(pNo, pNg), d2 = DIST(c2o, c2g), d3 = DIST(c3o, c3g);
#undef DIST
                const uint32_t bx = d0 > d3, by = d1 > d2, bz = d0 > d2, bw = d1 > d3, b4 = d2 > d3;
                cidx |= ((bx & b4) | (((by & bz) | (bx & bw)) << 1)) << (2 * i);
        }
        out[3] = cidx;
        const float insY = (float) fma((double) (mxY - mnY), 1.0 / 32.0, -((16.0 / 255.0) / 32.0)); /* :176-181 */
        const float nY = satadd(mnY, insY), xY = satadd(mxY, -insY);
        const uint32_t a0 = roundu(nY * 255.f), a1 = roundu(xY * 255.f); /* :350-357 */
        const float mid = (xY - nY) / 14.0f;                           /* :364 */
        const double dX = xY, dN = nY, dM = mid, k7 = 1.0 / 7.0;
        float ab[7];
        ab[0] = nY + mid;
        ab[1] = (float) fma(fma(dX, 6.0, dN), k7, dM);
        ab[2] = (float) fma(fma(dX, 5.0, dN + dN), k7, dM);
        ab[3] = (float) fma(fma(dX, 4.0, dN * 3.0), k7, dM);
        ab[4] = (float) fma(fma(dX, 3.0, dN * 4.0), k7, dM);
        ab[5] = (float) fma(fma(dX, 2.0, dN * 5.0), k7, dM);
        ab[6] = (float) fma(fma(dN, 6.0, dX), k7, dM);
        uint32_t ix = 0, iy = 0;
        for (int i = 0; i < 16; ++i) { /* :374-407 */
                uint32_t idx = 1;
                for (int k = 0; k < 7; ++k) {
                        idx += Y[i] <= ab[k];
                }
                idx &= 7u;
                idx ^= (2u > idx);
                if (i < 6) {
                        ix |= idx << (3 * i + 16);
                }
                if (i == 5) {
                        iy = idx >> 1;
                }
                if (i > 5) {
                        iy |= idx << (3 * i - 16);
                }
        }
        out[0] = (a0 << 8) | a1 | ix;
        out[1] = iy;
}

/* dxt_kernel + dxt_launch, cuda_dxt.cu:622-760: packed 3-byte source, negative size_y = bottom-up */
static int dxt_packed3(const uint8_t *src, uint32_t *out, int sx, int sy, int yuv, int type)
{
        int mirrored = 0;
        if (sy < 0) {
                mirrored = 1, sy = -sy;
        }
        if ((sx & 3) || (sy & 3)) {
                return -1;
        }
        for (int by = 0; by < sy / 4; ++by) {
                for (int bx = 0; bx < sx / 4; ++bx) {
                        float r[16], g[16], b[16];
                        for (int y = 0; y < 4; ++y) {
                                int row = by * 4 + y;
                                if (mirrored) {
                                        row = sy - 1 - row;
                                }
                                const uint8_t *p = src + ((size_t) row * sx + bx * 4) * 3;
                                for (int x = 0; x < 4; ++x, p += 3) {
                                        const int i = 4 * y + x;
                                        if (yuv) {
                                                yuv_px(p[0], p[1], p[2], &r[i], &g[i], &b[i]);
                                        } else {
                                                r[i] = p[0] * K255, g[i] = p[1] * K255, b[i] = p[2] * K255;
                                        }
                                }
                        }
                        if (type == 1) {
                                dxt1_block(r, g, b, out + 2 * ((size_t) by * (sx / 4) + bx));
                        } else {
                                dxt6_block(r, g, b, out + 4 * ((size_t) by * (sx / 4) + bx));
                        }
                }
        }
        return 0;
}
API int orc_rgb_to_dxt1(const uint8_t *src, uint32_t *out, int sx, int sy) { return dxt_packed3(src, out, sx, sy, 0, 1); }
API int orc_yuv_to_dxt1(const uint8_t *src, uint32_t *out, int sx, int sy) { return dxt_packed3(src, out, sx, sy, 1, 1); }
API int orc_rgb_to_dxt6(const uint8_t *src, uint32_t *out, int sx, int sy) { return dxt_packed3(src, out, sx, sy, 0, 6); }
API int orc_yuv_to_dxt6(const uint8_t *src, uint32_t *out, int sx, int sy) { return dxt_packed3(src, out, sx, sy, 1, 6); }

/* yuv422_to_yuv444_kernel, cuda_dxt.cu:697-732: chroma replication */
API void orc_yuv422_to_yuv444(const uint8_t *src, uint8_t *out, int pix_count)
{
        for (int i = 0; i + 1 < pix_count; i += 2, src += 4, out += 6) {
                out[0] = src[1], out[1] = src[0], out[2] = src[2];
                out[3] = src[3], out[4] = src[0], out[5] = src[2];
        }
}

/* the pair run by src/video_compress/cuda_dxt.cpp:223-257 for UYVY input: 422->444 then yuv_to_dxt1 */
static int uyvy_to_dxt(const uint8_t *src, uint32_t *out, int sx, int sy, long pitch, int type)
{
        int mirrored = 0;
        if (sy < 0) {
                mirrored = 1, sy = -sy;
        }
        if ((sx & 3) || (sy & 3)) {
                return -1;
        }
        if (pitch == 0) {
                pitch = (long) sx * 2;
        }
#pragma omp parallel for schedule(static)  /* block rows are independent; used by bench.py's cpu_baseline */
        for (int by = 0; by < sy / 4; ++by) {
                for (int bx = 0; bx < sx / 4; ++bx) {
                        float r[16], g[16], b[16];
                        for (int y = 0; y < 4; ++y) {
                                int row = by * 4 + y;
                                if (mirrored) {
                                        row = sy - 1 - row;
                                }
                                const uint8_t *p = src + (size_t) row * pitch + bx * 8;
                                for (int x = 0; x < 4; ++x) {
                                        const uint8_t *q = p + (x / 2) * 4;
                                        yuv_px(q[1 + 2 * (x & 1)], q[0], q[2], &r[4 * y + x], &g[4 * y + x], &b[4 * y + x]);
                                }
                        }
                        if (type == 1) {
                                dxt1_block(r, g, b, out + 2 * ((size_t) by * (sx / 4) + bx));
                        } else {
                                dxt6_block(r, g, b, out + 4 * ((size_t) by * (sx / 4) + bx));
                        }
                }
        }
        return 0;
}
API int orc_uyvy_to_dxt1(const uint8_t *src, uint32_t *out, int sx, int sy, long pitch) { return uyvy_to_dxt(src, out, sx, sy, pitch, 1); }
API int orc_uyvy_to_dxt6(const uint8_t *src, uint32_t *out, int sx, int sy, long pitch) { return uyvy_to_dxt(src, out, sx, sy, pitch, 6); }

/* DXT1 block decoder (S3TC, 4-colour mode) — used only to sanity-check that encoded blocks reproduce the
 * image (PSNR), never for bit-exact comparisons. */
API void orc_dxt1_decode(const uint32_t *in, uint8_t *rgb, int sx, int sy)
{
        for (int by = 0; by < sy / 4; ++by) {
                for (int bx = 0; bx < sx / 4; ++bx) {
                        const uint32_t pal = in[2 * ((size_t) by * (sx / 4) + bx)], idx = in[2 * ((size_t) by * (sx / 4) + bx) + 1];
                        int c[4][3];
                        for (int k = 0; k < 2; ++k) {
                                const uint32_t v = (pal >> (16 * k)) & 0xffff;
                                const int r5 = v >> 11, g6 = (v >> 5) & 63, b5 = v & 31;
                                c[k][0] = (r5 << 3) | (r5 >> 2), c[k][1] = (g6 << 2) | (g6 >> 4), c[k][2] = (b5 << 3) | (b5 >> 2);
                        }
                        for (int ch = 0; ch < 3; ++ch) {
                                c[2][ch] = (2 * c[0][ch] + c[1][ch] + 1) / 3;
                                c[3][ch] = (c[0][ch] + 2 * c[1][ch] + 1) / 3;
                        }
                        for (int i = 0; i < 16; ++i) {
                                const int k = (idx >> (2 * i)) & 3;
                                uint8_t *p = rgb + (((size_t) by * 4 + i / 4) * sx + bx * 4 + i % 4) * 3;
                                p[0] = c[k][0], p[1] = c[k][1], p[2] = c[k][2];
                        }
                }
        }
}

/* DXT5-YCoCg -> RGB/BGR as the reference tool cuda_dxt/dxt62tga.c:24-108 does it (all in double, (int)(255 v + 0.5), clamp), and DXT1
 * by the same rule for the colour block (+ 3-colour mode).  Built with -ffp-contract=off like the tool's plain gcc build. */
static uint8_t to_byte(double s)
{
        const int is = (int) (s + 0.5);
        return is > 255 ? 255 : is < 0 ? 0 : is;
}
API void orc_dxt5ycocg_to_rgb(const uint64_t *in, uint8_t *out, int sx, int sy, long pitch, int bgr)
{
        for (int by = 0; by < sy / 4; ++by) {
                for (int bx = 0; bx < sx / 4; ++bx, in += 2) {
                        uint64_t ac = in[0], cc = in[1];
                        const double a0 = (ac & 0xFF) / 255.0, a1 = ((ac >> 8) & 0xFF) / 255.0;
                        double ap[8] = { a0, a1 };
                        for (int k = 2; k < 8; ++k) {
                                ap[k] = a0 > a1 ? ((8 - k) * a0 + (k - 1) * a1) / 7.0 : k < 6 ? ((6 - k) * a0 + (k - 1) * a1) / 5.0 : k == 6 ? 0.0 : 1.0;
                        }
                        double r[4], g[4], b[4];
                        for (int k = 0; k < 2; ++k) {
                                b[k] = ((cc >> (16 * k)) & 0x1F) / 31.0, g[k] = ((cc >> (16 * k + 5)) & 0x3F) / 63.0, r[k] = ((cc >> (16 * k + 11)) & 0x1F) / 31.0;
                        }
                        b[2] = (2.0 * b[0] + 1.0 * b[1]) / 3.0, g[2] = (2.0 * g[0] + 1.0 * g[1]) / 3.0, r[2] = (2.0 * r[0] + 1.0 * r[1]) / 3.0;
                        b[3] = (1.0 * b[0] + 2.0 * b[1]) / 3.0, g[3] = (1.0 * g[0] + 2.0 * g[1]) / 3.0, r[3] = (1.0 * r[0] + 2.0 * r[1]) / 3.0;
                        ac >>= 16, cc >>= 32;
                        for (int i = 0; i < 16; ++i, ac >>= 3, cc >>= 2) {
                                const double a = ap[ac & 7];
                                const int k = cc & 3;
                                const double scale = 1.0 / (31.875 * b[k] + 1.0);
                                const double co = (r[k] - 5.01960814E-01) * scale, cg = (g[k] - 5.01960814E-01) * scale;
                                uint8_t *p = out + (size_t) (by * 4 + i / 4) * pitch + (size_t) (bx * 4 + i % 4) * 3;
                                p[bgr ? 2 : 0] = to_byte(((a + co) - cg) * 255.0), p[1] = to_byte((a + cg) * 255.0), p[bgr ? 0 : 2] = to_byte(((a - co) - cg) * 255.0);
                        }
                }
        }
}
API void orc_dxt1_to_rgb(const uint32_t *in, uint8_t *out, int sx, int sy, long pitch, int bgr)
{
        for (int by = 0; by < sy / 4; ++by) {
                for (int bx = 0; bx < sx / 4; ++bx, in += 2) {
                        const uint32_t c0 = in[0] & 0xffff, c1 = in[0] >> 16;
                        double r[4], g[4], b[4];
                        r[0] = (c0 >> 11) / 31.0, g[0] = ((c0 >> 5) & 63) / 63.0, b[0] = (c0 & 31) / 31.0;
                        r[1] = (c1 >> 11) / 31.0, g[1] = ((c1 >> 5) & 63) / 63.0, b[1] = (c1 & 31) / 31.0;
                        if (c0 > c1) {
                                r[2] = (2.0 * r[0] + r[1]) / 3.0, g[2] = (2.0 * g[0] + g[1]) / 3.0, b[2] = (2.0 * b[0] + b[1]) / 3.0;
                                r[3] = (r[0] + 2.0 * r[1]) / 3.0, g[3] = (g[0] + 2.0 * g[1]) / 3.0, b[3] = (b[0] + 2.0 * b[1]) / 3.0;
                        } else {
                                r[2] = (r[0] + r[1]) * 0.5, g[2] = (g[0] + g[1]) * 0.5, b[2] = (b[0] + b[1]) * 0.5;
                                r[3] = g[3] = b[3] = 0.0;
                        }
                        uint32_t idx = in[1];
                        for (int i = 0; i < 16; ++i, idx >>= 2) {
                                const int k = idx & 3;
                                uint8_t *p = out + (size_t) (by * 4 + i / 4) * pitch + (size_t) (bx * 4 + i % 4) * 3;
                                p[bgr ? 2 : 0] = to_byte(r[k] * 255.0), p[1] = to_byte(g[k] * 255.0), p[bgr ? 0 : 2] = to_byte(b[k] * 255.0);
                        }
                }
        }
}

API void orc_set_threads(int n) { omp_set_num_threads(n > 0 ? n : omp_get_num_procs()); }
API int orc_get_max_threads(void) { return omp_get_max_threads(); }
