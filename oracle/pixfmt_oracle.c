/* TEST INFRASTRUCTURE — CPU restatement of the pixel-format part of the hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this; the product
 * (ultragrid_b200/) never links or loads it.  Every function cites the UltraGrid source it restates
 * (paths relative to /root/reference).  Pinned by tests/test_oracle_pinning.py against
 *   (1) the unmodified reference objects built into oracle/_ref/libugref.so (in this container),
 *   (2) the known-answer checksums recorded in SURVEY.md section 6 / BASELINE.md section 2,
 *   (3) the patterns of the reference's own unit tests (test/codec_conversions_test.cpp),
 *   (4) the committed golden vectors tests/golden/ (generated from (1) by tests/golden/make_golden.py).
 * Integer arithmetic only: results must be byte-identical to the reference.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define API __attribute__((visibility("default")))

/* codec_t values, src/types.h:62-112 */
enum { C_RGBA = 1, C_UYVY = 2, C_YUYV = 3, C_VUYA = 4, C_R10k = 5, C_R12L = 6, C_v210 = 7, C_DVS10 = 8, C_RGB = 12, C_BGR = 20,
       C_RG48 = 27, C_I420 = 29, C_Y216 = 30, C_Y416 = 31 };

/* ---- src/color_space.{h,c} ---------------------------------------------------------------------- */
enum { COMP_BASE = 14 }; /* color_space.h:70 */
struct coeffs {
        int y_r, y_g, y_b, cb_r, cb_g, cb_b, cr_r, cr_g, cr_b, y_scale, r_cr, g_cb, g_cr, b_cb;
};

/* COEFFS(), color_space.c:46-128.  depth 0 = full range. */
static struct coeffs compute_coeffs(double kr, double kb, int depth)
{
        const double kg = 1. - kr - kb;
        const double D = 2. * (kr + kg), E = 2. * (1. - kr);
        const double yl = depth == 0 ? 1.0 : (219. * (1 << (depth - 8)) / ((1 << depth) - 1));
        const double cl = depth == 0 ? 1.0 : (224. * (1 << (depth - 8)) / ((1 << depth) - 1));
        const double B = 1 << COMP_BASE;
#define SCALED(x) ((int) (((x) * B) + ((x) > 0 ? 1. : -1.) * 0.5))
        struct coeffs c;
        c.y_r = (int) (kr * yl * B + 0.5);
        c.y_g = (int) (kg * yl * B + 0.5);
        c.y_b = (int) (kb * yl * B + 0.5);
        c.cb_r = (int) (-kr / D * cl * B - 0.5);
        c.cb_g = (int) (-kg / D * cl * B - 0.5);
        c.cb_b = (int) ((1 - kb) / D * cl * B + 0.5);
        c.cr_r = (int) ((1 - kr) / E * cl * B - 0.5);
        c.cr_g = (int) (-kg / E * cl * B - 0.5);
        c.cr_b = (int) (-kb / E * cl * B + 0.5);
        c.y_scale = SCALED(1. / yl);
        c.r_cr = SCALED((2. * (1. - kr)) / cl);
        c.g_cb = SCALED((-kb * (2. * (kr + kg)) / kg) / cl);
        c.g_cr = SCALED((-kr * (2. * (1. - kr)) / kg) / cl);
        c.b_cb = SCALED((2. * (kr + kg)) / cl);
#undef SCALED
        return c;
}

/* get_color_coeffs(CS_DFL|CS_709 -> BT.709, CS_601 -> BT.601), color_space.c:149-183 with the default
 * colour space (no "color-601" param).  cs: 0 = default, 1 = 601, 2 = 709. */
API void orc_get_color_coeffs(int cs, int depth, int out[14])
{
        const struct coeffs c = cs == 1 ? compute_coeffs(.299, .114, depth) : compute_coeffs(.212639, .072192, depth);
        memcpy(out, &c, sizeof c);
}
static struct coeffs cfs709(int depth) { return compute_coeffs(.212639, .072192, depth); }

/* ---- src/video_codec.c --------------------------------------------------------------------------- */
/* codec_info[] block_size_bytes / block_size_pixels / h_align, video_codec.c:120-206 */
static int info(int codec, int *bpx, int *h_align)
{
        switch (codec) {
        case C_RGBA: case C_VUYA: *bpx = 1, *h_align = 1; return 4;
        case C_UYVY: case C_YUYV: *bpx = 2, *h_align = 2; return 4;
        case C_R10k: *bpx = 1, *h_align = 64; return 4;
        case C_R12L: *bpx = 8, *h_align = 8; return 36;
        case C_v210: case C_DVS10: *bpx = 6, *h_align = 48; return 16;
        case C_RGB: case C_BGR: *bpx = 1, *h_align = 1; return 3;
        case C_RG48: *bpx = 1, *h_align = 1; return 6;
        case C_I420: *bpx = 2, *h_align = 2; return 3;
        case C_Y216: *bpx = 2, *h_align = 2; return 8;
        case C_Y416: *bpx = 1, *h_align = 1; return 8;
        }
        *bpx = 1, *h_align = 0;
        return 0;
}
/* vc_get_linesize, video_codec.c:507-521 */
API int orc_vc_get_linesize(unsigned width, int codec)
{
        int bpx, ha;
        const int bytes = info(codec, &bpx, &ha);
        if (bytes == 0) {
                return 0;
        }
        if (ha) {
                width = (width + ha - 1) / ha * ha;
        }
        return (width + bpx - 1) / bpx * bytes;
}
/* vc_get_size, video_codec.c:530-538 */
API int orc_vc_get_size(unsigned width, int codec)
{
        int bpx, ha;
        const int bytes = info(codec, &bpx, &ha);
        return bytes == 0 ? 0 : (width + bpx - 1) / bpx * bytes;
}

/* ---- src/pixfmt_conv.c line converters ----------------------------------------------------------- */
typedef void line_fn(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs);

static int clamp255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static uint32_t rd32(const unsigned char *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static void wr32(unsigned char *p, uint32_t v) { memcpy(p, &v, 4); }

/* vc_copylinev210, pixfmt_conv.c:86-130 */
static void v210_to_uyvy(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
#define A(w) ((((w) >> 0) & 0x3ff) >> 2)
#define B(w) ((((w) >> 10) & 0x3ff) >> 2)
#define C(w) ((((w) >> 20) & 0x3ff) >> 2)
        while (dst_len >= 12) {
                const uint32_t w0 = rd32(src), w1 = rd32(src + 4), w2 = rd32(src + 8), w3 = rd32(src + 12);
                wr32(dst, A(w0) | B(w0) << 8 | C(w0) << 16 | A(w1) << 24);
                wr32(dst + 4, B(w1) | C(w1) << 8 | A(w2) << 16 | B(w2) << 24);
                wr32(dst + 8, C(w2) | A(w3) << 8 | B(w3) << 16 | C(w3) << 24);
                src += 16, dst += 12, dst_len -= 12;
        }
        if (dst_len >= 4) { /* :118-122 */
                const uint32_t w0 = rd32(src), w1 = rd32(src + 4);
                wr32(dst, A(w0) | B(w0) << 8 | C(w0) << 16 | A(w1) << 24);
        }
        if (dst_len >= 8) { /* :123-127 */
                const uint32_t w1 = rd32(src + 4), w2 = rd32(src + 8);
                wr32(dst + 4, B(w1) | C(w1) << 8 | A(w2) << 16 | B(w2) << 24);
        }
#undef A
#undef B
#undef C
}

/* vc_copylineYUYV, pixfmt_conv.c:136-198 (word-wise byte swap; dst_len % 4 == 0 asserted at :156) */
static void yuyv_uyvy(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x <= dst_len - 4; x += 4) {
                dst[x] = src[x + 1], dst[x + 1] = src[x], dst[x + 2] = src[x + 3], dst[x + 3] = src[x + 2];
        }
}

/* copylineYUVtoRGB, pixfmt_conv.c:1065-1094 (rgb16 = 0) */
static void yuv422_to_rgb(unsigned char *dst, const unsigned char *src, int dst_len, int y1o, int y2o, int uo, int vo)
{
        const struct coeffs c = cfs709(8);
        for (int x = 0; x <= dst_len - 6; x += 6) {
                const int y1 = c.y_scale * (src[y1o] - 16), y2 = c.y_scale * (src[y2o] - 16);
                const int u = src[uo] - 128, v = src[vo] - 128;
                src += 4;
                *dst++ = clamp255((y1 + v * c.r_cr) >> COMP_BASE);
                *dst++ = clamp255((y1 + u * c.g_cb + v * c.g_cr) >> COMP_BASE);
                *dst++ = clamp255((y1 + u * c.b_cb) >> COMP_BASE);
                *dst++ = clamp255((y2 + v * c.r_cr) >> COMP_BASE);
                *dst++ = clamp255((y2 + u * c.g_cb + v * c.g_cr) >> COMP_BASE);
                *dst++ = clamp255((y2 + u * c.b_cb) >> COMP_BASE);
        }
}
/* vc_copylineUYVYtoRGB :1102-1108, vc_copylineYUYVtoRGB :1116-1122 */
static void uyvy_to_rgb(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; yuv422_to_rgb(d, s, n, 1, 3, 0, 2); }
static void yuyv_to_rgb(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; yuv422_to_rgb(d, s, n, 0, 2, 1, 3); }

/* vc_copylineUYVYtoRGBA, pixfmt_conv.c:1137-1163: double arithmetic (built with -ffp-contract=off, the
 * reference build has no FMA target either), truncating int conversion */
static void uyvy_to_rgba(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        const uint32_t amask = 0xFFFFFFFFU ^ (0xFFU << rs) ^ (0xFFU << gs) ^ (0xFFU << bs);
        for (int x = 0; x <= dst_len - 8; x += 8) {
                const int u = src[0], y1 = src[1], v = src[2], y2 = src[3];
                src += 4;
                for (int k = 0; k < 2; ++k) {
                        const int y = k ? y2 : y1;
                        int r = 1.164 * (y - 16) + 1.793 * (v - 128);
                        int g = 1.164 * (y - 16) - 0.534 * (v - 128) - 0.213 * (u - 128);
                        int b = 1.164 * (y - 16) + 2.115 * (u - 128);
                        r = clamp255(r), g = clamp255(g), b = clamp255(b);
                        wr32(dst, amask | (uint32_t) r << rs | (uint32_t) g << gs | (uint32_t) b << bs);
                        dst += 4;
                }
        }
}

/* vc_copylineToUYVY, pixfmt_conv.c:1008-1053 */
static void to_uyvy(unsigned char *dst, const unsigned char *src, int dst_len, int roff, int goff, int boff, int pix)
{
        const struct coeffs c = cfs709(8);
        const int count = (dst_len + 3) / 4; /* :1045 */
        for (int x = 0; x < count; ++x) {
                int r = src[roff], g = src[goff], b = src[boff];
                src += pix;
                const int y1 = ((r * c.y_r + g * c.y_g + b * c.y_b) >> COMP_BASE) + 16;
                int u = r * c.cb_r + g * c.cb_g + b * c.cb_b;
                int v = r * c.cr_r + g * c.cr_g + b * c.cr_b;
                r = src[roff], g = src[goff], b = src[boff];
                src += pix;
                const int y2 = ((r * c.y_r + g * c.y_g + b * c.y_b) >> COMP_BASE) + 16;
                u += r * c.cb_r + g * c.cb_g + b * c.cb_b;
                v += r * c.cr_r + g * c.cr_g + b * c.cr_b;
                u = ((u / 2) >> COMP_BASE) + 128; /* C division truncates toward zero, then arithmetic shift */
                v = ((v / 2) >> COMP_BASE) + 128;
                wr32(dst + 4 * x, ((uint32_t) (y2 & 0xFF) << 24) | ((v & 0xFF) << 16) | ((y1 & 0xFF) << 8) | (u & 0xFF));
        }
}
static void rgb_to_uyvy(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; to_uyvy(d, s, n, 0, 1, 2, 3); }   /* :2061-2068 */
static void bgr_to_uyvy(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; to_uyvy(d, s, n, 2, 1, 0, 3); }   /* :2271-2278 */
static void rgba_to_uyvy(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; to_uyvy(d, s, n, 0, 1, 2, 4); }  /* :2309-2316 */
static void rg48_to_uyvy(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; to_uyvy(d, s, n, 1, 3, 5, 6); }  /* :2336-2343 */

/* vc_copylineRGBtoRGBA, pixfmt_conv.c:944-990 */
static void rgb_to_rgba(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        const uint32_t amask = 0xFFFFFFFFU ^ (0xFFU << rs) ^ (0xFFU << gs) ^ (0xFFU << bs);
        for (int x = 0; x <= dst_len - 4; x += 4) {
                const uint32_t r = src[0], g = src[1], b = src[2];
                src += 3;
                wr32(dst + x, amask | r << rs | g << gs | b << bs);
        }
}
/* vc_copylineRGBAtoRGB, pixfmt_conv.c:866-900, SSSE3 build (the reference's tools/Makefile and oracle/_ref use
 * -msse4.1).  QUIRK kept on purpose: the scalar tail loop at :889-895 never advances `src`, so every pixel after
 * the pshufb loop (which runs while x <= dst_len - 24) repeats the first tail pixel. */
static void rgba_to_rgb(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        int x = 0;
        for (; x <= dst_len - 24; x += 12) { /* :880-887: 4 px per step */
                for (int k = 0; k < 4; ++k) {
                        dst[3 * k] = src[4 * k], dst[3 * k + 1] = src[4 * k + 1], dst[3 * k + 2] = src[4 * k + 2];
                }
                src += 16, dst += 12;
        }
        for (; x <= dst_len - 3; x += 3) { /* :889-895 */
                const uint32_t in = rd32(src);
                *dst++ = in & 0xff, *dst++ = (in >> 8) & 0xff, *dst++ = (in >> 16) & 0xff;
        }
}
/* vc_copylineRGBA, pixfmt_conv.c:538-589 */
static void rgba_to_rgba(unsigned char *dst, const unsigned char *src, int len, int rs, int gs, int bs)
{
        if (rs == 0 && gs == 8 && bs == 16) {
                memcpy(dst, src, len);
                return;
        }
        const uint32_t amask = 0xFFFFFFFFU ^ (0xFFU << rs) ^ (0xFFU << gs) ^ (0xFFU << bs);
        for (; len >= 4; len -= 4, src += 4, dst += 4) {
                const uint32_t t = rd32(src);
                wr32(dst, amask | (t & 0xff) << rs | ((t >> 8) & 0xff) << gs | ((t >> 16) & 0xff) << bs);
        }
}
/* vc_copylineRGB, pixfmt_conv.c:732-753 */
static void rgb_to_rgb(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        if (rs == 0 && gs == 8 && bs == 16) {
                memcpy(dst, src, dst_len);
                return;
        }
        for (int x = 0; x <= dst_len - 3; x += 3) {
                const uint32_t w = (uint32_t) src[0] << rs | (uint32_t) src[1] << gs | (uint32_t) src[2] << bs;
                src += 3;
                *dst++ = w & 0xff, *dst++ = (w >> 8) & 0xff, *dst++ = (w >> 16) & 0xff;
        }
}
/* vc_copylineBGRtoRGB, pixfmt_conv.c:2520-2527 */
static void bgr_to_rgb(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; rgb_to_rgb(d, s, n, 16, 8, 0); }
/* vc_memcpy, pixfmt_conv.c:2529-2536 */
static void copy_line(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; memcpy(d, s, n); }

/* ---- v210 family ---------------------------------------------------------------------------------------- */
#define S10(w, sh) (((w) >> (sh)) & 0x3ffu)
/* vc_copylineUYVYtoV210, pixfmt_conv.c:2581-2607 */
static void uyvy_to_v210(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (; dst_len >= 4; dst_len -= 4, dst += 4, src += 3) {
                wr32(dst, (uint32_t) src[0] << 2 | (uint32_t) src[1] << 12 | (uint32_t) src[2] << 22);
        }
}
/* vc_copylineY216toV210, pixfmt_conv.c:2761-2790 */
static void y216_to_v210(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x < (dst_len + 15) / 16; ++x) {
                uint16_t s[12];
                memcpy(s, src + x * 24, 24);
                unsigned char *d = dst + x * 16;
                wr32(d, (uint32_t) (s[1] >> 6) | (uint32_t) (s[0] >> 6) << 10 | (uint32_t) (s[3] >> 6) << 20);
                wr32(d + 4, (uint32_t) (s[2] >> 6) | (uint32_t) (s[5] >> 6) << 10 | (uint32_t) (s[4] >> 6) << 20);
                wr32(d + 8, (uint32_t) (s[7] >> 6) | (uint32_t) (s[6] >> 6) << 10 | (uint32_t) (s[9] >> 6) << 20);
                wr32(d + 12, (uint32_t) (s[8] >> 6) | (uint32_t) (s[11] >> 6) << 10 | (uint32_t) (s[10] >> 6) << 20);
        }
}
static void v210_group(const unsigned char *src, unsigned y[6], unsigned u[3], unsigned v[3])
{
        const uint32_t w0 = rd32(src), w1 = rd32(src + 4), w2 = rd32(src + 8), w3 = rd32(src + 12);
        y[0] = S10(w0, 10), y[1] = S10(w1, 0), y[2] = S10(w1, 20), y[3] = S10(w2, 10), y[4] = S10(w3, 0), y[5] = S10(w3, 20);
        u[0] = S10(w0, 0), u[1] = S10(w1, 10), u[2] = S10(w2, 20);
        v[0] = S10(w0, 20), v[1] = S10(w2, 0), v[2] = S10(w3, 10);
}
/* vc_copylineV210toY216, pixfmt_conv.c:2792-2832 */
static void v210_to_y216(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x < dst_len / 24; ++x) {
                unsigned y[6], u[3], v[3];
                v210_group(src + x * 16, y, u, v);
                uint16_t d[12];
                for (int i = 0; i < 3; ++i) {
                        d[4 * i] = y[2 * i] << 6, d[4 * i + 1] = u[i] << 6, d[4 * i + 2] = y[2 * i + 1] << 6, d[4 * i + 3] = v[i] << 6;
                }
                memcpy(dst + x * 24, d, 24);
        }
}
/* vc_copylineV210toY416, pixfmt_conv.c:2834-2882 */
static void v210_to_y416(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x < dst_len / 48; ++x) {
                unsigned y[6], u[3], v[3];
                v210_group(src + x * 16, y, u, v);
                uint16_t d[24];
                for (int i = 0; i < 6; ++i) {
                        d[4 * i] = u[i / 2] << 6, d[4 * i + 1] = y[i] << 6, d[4 * i + 2] = v[i / 2] << 6, d[4 * i + 3] = 0xFFFFU;
                }
                memcpy(dst + x * 48, d, 48);
        }
}
/* vc_copylineV210toRGB, pixfmt_conv.c:2884-2940 */
static void v210_to_rgb(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        const struct coeffs c = cfs709(8);
        for (int x = 0; x < dst_len; x += 18, src += 16) {
                unsigned y[6], u[3], v[3];
                v210_group(src, y, u, v);
                for (int i = 0; i < 6; ++i) {
                        const int ys = c.y_scale * ((int) (y[i] >> 2) - 16), uu = (int) (u[i / 2] >> 2) - 128, vv = (int) (v[i / 2] >> 2) - 128;
                        int val = (ys + vv * c.r_cr) >> COMP_BASE;
                        *dst++ = val < 1 ? 1 : val > 254 ? 254 : val; /* CLAMP_FULL, color_space.h:96-98 */
                        val = (ys + uu * c.g_cb + vv * c.g_cr) >> COMP_BASE;
                        *dst++ = val < 1 ? 1 : val > 254 ? 254 : val;
                        val = (ys + uu * c.b_cb) >> COMP_BASE;
                        *dst++ = val < 1 ? 1 : val > 254 ? 254 : val;
                }
        }
}
#undef S10

/* ---- RG48 / Y216 / Y416 / VUYA / R10k byte and bit repackers ----------------------------------------------------- */
#define UNUSED3 (void) rs, (void) gs, (void) bs
/* vc_copylineRG48toRGB, pixfmt_conv.c:2030-2042 */
static void rg48_to_rgb(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (int x = 0; x <= dst_len - 3; x += 3, src += 6) {
                *dst++ = src[1], *dst++ = src[3], *dst++ = src[5];
        }
}
/* vc_copylineRG48toRGBA, :2044-2055 */
static void rg48_to_rgba(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        const uint32_t amask = 0xFFFFFFFFU ^ (0xFFU << rs) ^ (0xFFU << gs) ^ (0xFFU << bs);
        for (int x = 0; x <= dst_len - 4; x += 4, src += 6) {
                wr32(dst + x, amask | (uint32_t) src[1] << rs | (uint32_t) src[3] << gs | (uint32_t) src[5] << bs);
        }
}
/* vc_copylineRG48toR10k, :2008-2028 */
static void rg48_to_r10k(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (int x = 0; x <= dst_len - 4; x += 4, src += 6) {
                uint16_t in[3];
                memcpy(in, src, 6);
                const unsigned r = in[0] >> 6, g = in[1] >> 6, b = in[2] >> 6;
                wr32(dst + x, (b & 0x3FU) << 26U | 0x3000000U | (g & 0xFU) << 20U | (b >> 6U) << 16U | (r & 0x3U) << 14U | (g >> 4U) << 8U | r >> 2U);
        }
}
/* vc_copylineRGBAtoRG48, :1336-1351 */
static void rgba_to_rg48(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (int x = 0; x <= dst_len - 6; x += 6, src += 4) {
                *dst++ = 0, *dst++ = src[0], *dst++ = 0, *dst++ = src[1], *dst++ = 0, *dst++ = src[2];
        }
}
/* vc_copylineRGBtoRG48, :1353-1363 */
static void rgb_to_rg48(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (int x = 0; x <= dst_len - 2; x += 2) {
                *dst++ = 0, *dst++ = *src++;
        }
}
/* vc_copylineUYVYtoY216, :2609-2627 */
static void uyvy_to_y216(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (; dst_len >= 8; dst_len -= 8, src += 4) {
                *dst++ = 0, *dst++ = src[1], *dst++ = 0, *dst++ = src[0], *dst++ = 0, *dst++ = src[3], *dst++ = 0, *dst++ = src[2];
        }
}
/* vc_copylineUYVYtoY416, :2629-2665 — the loop tests >= 12 but consumes 16 (kept) */
static void uyvy_to_y416(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        while (dst_len >= 12) {
                *dst++ = 0, *dst++ = src[0], *dst++ = 0, *dst++ = src[1], *dst++ = 0, *dst++ = src[2], *dst++ = 0xFF, *dst++ = 0xFF;
                *dst++ = 0, *dst++ = src[0], *dst++ = 0, *dst++ = src[3], *dst++ = 0, *dst++ = src[2], *dst++ = 0xFF, *dst++ = 0xFF;
                src += 4;
                dst_len -= 16;
        }
        if (dst_len >= 8) {
                *dst++ = 0, *dst++ = src[0], *dst++ = 0, *dst++ = src[1], *dst++ = 0, *dst++ = src[2], *dst++ = 0xFF, *dst++ = 0xFF;
        }
}
/* vc_copylineY216toUYVY, :2728-2743 */
static void y216_to_uyvy(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (; dst_len >= 4; dst_len -= 4, src += 8) {
                *dst++ = src[3], *dst++ = src[1], *dst++ = src[7], *dst++ = src[5];
        }
}
/* vc_copylineY416toUYVY, :2745-2759 */
static void y416_to_uyvy(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (; dst_len >= 4; dst_len -= 4, src += 16) {
                *dst++ = (src[1] + src[9]) / 2, *dst++ = src[3], *dst++ = (src[5] + src[13]) / 2, *dst++ = src[11];
        }
}
/* vc_copylineVUYAtoY416, :2667-2686 */
static void vuya_to_y416(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (; dst_len > 7; dst_len -= 8, src += 4) {
                *dst++ = 0, *dst++ = src[1], *dst++ = 0, *dst++ = src[2], *dst++ = 0, *dst++ = src[0], *dst++ = 0, *dst++ = src[3];
        }
}
/* vc_copylineVUYAtoUYVY, :2688-2703 — src[7] (alpha of the 2nd pixel) is what the reference stores as Y1 (kept) */
static void vuya_to_uyvy(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (; dst_len > 3; dst_len -= 4, src += 8) {
                *dst++ = (src[1] + src[5]) / 2, *dst++ = src[2], *dst++ = (src[0] + src[4]) / 2, *dst++ = src[7];
        }
}
/* vc_copylineVUYAtoRGB, :2705-2726 */
static void vuya_to_rgb(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        const struct coeffs c = cfs709(8);
        for (int x = 0; x < dst_len; x += 3, src += 4) {
                const int v = src[0] - 128, u = src[1] - 128, y = c.y_scale * (src[2] - 16);
                int val = (y + v * c.r_cr) >> COMP_BASE;
                *dst++ = val < 1 ? 1 : val > 254 ? 254 : val;
                val = (y + u * c.g_cb + v * c.g_cr) >> COMP_BASE;
                *dst++ = val < 1 ? 1 : val > 254 ? 254 : val;
                val = (y + u * c.b_cb) >> COMP_BASE;
                *dst++ = val < 1 ? 1 : val > 254 ? 254 : val;
        }
}
/* vc_copylineRGBAtoVUYA, :2280-2302 */
static void rgba_to_vuya(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        const struct coeffs c = cfs709(8);
        for (; dst_len > 3; dst_len -= 4, src += 4) {
                const int r = src[0], g = src[1], b = src[2];
                *dst++ = ((r * c.cr_r + g * c.cr_g + b * c.cr_b) >> COMP_BASE) + 128;
                *dst++ = ((r * c.cb_r + g * c.cb_g + b * c.cb_b) >> COMP_BASE) + 128;
                *dst++ = ((r * c.y_r + g * c.y_g + b * c.y_b) >> COMP_BASE) + 16;
                *dst++ = src[3];
        }
}
/* vc_copyliner10k, :211-272 */
static void r10k_to_rgba(unsigned char *dst, const unsigned char *src, int len, int rs, int gs, int bs)
{
        const uint32_t amask = 0xFFFFFFFFU ^ (0xFFU << rs) ^ (0xFFU << gs) ^ (0xFFU << bs);
        for (; len >= 4; len -= 4, src += 4, dst += 4) {
                const uint32_t r = src[0], g = (src[1] & 0x3FU) << 2 | src[2] >> 6, b = (src[2] & 0xFU) << 4 | src[3] >> 4;
                wr32(dst, amask | r << rs | g << gs | b << bs);
        }
}
/* vc_copyliner10ktoRGB, :331-340 */
static void r10k_to_rgb(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (int x = 0; x < dst_len; x += 3, src += 4) {
                *dst++ = src[0], *dst++ = src[1] << 2 | src[2] >> 6, *dst++ = src[2] << 4 | src[3] >> 4;
        }
}
/* vc_copyliner10ktoRG48, :274-292 */
static void r10k_to_rg48(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (; dst_len > 0; dst_len -= 6, dst += 6, src += 4) {
                const unsigned b2 = src[1], b3 = src[2], b4 = src[3];
                dst[1] = src[0], dst[0] = b2 & 0xC0U, dst[3] = b2 << 2U | b3 >> 6U, dst[2] = (b3 & 0x30U) << 2U;
                dst[5] = (b3 & 0xFU) << 4U | b4 >> 4U, dst[4] = (b4 & 0xCU) << 4U;
        }
}
/* vc_copylineRGBAtoR10k, :2538-2577 */
static void rgba_to_r10k(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        UNUSED3;
        for (; dst_len >= 4; dst_len -= 4, src += 4, dst += 4) {
                const unsigned r = src[0], g = src[1], b = src[2];
                dst[0] = r, dst[1] = g >> 2, dst[2] = (b >> 4) | (g & 3) << 6, dst[3] = 0x3 | (b & 0xF) << 4;
        }
}
#undef UNUSED3

/* ---- 16-bit colour-space converters ----------------------------------------------------------------------------- */
static int clampr(int v, int lo, int hi) { return v < lo ? lo : v > hi ? hi : v; }
static void y416_px(const unsigned char *src, int shift, int lo, int hi, int *r, int *g, int *b)
{
        const struct coeffs c = cfs709(16);
        uint16_t in[4];
        memcpy(in, src, 8);
        const int u = in[0] - 32768, y = c.y_scale * (in[1] - 4096), v = in[2] - 32768;
        *r = clampr((y + v * c.r_cr) >> shift, lo, hi);
        *g = clampr((y + u * c.g_cb + v * c.g_cr) >> shift, lo, hi);
        *b = clampr((y + u * c.b_cb) >> shift, lo, hi);
}
/* vc_copylineY416toRG48, pixfmt_conv.c:2485-2514 */
static void y416_to_rg48(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x < dst_len; x += 6, src += 8, dst += 6) {
                int r, g, b;
                y416_px(src, COMP_BASE, 256, 65279, &r, &g, &b);
                const uint16_t o[3] = { r, g, b };
                memcpy(dst, o, 6);
        }
}
/* vc_copylineY416toRGB, :1948-1976 */
static void y416_to_rgb(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x < dst_len; x += 3, src += 8) {
                int r, g, b;
                y416_px(src, COMP_BASE + 8, 1, 254, &r, &g, &b);
                *dst++ = r, *dst++ = g, *dst++ = b;
        }
}
/* vc_copylineY416toRGBA, :1978-2006 */
static void y416_to_rgba(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        const uint32_t amask = 0xFFFFFFFFU ^ (0xFFU << rs) ^ (0xFFU << gs) ^ (0xFFU << bs);
        for (int x = 0; x < dst_len; x += 4, src += 8) {
                int r, g, b;
                y416_px(src, COMP_BASE + 8, 1, 254, &r, &g, &b);
                wr32(dst + x, amask | (uint32_t) r << rs | (uint32_t) g << gs | (uint32_t) b << bs);
        }
}
/* vc_copylineY416toR10k, :1917-1946 */
static void y416_to_r10k(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x < dst_len; x += 4, src += 8) {
                int r, g, b;
                y416_px(src, COMP_BASE + 6, 4, 1019, &r, &g, &b);
                *dst++ = r >> 2, *dst++ = (r & 0x3) << 6 | g >> 4, *dst++ = (g & 0xF) << 4 | b >> 6, *dst++ = (b & 0x3F) << 2;
        }
}
/* vc_copylineY416toV210, :3004-3033 */
static void y416_to_v210(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x < dst_len / 16; ++x) {
                uint16_t s[24];
                memcpy(s, src + x * 48, 48);
                unsigned char *d = dst + x * 16;
#define AVG(a, b) ((uint32_t) (uint16_t) ((s[a] + s[b]) / 2) >> 6)
#define YY(a) ((uint32_t) s[a] >> 6)
                wr32(d, AVG(0, 4) | YY(1) << 10 | AVG(2, 6) << 20);
                wr32(d + 4, YY(5) | AVG(8, 12) << 10 | YY(9) << 20);
                wr32(d + 8, AVG(10, 14) | YY(13) << 10 | AVG(16, 20) << 20);
                wr32(d + 12, YY(17) | AVG(18, 22) << 10 | YY(21) << 20);
#undef AVG
#undef YY
        }
}
/* vc_copylineRG48toY416, :2451-2483 */
static void rg48_to_y416(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        const struct coeffs c = cfs709(16);
        for (int x = 0; x < dst_len; x += 8, src += 6, dst += 8) {
                uint16_t in[3];
                memcpy(in, src, 6);
                const int r = in[0], g = in[1], b = in[2];
                const uint16_t o[4] = { ((r * c.cb_r + g * c.cb_g + b * c.cb_b) >> COMP_BASE) + 32768, ((r * c.y_r + g * c.y_g + b * c.y_b) >> COMP_BASE) + 4096,
                                        ((r * c.cr_r + g * c.cr_g + b * c.cr_b) >> COMP_BASE) + 32768, 0xFFFFU };
                memcpy(dst, o, 8);
        }
}
/* vc_copylineRG48toY216, :2410-2449 */
static void rg48_to_y216(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        const struct coeffs c = cfs709(16);
        for (int x = 0; x < dst_len; x += 8, src += 12, dst += 8) {
                uint16_t in[6];
                memcpy(in, src, 12);
                int r = in[0], g = in[1], b = in[2];
                const int y0 = ((r * c.y_r + g * c.y_g + b * c.y_b) >> COMP_BASE) + 4096;
                int u = (r * c.cb_r + g * c.cb_g + b * c.cb_b) >> COMP_BASE, v = (r * c.cr_r + g * c.cr_g + b * c.cr_b) >> COMP_BASE;
                r = in[3], g = in[4], b = in[5];
                u = ((u + ((r * c.cb_r + g * c.cb_g + b * c.cb_b) >> COMP_BASE)) / 2) + 32768;
                const int y1 = ((r * c.y_r + g * c.y_g + b * c.y_b) >> COMP_BASE) + 4096;
                v = ((v + ((r * c.cr_r + g * c.cr_g + b * c.cr_b) >> COMP_BASE)) / 2) + 32768;
                const uint16_t o[4] = { y0, u, y1, v };
                memcpy(dst, o, 8);
        }
}
/* vc_copylineRG48toV210, :2354-2407 */
static void rg48_to_v210(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        const struct coeffs c = cfs709(10);
        enum { OFF = COMP_BASE + 6 };
        for (int x = 0; x <= dst_len - 16; x += 16, dst += 16) {
                uint32_t y[6], u[3], v[3];
                for (int p = 0; p < 3; ++p, src += 12) {
                        uint16_t in[6];
                        memcpy(in, src, 12);
                        int r = in[0], g = in[1], b = in[2];
                        y[2 * p] = ((r * c.y_r + g * c.y_g + b * c.y_b) >> OFF) + 64;
                        int uu = (r * c.cb_r + g * c.cb_g + b * c.cb_b) >> OFF, vv = (r * c.cr_r + g * c.cr_g + b * c.cr_b) >> OFF;
                        r = in[3], g = in[4], b = in[5];
                        y[2 * p + 1] = ((r * c.y_r + g * c.y_g + b * c.y_b) >> OFF) + 64;
                        uu += (r * c.cb_r + g * c.cb_g + b * c.cb_b) >> OFF, vv += (r * c.cr_r + g * c.cr_g + b * c.cr_b) >> OFF;
                        u[p] = uu / 2 + 512, v[p] = vv / 2 + 512;
                }
                wr32(dst, u[0] | y[0] << 10 | v[0] << 20);
                wr32(dst + 4, y[1] | u[1] << 10 | y[2] << 20);
                wr32(dst + 8, v[1] | y[3] << 10 | u[2] << 20);
                wr32(dst + 12, y[4] | v[2] << 10 | y[5] << 20);
        }
}
/* vc_copylineUYVYtoRG48, :1124-1130: copylineYUVtoRGB (:1065-1094) with rgb16 = 1 */
static void uyvy_to_rg48(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        const struct coeffs c = cfs709(8);
        for (int x = 0; x <= dst_len - 12; x += 12, src += 4) {
                const int y1 = c.y_scale * (src[1] - 16), y2 = c.y_scale * (src[3] - 16), u = src[0] - 128, v = src[2] - 128;
                const int vals[6] = { (y1 + v * c.r_cr) >> COMP_BASE, (y1 + u * c.g_cb + v * c.g_cr) >> COMP_BASE, (y1 + u * c.b_cb) >> COMP_BASE,
                                      (y2 + v * c.r_cr) >> COMP_BASE, (y2 + u * c.g_cb + v * c.g_cr) >> COMP_BASE, (y2 + u * c.b_cb) >> COMP_BASE };
                for (int k = 0; k < 6; ++k) {
                        *dst++ = 0, *dst++ = clamp255(vals[k]);
                }
        }
}
/* vc_copyliner10ktoY416, :294-329 */
static void r10k_to_y416(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        const struct coeffs c = cfs709(16);
        for (int x = 0; x < dst_len; x += 8, src += 4, dst += 8) {
                const int r = src[0] << 8 | (src[1] & 0xC0), g = (src[1] & 0x3F) << 10 | (src[2] & 0xF0) << 2, b = (src[2] & 0xF) << 12 | (src[3] & 0xFC) << 4;
                const uint16_t o[4] = { ((r * c.cb_r + g * c.cb_g + b * c.cb_b) >> COMP_BASE) + 32768, ((r * c.y_r + g * c.y_g + b * c.y_b) >> COMP_BASE) + 4096,
                                        ((r * c.cr_r + g * c.cr_g + b * c.cr_b) >> COMP_BASE) + 32768, 0xFFFFU };
                memcpy(dst, o, 8);
        }
}
/* vc_copylineR10ktoUYVY, :2318-2334 */
static void r10k_to_uyvy(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (const unsigned char *end = dst + dst_len; dst < end; dst += 4, src += 8) {
                unsigned char rgb[6];
                for (int k = 0; k < 2; ++k) {
                        rgb[3 * k] = src[4 * k], rgb[3 * k + 1] = src[4 * k + 1] << 2 | src[4 * k + 2] >> 6, rgb[3 * k + 2] = src[4 * k + 2] << 4 | src[4 * k + 3] >> 4;
                }
                rgb_to_uyvy(dst, rgb, 4, 0, 8, 16);
        }
}

/* ---- R12L: 8 px x 3 x 12 bit = 36 bytes, component k of a group at bit 12k (little endian) -------------------------------- */
static unsigned r12_get(const unsigned char *blk, int k)
{
        const int off = 12 * k;
        const uint32_t v = blk[off >> 3] | (uint32_t) blk[(off >> 3) + 1] << 8;
        return (v >> (off & 7)) & 0xfff;
}
static void r12_put(unsigned char *blk, int k, unsigned v)
{
        const int off = 12 * k;
        if (off & 7) {
                blk[off >> 3] |= (v & 0xf) << 4, blk[(off >> 3) + 1] = v >> 4;
        } else {
                blk[off >> 3] = v & 0xff, blk[(off >> 3) + 1] = (blk[(off >> 3) + 1] & 0xf0) | v >> 8;
        }
}
/* vc_copylineR12LtoRGB, pixfmt_conv.c:353-430 */
static void r12l_to_rgb(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x <= dst_len - 24; x += 24, src += 36) {
                for (int k = 0; k < 24; ++k) {
                        *dst++ = r12_get(src, k) >> 4;
                }
        }
}
/* vc_copylineR12L, :438-523 (the last, possibly partial group goes through a temporary: exactly dst_len bytes are written) */
static void r12l_to_rgba(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        const uint32_t amask = 0xFFFFFFFFU ^ (0xFFU << rs) ^ (0xFFU << gs) ^ (0xFFU << bs);
        for (int x = 0; x < dst_len; x += 32, src += 36) {
                unsigned char tmp[32];
                for (int i = 0; i < 8; ++i) {
                        wr32(tmp + 4 * i, amask | (r12_get(src, 3 * i) >> 4) << rs | (r12_get(src, 3 * i + 1) >> 4) << gs | (r12_get(src, 3 * i + 2) >> 4) << bs);
                }
                memcpy(dst + x, tmp, dst_len - x < 32 ? dst_len - x : 32);
        }
}
/* vc_copylineR12LtoRG48, :1371-1476 (partial last group through a temporary, like R12L -> RGBA) */
static void r12l_to_rg48(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x < dst_len; x += 48, src += 36) {
                unsigned char tmp[48];
                for (int k = 0; k < 24; ++k) {
                        const uint16_t v = r12_get(src, k) << 4;
                        memcpy(tmp + 2 * k, &v, 2);
                }
                memcpy(dst + x, tmp, dst_len - x < 48 ? dst_len - x : 48);
        }
}
/* vc_copylineR12LtoR10k, :1640-1699.  Not a clean R10k: the 4th byte keeps all of B[7:0] (so the two padding bits carry B[1:0]), and for
 * pixel 1 of each group its low nibble is R[3:0] instead of B[3:0] (:1661 reads src[4] where src[7] holds B) */
static void r12l_to_r10k(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x <= dst_len - 32; x += 32, src += 36) {
                for (int i = 0; i < 8; ++i) {
                        const unsigned r = r12_get(src, 3 * i), g = r12_get(src, 3 * i + 1), b = r12_get(src, 3 * i + 2);
                        *dst++ = r >> 4, *dst++ = (r & 0xC) << 4 | g >> 6, *dst++ = ((g >> 2) & 0xF) << 4 | b >> 8;
                        *dst++ = i == 1 ? (b & 0xF0) | (r & 0xF) : b & 0xFF;
                }
        }
}
/* vc_copylineR12LtoY416, :1478-1542 */
static void r12l_to_y416(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        const struct coeffs c = cfs709(16);
        for (int x = 0; x < dst_len; x += 64, src += 36) {
                for (int i = 0; i < 8; ++i, dst += 8) {
                        const int r = r12_get(src, 3 * i) << 4, g = r12_get(src, 3 * i + 1) << 4, b = r12_get(src, 3 * i + 2) << 4;
                        const uint16_t o[4] = { ((r * c.cb_r + g * c.cb_g + b * c.cb_b) >> COMP_BASE) + 32768, ((r * c.y_r + g * c.y_g + b * c.y_b) >> COMP_BASE) + 4096,
                                                ((r * c.cr_r + g * c.cr_g + b * c.cr_b) >> COMP_BASE) + 32768, 0xFFFFU };
                        memcpy(dst, o, 8);
                }
        }
}
/* vc_copylineR12LtoUYVY, :1544-1638 */
static void r12l_to_uyvy(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        const struct coeffs c = cfs709(8);
        for (int x = 0; x < dst_len; x += 16, src += 36) {
                for (int i = 0; i < 4; ++i) {
                        int r[2], g[2], b[2];
                        for (int k = 0; k < 2; ++k) {
                                r[k] = r12_get(src, 6 * i + 3 * k) << 4, g[k] = r12_get(src, 6 * i + 3 * k + 1) << 4, b[k] = r12_get(src, 6 * i + 3 * k + 2) << 4;
                        }
                        *dst++ = (((r[0] * c.cb_r + g[0] * c.cb_g + b[0] * c.cb_b) + (r[1] * c.cb_r + g[1] * c.cb_g + b[1] * c.cb_b)) >> (COMP_BASE + 9)) + 128;
                        *dst++ = ((r[0] * c.y_r + g[0] * c.y_g + b[0] * c.y_b) >> (COMP_BASE + 8)) + 16;
                        *dst++ = (((r[0] * c.cr_r + g[0] * c.cr_g + b[0] * c.cr_b) + (r[1] * c.cr_r + g[1] * c.cr_g + b[1] * c.cr_b)) >> (COMP_BASE + 9)) + 128;
                        *dst++ = ((r[1] * c.y_r + g[1] * c.y_g + b[1] * c.y_b) >> (COMP_BASE + 8)) + 16;
                }
        }
}
/* vc_copylineRGB_AtoR12L, :1258-1334 (pix = 3: RGB :1324, pix = 4: RGBA :1330); vc_copylineRG48toR12L, :1701-1826; vc_copylineY416toR12L, :1828-1915 */
static void x_to_r12l(unsigned char *dst, const unsigned char *src, int dst_len, int kind)
{
        const struct coeffs c = cfs709(16);
        for (int x = 0; kind == 3 ? x < dst_len : x <= dst_len - 36; x += 36, dst += 36) {
                memset(dst, 0, 36);
                for (int i = 0; i < 8; ++i) {
                        unsigned r, g, b;
                        if (kind == 0 || kind == 1) {
                                r = src[0] << 4, g = src[1] << 4, b = src[2] << 4;
                                src += kind == 0 ? 3 : 4;
                        } else if (kind == 2) {
                                uint16_t in[3];
                                memcpy(in, src, 6);
                                r = in[0] >> 4, g = in[1] >> 4, b = in[2] >> 4;
                                src += 6;
                        } else {
                                uint16_t in[4];
                                memcpy(in, src, 8);
                                const int u = in[0] - 32768, y = c.y_scale * (in[1] - 4096), v = in[2] - 32768;
                                r = clampr((y + v * c.r_cr) >> (COMP_BASE + 4), 16, 4079);
                                g = clampr((y + u * c.g_cb + v * c.g_cr) >> (COMP_BASE + 4), 16, 4079);
                                b = clampr((y + u * c.b_cb) >> (COMP_BASE + 4), 16, 4079);
                                src += 8;
                        }
                        r12_put(dst, 3 * i, r), r12_put(dst, 3 * i + 1, g), r12_put(dst, 3 * i + 2, b);
                }
        }
}
static void rgb_to_r12l(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; x_to_r12l(d, s, n, 0); }
static void rgba_to_r12l(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; x_to_r12l(d, s, n, 1); }
static void rg48_to_r12l(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; x_to_r12l(d, s, n, 2); }
static void y416_to_r12l(unsigned char *d, const unsigned char *s, int n, int rs, int gs, int bs) { (void) rs, (void) gs, (void) bs; x_to_r12l(d, s, n, 3); }

/* vc_copylineV210toRG48, pixfmt_conv.c:2942-3002 */
static void v210_to_rg48(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        const struct coeffs c = cfs709(10);
        for (int x = 0; x < dst_len; x += 36, src += 16) {
                const uint32_t w0 = rd32(src), w1 = rd32(src + 4), w2 = rd32(src + 8), w3 = rd32(src + 12);
                const int y[6] = { (w0 >> 10) & 0x3ff, w1 & 0x3ff, (w1 >> 20) & 0x3ff, (w2 >> 10) & 0x3ff, w3 & 0x3ff, (w3 >> 20) & 0x3ff };
                const int u[3] = { (int) (w0 & 0x3ff) - 512, (int) ((w1 >> 10) & 0x3ff) - 512, (int) ((w2 >> 20) & 0x3ff) - 512 };
                const int v[3] = { (int) ((w0 >> 20) & 0x3ff) - 512, (int) (w2 & 0x3ff) - 512, (int) ((w3 >> 10) & 0x3ff) - 512 };
                for (int i = 0; i < 6; ++i, dst += 6) {
                        const int ys = c.y_scale * (y[i] - 64);
                        const uint16_t o[3] = { clampr((ys + v[i / 2] * c.r_cr) >> (COMP_BASE - 6), 256, 65279),
                                                clampr((ys + u[i / 2] * c.g_cb + v[i / 2] * c.g_cr) >> (COMP_BASE - 6), 256, 65279),
                                                clampr((ys + u[i / 2] * c.b_cb) >> (COMP_BASE - 6), 256, 65279) };
                        memcpy(dst, o, 6);
                }
        }
}
/* vc_copylineDVS10 (the C variant that is compiled, pixfmt_conv.c:690-720): keeps bytes 0..2 of every 32-bit word */
static void dvs10_to_uyvy(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        const int src_len = dst_len / 1.5;
        for (int x = 0; x <= src_len - 16; x += 16, src += 32) {
                for (int k = 0; k < 8; ++k) {
                        *dst++ = src[4 * k], *dst++ = src[4 * k + 1], *dst++ = src[4 * k + 2];
                }
        }
}
/* vc_copylineDVS10toV210, :595-618 */
static void dvs10_to_v210(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs)
{
        (void) rs, (void) gs, (void) bs;
        for (int x = 0; x <= dst_len - 4; x += 4) {
                const uint32_t a = rd32(src + x);
                wr32(dst + x, (((a >> 24) * 0x00010101U) & 0x00300c03U) | ((a << 2) & (0xffU << 2)) | ((a << 4) & (0xff00U << 4)) | ((a << 6) & (0xff0000U << 6)));
        }
}

/* get_decoder_from_to, pixfmt_conv.c:3110-3125 (subset of decoders[] :3041-3103 restated so far) */
static line_fn *decoder_from_to(int in, int out)
{
        if (in == out && out != C_RGBA && out != C_RGB) {
                return copy_line;
        }
        switch (in * 256 + out) {
        case C_v210 * 256 + C_UYVY: return v210_to_uyvy;
        case C_YUYV * 256 + C_UYVY: case C_UYVY * 256 + C_YUYV: return yuyv_uyvy;
        case C_UYVY * 256 + C_RGB: return uyvy_to_rgb;
        case C_YUYV * 256 + C_RGB: return yuyv_to_rgb;
        case C_UYVY * 256 + C_RGBA: return uyvy_to_rgba;
        case C_RGB * 256 + C_UYVY: return rgb_to_uyvy;
        case C_BGR * 256 + C_UYVY: return bgr_to_uyvy;
        case C_RGBA * 256 + C_UYVY: return rgba_to_uyvy;
        case C_RG48 * 256 + C_UYVY: return rg48_to_uyvy;
        case C_RGB * 256 + C_RGBA: return rgb_to_rgba;
        case C_RGBA * 256 + C_RGB: return rgba_to_rgb;
        case C_RGBA * 256 + C_RGBA: return rgba_to_rgba;
        case C_RGB * 256 + C_RGB: return rgb_to_rgb;
        case C_BGR * 256 + C_RGB: return bgr_to_rgb;
        case C_UYVY * 256 + C_v210: return uyvy_to_v210;
        case C_Y216 * 256 + C_v210: return y216_to_v210;
        case C_v210 * 256 + C_Y216: return v210_to_y216;
        case C_v210 * 256 + C_Y416: return v210_to_y416;
        case C_v210 * 256 + C_RGB: return v210_to_rgb;
        case C_RG48 * 256 + C_RGB: return rg48_to_rgb;
        case C_RG48 * 256 + C_RGBA: return rg48_to_rgba;
        case C_RG48 * 256 + C_R10k: return rg48_to_r10k;
        case C_RGBA * 256 + C_RG48: return rgba_to_rg48;
        case C_RGB * 256 + C_RG48: return rgb_to_rg48;
        case C_UYVY * 256 + C_Y216: return uyvy_to_y216;
        case C_UYVY * 256 + C_Y416: return uyvy_to_y416;
        case C_Y216 * 256 + C_UYVY: return y216_to_uyvy;
        case C_Y416 * 256 + C_UYVY: return y416_to_uyvy;
        case C_VUYA * 256 + C_Y416: return vuya_to_y416;
        case C_VUYA * 256 + C_UYVY: return vuya_to_uyvy;
        case C_VUYA * 256 + C_RGB: return vuya_to_rgb;
        case C_RGBA * 256 + C_VUYA: return rgba_to_vuya;
        case C_R10k * 256 + C_RGBA: return r10k_to_rgba;
        case C_R10k * 256 + C_RGB: return r10k_to_rgb;
        case C_R10k * 256 + C_RG48: return r10k_to_rg48;
        case C_RGBA * 256 + C_R10k: return rgba_to_r10k;
        case C_Y416 * 256 + C_RG48: return y416_to_rg48;
        case C_Y416 * 256 + C_RGB: return y416_to_rgb;
        case C_Y416 * 256 + C_RGBA: return y416_to_rgba;
        case C_Y416 * 256 + C_R10k: return y416_to_r10k;
        case C_Y416 * 256 + C_v210: return y416_to_v210;
        case C_RG48 * 256 + C_Y416: return rg48_to_y416;
        case C_RG48 * 256 + C_Y216: return rg48_to_y216;
        case C_RG48 * 256 + C_v210: return rg48_to_v210;
        case C_UYVY * 256 + C_RG48: return uyvy_to_rg48;
        case C_R10k * 256 + C_Y416: return r10k_to_y416;
        case C_R10k * 256 + C_UYVY: return r10k_to_uyvy;
        case C_v210 * 256 + C_RG48: return v210_to_rg48;
        case C_DVS10 * 256 + C_UYVY: return dvs10_to_uyvy;
        case C_DVS10 * 256 + C_v210: return dvs10_to_v210;
        case C_R12L * 256 + C_RGB: return r12l_to_rgb;
        case C_R12L * 256 + C_RGBA: return r12l_to_rgba;
        case C_R12L * 256 + C_RG48: return r12l_to_rg48;
        case C_R12L * 256 + C_R10k: return r12l_to_r10k;
        case C_R12L * 256 + C_Y416: return r12l_to_y416;
        case C_R12L * 256 + C_UYVY: return r12l_to_uyvy;
        case C_RGB * 256 + C_R12L: return rgb_to_r12l;
        case C_RGBA * 256 + C_R12L: return rgba_to_r12l;
        case C_RG48 * 256 + C_R12L: return rg48_to_r12l;
        case C_Y416 * 256 + C_R12L: return y416_to_r12l;
        }
        return NULL;
}

API int orc_has_decoder(int in, int out) { return decoder_from_to(in, out) != NULL; }

/* row loop of tools/convert.cpp:148-152 */
API int orc_convert(int in_codec, int out_codec, unsigned char *dst, long dst_pitch, const unsigned char *src, long src_pitch,
                    int dst_len, int height, int rs, int gs, int bs)
{
        line_fn *f = decoder_from_to(in_codec, out_codec);
        if (f == NULL) {
                return -4;
        }
        for (int y = 0; y < height; ++y) {
                f(dst + y * dst_pitch, src + y * src_pitch, dst_len, rs, gs, bs);
        }
        return 0;
}

/* ---- src/to_planar.c ------------------------------------------------------------------------------ */
/* v210_to_p010le, to_planar.c:64-155 */
API void orc_v210_to_p010le(int width, int height, unsigned char *out_y, unsigned ls_y, unsigned char *out_c, unsigned ls_c,
                            const unsigned char *in)
{
        const long in_ls = orc_vc_get_linesize(width, C_v210);
        void *garbage = NULL;
        for (int y = 0; y < height; y += 2) {
                const uint32_t *src = (const uint32_t *) (const void *) (in + y * in_ls);
                const uint32_t *src2 = (const uint32_t *) (const void *) (in + (y + 1) * in_ls);
                uint16_t *dst_y = (uint16_t *) (void *) (out_y + (size_t) ls_y * y);
                uint16_t *dst_y2 = (uint16_t *) (void *) (out_y + (size_t) ls_y * (y + 1));
                uint16_t *dst_c = (uint16_t *) (void *) (out_c + (size_t) ls_c * y / 2);
                if (height - y == 1) { /* :84-87 */
                        dst_y2 = garbage = malloc(ls_y);
                        src2 = src;
                }
                int w = (width + 5) / 6 * 6; /* :89 */
                if (height - y == 1 || height - y == 2) {
                        w = width;
                }
                for (int x = 0; x < w / 6; ++x) {
                        const uint32_t a0 = *src++, a1 = *src++, a2 = *src++, a3 = *src++;
                        const uint32_t b0 = *src2++, b1 = *src2++, b2 = *src2++, b3 = *src2++;
#define S(w, sh) (((w) >> (sh)) & 0x3ff)
                        *dst_y++ = S(a0, 10) << 6, *dst_y++ = S(a1, 0) << 6, *dst_y++ = S(a1, 20) << 6;
                        *dst_y++ = S(a2, 10) << 6, *dst_y++ = S(a3, 0) << 6, *dst_y++ = S(a3, 20) << 6;
                        *dst_y2++ = S(b0, 10) << 6, *dst_y2++ = S(b1, 0) << 6, *dst_y2++ = S(b1, 20) << 6;
                        *dst_y2++ = S(b2, 10) << 6, *dst_y2++ = S(b3, 0) << 6, *dst_y2++ = S(b3, 20) << 6;
                        *dst_c++ = ((S(a0, 0) + S(b0, 0)) / 2) << 6;   /* Cb */
                        *dst_c++ = ((S(a0, 20) + S(b0, 20)) / 2) << 6; /* Cr */
                        *dst_c++ = ((S(a1, 10) + S(b1, 10)) / 2) << 6;
                        *dst_c++ = ((S(a2, 0) + S(b2, 0)) / 2) << 6;
                        *dst_c++ = ((S(a2, 20) + S(b2, 20)) / 2) << 6;
                        *dst_c++ = ((S(a3, 10) + S(b3, 10)) / 2) << 6;
#undef S
                }
                /* :141-151 — pointer arithmetic on uint16_t*: "- out_linesize" moves back out_linesize ELEMENTS,
                 * i.e. two rows.  Where that would read before the buffer (reference UB) nothing is copied. */
                if ((height - y == 1 || height - y == 2) && width % 6 != 0) {
                        const size_t pix_cnt = width % 6;
                        if (y >= 2) {
                                memcpy(dst_y, dst_y - ls_y, pix_cnt * 2);
                                if (height - y == 2) {
                                        memcpy(dst_y2, dst_y - ls_y, pix_cnt * 2);
                                }
                        }
                        if (y / 2 >= 2) {
                                memcpy(dst_c, dst_c - ls_c, pix_cnt * 2);
                        }
                }
        }
        free(garbage);
}

/* ---- line converters exported outside decoders[] (pixfmt_conv.h:93-101) -------------------------------------------------------------------
 * func: 1 vc_copylineABGRtoRGB (pixfmt_conv.c:809-843, SSSE3 build: the scalar tail never advances src), 2 vc_copylineBGRAtoRGB (:845-860),
 * 3 vc_copylineToRGBA_inplace (:907-921), 4 vc_copylineUYVYtoGrayscale (:927-938).  Same whole-buffer row loop as orc_convert. */
static void x32_to_rgb(unsigned char *dst, const unsigned char *src, int dst_len, int rs, int gs, int bs, int quirk)
{
        const int tail_px = !quirk ? 0x7fffffff : dst_len >= 24 ? ((dst_len - 24) / 12 + 1) * 4 : 0;
        for (int x = 0, px = 0; x <= dst_len - 3; x += 3, ++px) {
                const uint32_t in = rd32(src + 4 * (size_t) (px < tail_px ? px : tail_px));
                dst[x] = (in >> rs) & 0xff, dst[x + 1] = (in >> gs) & 0xff, dst[x + 2] = (in >> bs) & 0xff;
        }
}
API int orc_copyline_named(int func, unsigned char *dst, long dst_pitch, const unsigned char *src, long src_pitch, int dst_len, int height, int rs, int gs, int bs)
{
        for (int y = 0; y < height; ++y) {
                unsigned char *d = dst + (size_t) y * dst_pitch;
                const unsigned char *s = src + (size_t) y * src_pitch;
                switch (func) {
                case 1: x32_to_rgb(d, s, dst_len, 24, 16, 8, 1); break;
                case 2: x32_to_rgb(d, s, dst_len, 16, 8, 0, 0); break;
                case 3:
                        for (int x = 0; x + 4 <= dst_len; x += 4) {
                                const uint32_t in = rd32(s + x);
                                wr32(d + x, ((in >> rs) & 0xff) | ((in >> gs) & 0xff) << 8 | ((in >> bs) & 0xff) << 16);
                        }
                        break;
                case 4:
                        for (int x = 0; x <= dst_len - 2; x += 2) {
                                d[x] = s[2 * x + 1], d[x + 1] = s[2 * x + 3];
                        }
                        break;
                default: return -4;
                }
        }
        return 0;
}

/* ---- src/utils/cuda_pix_conv.cu (device-to-device helpers of the reference), restated on the CPU ---------------------------------------------
 * kind 0 RGB->RGBA (:7-29), 1 RGBA->RGB (:32-54), 2 UYVY->RGBA (:60-92), 3 RGBA->UYVY (:95-133).  The float matrix of kind 2 follows the FMA
 * contraction of the reference's sm_100a build (read from its SASS): y = 1.164f * (Y - 16); R = fma(v, 1.793f, y);
 * G = fma(v, -0.534f, y) - 0.213f * u; B = fma(u, 2.115f, y); x > 0 ? (x < 255 ? trunc : 255) : 0.  Pinned on the GPU against the unmodified file. */
#include <math.h>
static uint32_t sat_trunc(float x) { return x > 0.0f ? (x < 255.0f ? (uint32_t) (int) x : 255u) : 0u; }
API void orc_cuda_pix_conv(int kind, unsigned char *dst, size_t dpitch, const unsigned char *src, size_t spitch, int width, int height)
{
        for (int y = 0; y < height; ++y) {
                const unsigned char *s = src + (size_t) y * spitch;
                unsigned char *d = dst + (size_t) y * dpitch;
                for (int x = 0; x < width; ++x) {
                        if (kind == 0) {
                                d[4 * x] = s[3 * x], d[4 * x + 1] = s[3 * x + 1], d[4 * x + 2] = s[3 * x + 2], d[4 * x + 3] = 0;
                        } else if (kind == 1) {
                                d[3 * x] = s[4 * x], d[3 * x + 1] = s[4 * x + 1], d[3 * x + 2] = s[4 * x + 2];
                        } else if (kind == 2) {
                                const unsigned char *b = s + 4 * (x / 2);
                                const float u = (float) (b[0] - 128), v = (float) (b[2] - 128), yy = (float) (b[1 + 2 * (x & 1)] - 16) * 1.164f;
                                d[4 * x] = sat_trunc(fmaf(v, 1.793f, yy)), d[4 * x + 1] = sat_trunc(fmaf(v, -0.534f, yy) - u * 0.213f), d[4 * x + 2] = sat_trunc(fmaf(u, 2.115f, yy));
                                d[4 * x + 3] = 0;
                        } else if (x % 2 == 0 && x + 1 < width) {
                                const unsigned char *p = s + 4 * x;
                                const int y1 = 11993 * p[0] + 40239 * p[1] + 4063 * p[2] + (1 << 20), y2 = 11993 * p[4] + 40239 * p[5] + 4063 * p[6] + (1 << 20);
                                int u = (-6619 * p[0] - 22151 * p[1] + 28770 * p[2]) + (-6619 * p[4] - 22151 * p[5] + 28770 * p[6]);
                                int v = (28770 * p[0] - 26149 * p[1] - 2621 * p[2]) + (28770 * p[4] - 26149 * p[5] - 2621 * p[6]);
                                u = u / 2 + (1 << 23), v = v / 2 + (1 << 23);
                                const int lim = (1 << 24) - 1;
                                d[2 * x] = clampr(u, 0, lim) >> 16, d[2 * x + 1] = clampr(y1, 0, lim) >> 16, d[2 * x + 2] = clampr(v, 0, lim) >> 16, d[2 * x + 3] = clampr(y2, 0, lim) >> 16;
                        }
                }
        }
}
