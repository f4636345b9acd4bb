/* TEST INFRASTRUCTURE - not part of the product.
 * CPU restatement of UltraGrid's packed<->planar whole-buffer converters (src/to_planar.c, src/from_planar.c) except
 * v210_to_p010le (pixfmt_oracle.c).  Index-based loops, one function per reference routine; each cites the lines it follows.
 * Pinned against the unmodified reference objects (oracle/_ref/libugref.so) by tests/test_planar.py.
 * Same by-value argument structs as the reference (to_planar.h:53-59, from_planar.h:58-70). */
#include <stdint.h>
#include <string.h>

#define API __attribute__((visibility("default")))

struct to_planar_data {
        int width, height;
        unsigned char *out_data[4];
        unsigned out_linesize[4];
        const unsigned char *in_data;
};
struct from_planar_data {
        int width, height;
        unsigned char *out_data;
        unsigned out_pitch;
        const unsigned char *in_data[4];
        unsigned in_linesize[4];
        int in_depth, log2_chroma_h, rgb_shift[3];
};

static uint16_t rd16(const unsigned char *p) { uint16_t v; memcpy(&v, p, 2); return v; }
static void wr16(unsigned char *p, unsigned v) { const uint16_t t = (uint16_t) v; memcpy(p, &t, 2); }
static void wr32(unsigned char *p, uint32_t v) { memcpy(p, &v, 4); }

/* ---- to_planar ------------------------------------------------------------------------------------------------------ */

/* y216_to_p010le, to_planar.c:164-200.  The odd row of a pair is written directly behind the even row's `width` luma samples
 * (the reference keeps incrementing one pointer, :187-197), not at the next out_linesize[0]. */
API void orc_y216_to_p010le(struct to_planar_data d)
{
        const size_t in_ls = (size_t) ((d.width + 1) / 2) * 8;
        const int cw = (d.width + 1) & ~1;  /* an odd last pixel still emits Cb and Cr */
        for (int y = 0; y < d.height; ++y) {
                const unsigned char *s = d.in_data + y * in_ls;
                unsigned char *oy = d.out_data[0] + (size_t) (y & ~1) * d.out_linesize[0] + (y & 1 ? 2 * (size_t) d.width : 0);
                for (int x = 0; x < d.width; ++x) {
                        wr16(oy + 2 * x, rd16(s + 4 * x));
                }
                if (y % 2 == 0) {
                        unsigned char *oc = d.out_data[1] + (size_t) (y / 2) * d.out_linesize[1];
                        for (int x = 0; x < cw; ++x) {
                                wr16(oc + 2 * x, rd16(s + 4 * x + 2));
                        }
                }
        }
}

/* uyvy_to_nv12 (to_planar.c:207-302; built with SSE3: x < 16 * (width / 16) averages with round-half-up, the rest truncates) and
 * uyvy_to_i420 (:343-378, (a + b + 1) / 2 everywhere) */
static void uyvy_to_420(struct to_planar_data d, int i420)
{
        const size_t in_ls = i420 ? (size_t) ((d.width + 1) / 2) * 4 : (size_t) d.width * 2;
        const int sse_end = d.width / 16 * 16;
        for (int y = 0; y < d.height; y += 2) {
                const int y2 = y + 1 < d.height ? y + 1 : y;
                const unsigned char *a = d.in_data + y * in_ls, *b = d.in_data + y2 * in_ls;
                unsigned char *oy = d.out_data[0] + (size_t) y * d.out_linesize[0], *oy2 = d.out_data[0] + (size_t) y2 * d.out_linesize[0];
                for (int x = 0; x < d.width; x += 2) {
                        const int r = (i420 || x < sse_end) ? 1 : 0;
                        const unsigned char cb = (a[2 * x] + b[2 * x] + r) / 2, cr = (a[2 * x + 2] + b[2 * x + 2] + r) / 2;
                        if (i420) {
                                d.out_data[1][(size_t) (y / 2) * d.out_linesize[1] + x / 2] = cb;
                                d.out_data[2][(size_t) (y / 2) * d.out_linesize[2] + x / 2] = cr;
                        } else {
                                d.out_data[1][(size_t) (y / 2) * d.out_linesize[1] + x] = cb;
                                d.out_data[1][(size_t) (y / 2) * d.out_linesize[1] + x + 1] = cr;
                        }
                        oy[x] = a[2 * x + 1], oy2[x] = b[2 * x + 1];
                        if (x + 1 < d.width) {
                                oy[x + 1] = a[2 * x + 3], oy2[x + 1] = b[2 * x + 3];
                        }
                }
        }
}
API void orc_uyvy_to_nv12(struct to_planar_data d) { uyvy_to_420(d, 0); }
API void orc_uyvy_to_i420(struct to_planar_data d) { uyvy_to_420(d, 1); }

/* rgba_to_bgra, to_planar.c:304-319 */
API void orc_rgba_to_bgra(struct to_planar_data d)
{
        for (int y = 0; y < d.height; ++y) {
                const unsigned char *s = d.in_data + (size_t) y * d.width * 4;
                unsigned char *o = d.out_data[0] + (size_t) y * d.out_linesize[0];
                for (int x = 0; x < d.width; ++x) {
                        o[4 * x] = s[4 * x + 2], o[4 * x + 1] = s[4 * x + 1], o[4 * x + 2] = s[4 * x], o[4 * x + 3] = s[4 * x + 3];
                }
        }
}

/* vuya_to_i444, to_planar.c:321-337 */
API void orc_vuya_to_i444(struct to_planar_data d)
{
        for (int y = 0; y < d.height; ++y) {
                const unsigned char *s = d.in_data + (size_t) y * d.width * 4;
                for (int x = 0; x < d.width; ++x) {
                        d.out_data[2][(size_t) y * d.out_linesize[2] + x] = s[4 * x];
                        d.out_data[1][(size_t) y * d.out_linesize[1] + x] = s[4 * x + 1];
                        d.out_data[0][(size_t) y * d.out_linesize[0] + x] = s[4 * x + 2];
                }
        }
}

static unsigned r12_get(const unsigned char *blk, int k)
{
        const int off = 12 * k;
        return ((blk[off >> 3] | (unsigned) blk[(off >> 3) + 1] << 8) >> (off & 7)) & 0xfff;
}

/* r12l_to_gbrpXXle, to_planar.c:381-481: whole 8-pixel groups (up to 7 samples past width; in row order, so only the last
 * row's spill survives) */
static void r12l_to_planes(struct to_planar_data d, int depth, int rind, int gind, int bind)
{
        const size_t in_ls = (size_t) ((d.width + 7) / 8) * 36;
        const int ind[3] = { rind, gind, bind };
        for (int y = 0; y < d.height; ++y) {
                for (int x = 0; x < d.width; x += 8) {
                        const unsigned char *blk = d.in_data + y * in_ls + (size_t) (x / 8) * 36;
                        for (int i = 0; i < 8; ++i) {
                                for (int c = 0; c < 3; ++c) {
                                        wr16(d.out_data[ind[c]] + (size_t) y * d.out_linesize[ind[c]] + 2 * (x + i), r12_get(blk, 3 * i + c) << (depth - 12));
                                }
                        }
                }
        }
}
API void orc_r12l_to_gbrp12le(struct to_planar_data d) { r12l_to_planes(d, 12, 2, 0, 1); }
API void orc_r12l_to_gbrp16le(struct to_planar_data d) { r12l_to_planes(d, 16, 2, 0, 1); }
API void orc_r12l_to_rgbp12le(struct to_planar_data d) { r12l_to_planes(d, 12, 0, 1, 2); }

/* ---- from_planar ---------------------------------------------------------------------------------------------------- */
static unsigned s16(const struct from_planar_data *d, int plane, int y, int x)
{
        return rd16(d->in_data[plane] + (size_t) d->in_linesize[plane] * y + 2 * x);
}
static unsigned s8(const struct from_planar_data *d, int plane, int y, int x) { return d->in_data[plane][(size_t) d->in_linesize[plane] * y + x]; }

/* gbrpXXle_to_r12l, from_planar.c:61-129.  Samples beyond width in the last group: zero here (uninitialised in the reference). */
static void planes_to_r12l(struct from_planar_data d, int depth, int rind, int gind, int bind)
{
        const int ind[3] = { rind, gind, bind };
        for (int y = 0; y < d.height; ++y) {
                for (int x = 0; x < d.width; x += 8) {
                        unsigned f[24];
                        for (int i = 0; i < 8; ++i) {
                                for (int c = 0; c < 3; ++c) {
                                        f[3 * i + c] = x + i < d.width ? s16(&d, ind[c], y, x + i) >> (depth - 12) : 0;
                                }
                        }
                        unsigned char *o = d.out_data + (size_t) y * d.out_pitch + (size_t) (x / 8) * 36;
                        for (int p = 0; p < 12; ++p) {  /* two fields = three bytes, each truncated by the uint8 store */
                                const unsigned e = f[2 * p], od = f[2 * p + 1];
                                o[3 * p] = e & 0xff, o[3 * p + 1] = (od & 0xf) << 4 | e >> 8, o[3 * p + 2] = od >> 4;
                        }
                }
        }
}
API void orc_gbrp12le_to_r12l(struct from_planar_data d) { planes_to_r12l(d, 12, 2, 0, 1); }
API void orc_gbrp16le_to_r12l(struct from_planar_data d) { planes_to_r12l(d, 16, 2, 0, 1); }
API void orc_rgbpXXle_to_r12l(struct from_planar_data d) { planes_to_r12l(d, d.in_depth, 0, 1, 2); }

/* rgbpXXle_to_rg48_int, from_planar.c:157-177 */
static void planes_to_rg48(struct from_planar_data d, int depth, int rind, int gind, int bind)
{
        for (int y = 0; y < d.height; ++y) {
                unsigned char *o = d.out_data + (size_t) y * d.out_pitch;
                for (int x = 0; x < d.width; ++x) {
                        wr16(o + 6 * x, s16(&d, rind, y, x) << (16 - depth));
                        wr16(o + 6 * x + 2, s16(&d, gind, y, x) << (16 - depth));
                        wr16(o + 6 * x + 4, s16(&d, bind, y, x) << (16 - depth));
                }
        }
}
API void orc_gbrp10le_to_rg48(struct from_planar_data d) { planes_to_rg48(d, 10, 2, 0, 1); }
API void orc_gbrp12le_to_rg48(struct from_planar_data d) { planes_to_rg48(d, 12, 2, 0, 1); }
API void orc_gbrp16le_to_rg48(struct from_planar_data d) { planes_to_rg48(d, 16, 2, 0, 1); }
API void orc_rgbpXXle_to_rg48(struct from_planar_data d) { planes_to_rg48(d, d.in_depth, 0, 1, 2); }

/* gbrpXXle_to_r10k, from_planar.c:203-226 */
static void planes_to_r10k(struct from_planar_data d, int depth, int rind, int gind, int bind)
{
        for (int y = 0; y < d.height; ++y) {
                unsigned char *o = d.out_data + (size_t) y * d.out_pitch;
                for (int x = 0; x < d.width; ++x) {
                        const unsigned r = s16(&d, rind, y, x), g = s16(&d, gind, y, x), b = s16(&d, bind, y, x);
                        o[4 * x] = r >> (depth - 8);
                        o[4 * x + 1] = ((r >> (depth - 10)) & 0x3) << 6 | g >> (depth - 6);
                        o[4 * x + 2] = ((g >> (depth - 10)) & 0xf) << 4 | b >> (depth - 4);
                        o[4 * x + 3] = ((b >> (depth - 10)) & 0x3f) << 2 | 0x3;
                }
        }
}
API void orc_gbrp10le_to_r10k(struct from_planar_data d) { planes_to_r10k(d, 10, 2, 0, 1); }
API void orc_gbrp12le_to_r10k(struct from_planar_data d) { planes_to_r10k(d, 12, 2, 0, 1); }
API void orc_gbrp16le_to_r10k(struct from_planar_data d) { planes_to_r10k(d, 16, 2, 0, 1); }
API void orc_rgbpXXle_to_r10k(struct from_planar_data d) { planes_to_r10k(d, d.in_depth, 0, 1, 2); }

/* yuv422p10le_to_v210, from_planar.c:295-333 */
API void orc_yuv422p10le_to_v210(struct from_planar_data d)
{
        for (int y = 0; y < d.height; ++y) {
                for (int g = 0; g < d.width / 6; ++g) {
                        uint32_t Y[6], B[3], R[3];
                        for (int i = 0; i < 6; ++i) {
                                Y[i] = s16(&d, 0, y, 6 * g + i);
                        }
                        for (int i = 0; i < 3; ++i) {
                                B[i] = s16(&d, 1, y, 3 * g + i), R[i] = s16(&d, 2, y, 3 * g + i);
                        }
                        unsigned char *o = d.out_data + (size_t) y * d.out_pitch + 16 * (size_t) g;
                        wr32(o, B[0] | Y[0] << 10 | R[0] << 20), wr32(o + 4, Y[1] | B[1] << 10 | Y[2] << 20);
                        wr32(o + 8, R[1] | Y[3] << 10 | B[2] << 20), wr32(o + 12, Y[4] | R[2] << 10 | Y[5] << 20);
                }
        }
}

/* gbrap_to_rgb_rgba, from_planar.c:335-354 (all planes indexed with in_linesize[0]) */
static void gbrap_to_packed(struct from_planar_data d, int rind, int gind, int bind, int aind)
{
        const int n = aind < 0 ? 3 : 4;
        for (int y = 0; y < d.height; ++y) {
                for (int x = 0; x < d.width; ++x) {
                        unsigned char *o = d.out_data + (size_t) y * d.out_pitch + (size_t) n * x;
                        const size_t si = (size_t) y * d.in_linesize[0] + x;
                        o[0] = d.in_data[rind][si], o[1] = d.in_data[gind][si], o[2] = d.in_data[bind][si];
                        if (n == 4) {
                                o[3] = d.in_data[aind][si];
                        }
                }
        }
}
API void orc_gbrap_to_rgba(struct from_planar_data d) { gbrap_to_packed(d, 2, 0, 1, 3); }
API void orc_gbrap_to_rgb(struct from_planar_data d) { gbrap_to_packed(d, 2, 0, 1, -1); }

/* yuv420_to_i420, from_planar.c:368-390 */
API void orc_yuv420_to_i420(struct from_planar_data d)
{
        const size_t w = d.width, h = d.height;
        unsigned char *oy = d.out_data, *ou = oy + w * h, *ov = ou + (w / 2) * (h / 2);
        for (size_t y = 0; y < h; ++y) {
                memcpy(oy + y * w, d.in_data[0] + y * d.in_linesize[0], w);
        }
        for (size_t y = 0; y < h / 2; ++y) {
                memcpy(ou + y * (w / 2), d.in_data[1] + y * d.in_linesize[1], w / 2);
                memcpy(ov + y * (w / 2), d.in_data[2] + y * d.in_linesize[2], w / 2);
        }
}

/* yuv422p_to_uyvy_yuyv, from_planar.c:392-417; yuv422pXXle_to_uyvy_int, :425-441 */
static void yuv422p_to_packed(struct from_planar_data d, int yuyv, int depth)
{
        for (int y = 0; y < d.height; ++y) {
                unsigned char *o = d.out_data + (size_t) y * d.out_pitch;
                for (int j = 0; j < d.width / 2; ++j) {
                        unsigned char y0, y1, cb, cr;
                        if (depth == 8) {
                                y0 = s8(&d, 0, y, 2 * j), y1 = s8(&d, 0, y, 2 * j + 1), cb = s8(&d, 1, y, j), cr = s8(&d, 2, y, j);
                        } else {
                                y0 = s16(&d, 0, y, 2 * j) >> (depth - 8), y1 = s16(&d, 0, y, 2 * j + 1) >> (depth - 8);
                                cb = s16(&d, 1, y, j) >> (depth - 8), cr = s16(&d, 2, y, j) >> (depth - 8);
                        }
                        if (yuyv) {
                                o[4 * j] = y0, o[4 * j + 1] = cb, o[4 * j + 2] = y1, o[4 * j + 3] = cr;
                        } else {
                                o[4 * j] = cb, o[4 * j + 1] = y0, o[4 * j + 2] = cr, o[4 * j + 3] = y1;
                        }
                }
        }
}
API void orc_yuv422p_to_uyvy(struct from_planar_data d) { yuv422p_to_packed(d, 0, 8); }
API void orc_yuv422p_to_yuyv(struct from_planar_data d) { yuv422p_to_packed(d, 1, 8); }
API void orc_yuv422p10le_to_uyvy(struct from_planar_data d) { yuv422p_to_packed(d, 0, 10); }
API void orc_yuv422pXX_to_uyvy(struct from_planar_data d) { yuv422p_to_packed(d, 0, d.in_depth); }

/* gbrpXXle_to_rgb, from_planar.c:465-484 */
static void planes_to_rgb(struct from_planar_data d, int depth, int rind, int gind, int bind)
{
        for (int y = 0; y < d.height; ++y) {
                unsigned char *o = d.out_data + (size_t) y * d.out_pitch;
                for (int x = 0; x < d.width; ++x) {
                        o[3 * x] = s16(&d, rind, y, x) >> (depth - 8), o[3 * x + 1] = s16(&d, gind, y, x) >> (depth - 8), o[3 * x + 2] = s16(&d, bind, y, x) >> (depth - 8);
                }
        }
}
API void orc_gbrp10le_to_rgb(struct from_planar_data d) { planes_to_rgb(d, 10, 2, 0, 1); }
API void orc_gbrp12le_to_rgb(struct from_planar_data d) { planes_to_rgb(d, 12, 2, 0, 1); }
API void orc_gbrp16le_to_rgb(struct from_planar_data d) { planes_to_rgb(d, 16, 2, 0, 1); }
API void orc_rgbpXX_to_rgb(struct from_planar_data d)  /* :555-563 */
{
        if (d.in_depth == 8) {
                gbrap_to_packed(d, 0, 1, 2, -1);
        } else {
                planes_to_rgb(d, d.in_depth, 0, 1, 2);
        }
}

/* gbrpXXle_to_rgba, from_planar.c:486-517 (planes G, B, R) */
static void planes_to_rgba(struct from_planar_data d, int depth)
{
        const uint32_t amask = 0xFFFFFFFFU ^ (0xFFU << d.rgb_shift[0]) ^ (0xFFU << d.rgb_shift[1]) ^ (0xFFU << d.rgb_shift[2]);
        for (int y = 0; y < d.height; ++y) {
                for (int x = 0; x < d.width; ++x) {
                        wr32(d.out_data + (size_t) y * d.out_pitch + 4 * (size_t) x, amask | (s16(&d, 2, y, x) >> (depth - 8)) << d.rgb_shift[0] |
                                                                                           (s16(&d, 0, y, x) >> (depth - 8)) << d.rgb_shift[1] |
                                                                                           (s16(&d, 1, y, x) >> (depth - 8)) << d.rgb_shift[2]);
                }
        }
}
API void orc_gbrp10le_to_rgba(struct from_planar_data d) { planes_to_rgba(d, 10); }
API void orc_gbrp12le_to_rgba(struct from_planar_data d) { planes_to_rgba(d, 12); }
API void orc_gbrp16le_to_rgba(struct from_planar_data d) { planes_to_rgba(d, 16); }

/* yuv444p_to_vuya, from_planar.c:565-580 */
API void orc_yuv444p_to_vuya(struct from_planar_data d)
{
        for (int y = 0; y < d.height; ++y) {
                unsigned char *o = d.out_data + (size_t) y * d.out_pitch;
                for (int x = 0; x < d.width; ++x) {
                        o[4 * x] = s8(&d, 2, y, x), o[4 * x + 1] = s8(&d, 1, y, x), o[4 * x + 2] = s8(&d, 0, y, x), o[4 * x + 3] = 0xFF;
                }
        }
}

/* yuv420p_to_uyvy, from_planar.c:582-683 */
API void orc_yuv420p_to_uyvy(struct from_planar_data d)
{
        for (int y = 0; y < d.height; ++y) {
                unsigned char *o = d.out_data + (size_t) y * d.out_pitch;
                for (int x = 0; x < d.width; x += 2) {
                        o[2 * x] = s8(&d, 1, y / 2, x / 2), o[2 * x + 1] = s8(&d, 0, y, x), o[2 * x + 2] = s8(&d, 2, y / 2, x / 2);
                        o[2 * x + 3] = x + 1 < d.width ? s8(&d, 0, y, x + 1) : 0;
                }
        }
}
