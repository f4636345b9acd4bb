/* TEST INFRASTRUCTURE — CPU restatement of the baseline-JPEG encode stage (the work GPUJPEG does for UltraGrid's
 * src/video_compress/gpujpeg.cpp:617-630 as configured at :256-369).
 *
 * PARITY UNPINNED against GPUJPEG itself: libgpujpeg is an un-vendored external dependency (pkg-config
 * libgpujpeg >= 0.14, configure.ac:2634-2635; ext-deps/bootstrap_gpujpeg.sh:80 clones HEAD) and is absent here,
 * and the reference's only test at this boundary (test/gpujpeg_test.cpp:68-106) pins nothing about the bitstream.
 * What is pinned instead: (1) the stream is a valid ITU-T T.81 baseline JPEG that an independent decoder (libjpeg via
 * PIL) reads back; (2) PSNR at a given quality is within 0.3 dB of libjpeg's own encoder on the same image and tables;
 * (3) the reference's flat-grey round-trip test (max |diff| <= 1) passes; (4) the CUDA encoder produces the identical
 * byte stream (this file and the kernels use the same explicitly ordered float operations).
 *
 * Stream layout follows what the reference configures: UYVY input -> YCbCr stored as-is (no colour transform,
 * gpujpeg.cpp:304-305), 4:2:2, interleaved scan; RGB input -> stored as RGB, 4:4:4, one scan per component
 * (gpujpeg.cpp:303); restart intervals; Annex K tables scaled by IJG quality; header order as the reference's own
 * RFC 2435 writer (src/utils/jpeg_writer.c:215-382): SOI, APPn, DQT, SOF0, DHT x4, DRI, SOS.
 */
#include <math.h>
#include <omp.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../ultragrid_b200/csrc/jpeg_tables.h" /* ITU-T T.81 Annex K constants only (shared data, no code path) */

#define API __attribute__((visibility("default")))

enum { FMT_UYVY_422 = 0, FMT_RGB_444 = 1 };

/* 8-point AAN forward DCT (Arai/Agui/Nakajima), explicit operation order; all multiply-adds are fused on
 * purpose (fmaf) so that CPU and GPU round identically */
static void fdct8(float *d, int stride)
{
        const float t0 = d[0 * stride] + d[7 * stride], t7 = d[0 * stride] - d[7 * stride];
        const float t1 = d[1 * stride] + d[6 * stride], t6 = d[1 * stride] - d[6 * stride];
        const float t2 = d[2 * stride] + d[5 * stride], t5 = d[2 * stride] - d[5 * stride];
        const float t3 = d[3 * stride] + d[4 * stride], t4 = d[3 * stride] - d[4 * stride];
        const float e0 = t0 + t3, e3 = t0 - t3, e1 = t1 + t2, e2 = t1 - t2;
        d[0 * stride] = e0 + e1;
        d[4 * stride] = e0 - e1;
        const float s1 = e2 + e3; /* every multiply-add of this transform is ONE fused operation, written out: no product ever feeds a */
        d[2 * stride] = fmaf(s1, 0.707106781f, e3);  /* separate add, so there is nothing a compiler could contract differently     */
        d[6 * stride] = fmaf(s1, -0.707106781f, e3);
        const float o0 = t4 + t5, o1 = t5 + t6, o2 = t6 + t7;
        const float z5 = (o0 - o2) * 0.382683433f;
        const float z2 = fmaf(0.541196100f, o0, z5);
        const float z4 = fmaf(1.306562965f, o2, z5);
        const float z11 = fmaf(o1, 0.707106781f, t7), z13 = fmaf(o1, -0.707106781f, t7);
        d[5 * stride] = z13 + z2;
        d[3 * stride] = z13 - z2;
        d[1 * stride] = z11 + z4;
        d[7 * stride] = z11 - z4;
}

/* level shift, 2-D DCT, quantise (round half to even), zig-zag.  AC clamped to +-1023 (10-bit category limit). */
static void block_to_coeffs(const uint8_t px[64], const float qmul[64], int16_t zz[64])
{
        float f[64];
        for (int i = 0; i < 64; ++i) {
                f[i] = (float) ((int) px[i] - 128);
        }
        for (int r = 0; r < 8; ++r) {
                fdct8(f + 8 * r, 1);
        }
        for (int c = 0; c < 8; ++c) {
                fdct8(f + c, 8);
        }
        for (int k = 0; k < 64; ++k) {
                const int n = ugb_jpeg_zigzag[k];
                int v = (int) (fmaf(f[n], qmul[n], 12582912.0f) - 12582912.0f); /* round-to-nearest-even of the exact product (1.5 * 2^23 trick) */
                if (k > 0) {
                        v = v < -1023 ? -1023 : v > 1023 ? 1023 : v;
                }
                zz[k] = (int16_t) v;
        }
}

struct bitw {
        uint8_t *p;
        uint64_t acc;
        int nbits;
};
static void put_bits(struct bitw *w, uint32_t code, int len)
{
        w->acc = (w->acc << len) | (code & ((1u << len) - 1));
        w->nbits += len;
        while (w->nbits >= 8) {
                const uint8_t b = (uint8_t) (w->acc >> (w->nbits - 8));
                *w->p++ = b;
                if (b == 0xFF) {
                        *w->p++ = 0; /* byte stuffing, T.81 B.1.1.5 */
                }
                w->nbits -= 8;
        }
}
static void flush_bits(struct bitw *w)
{
        if (w->nbits > 0) {
                put_bits(w, 0x7F, 8 - w->nbits); /* pad with 1-bits, T.81 F.1.2.3 */
        }
        w->acc = 0, w->nbits = 0;
}
static int category(int v)
{
        int a = v < 0 ? -v : v, n = 0;
        while (a) {
                ++n, a >>= 1;
        }
        return n;
}

struct huff {
        uint16_t code[256];
        uint8_t len[256];
};

/* T.81 F.1.2: DC difference + AC run-lengths for one block */
static void encode_block(struct bitw *w, const int16_t zz[64], int *pred, const struct huff *dc, const struct huff *ac)
{
        const int diff = zz[0] - *pred;
        *pred = zz[0];
        int s = category(diff);
        put_bits(w, dc->code[s], dc->len[s]);
        if (s) {
                put_bits(w, (uint32_t) (diff < 0 ? diff - 1 : diff), s);
        }
        int run = 0;
        for (int k = 1; k < 64; ++k) {
                const int v = zz[k];
                if (v == 0) {
                        ++run;
                        continue;
                }
                while (run > 15) {
                        put_bits(w, ac->code[0xF0], ac->len[0xF0]); /* ZRL */
                        run -= 16;
                }
                s = category(v);
                put_bits(w, ac->code[(run << 4) | s], ac->len[(run << 4) | s]);
                put_bits(w, (uint32_t) (v < 0 ? v - 1 : v), s);
                run = 0;
        }
        if (run > 0) {
                put_bits(w, ac->code[0x00], ac->len[0x00]); /* EOB */
        }
}

static uint8_t *put16(uint8_t *p, unsigned v)
{
        *p++ = (uint8_t) (v >> 8), *p++ = (uint8_t) v;
        return p;
}
static uint8_t *put_dht(uint8_t *p, int tc_th, const uint8_t bits[16], const uint8_t *vals, int n)
{
        *p++ = 0xFF, *p++ = 0xC4;
        p = put16(p, 2 + 1 + 16 + n);
        *p++ = (uint8_t) tc_th;
        memcpy(p, bits, 16), p += 16;
        memcpy(p, vals, n), p += n;
        return p;
}

/* everything up to (not including) the first SOS; shared with the product's header writer in spirit, restated here */
static uint8_t *write_headers(uint8_t *p, int w, int h, int fmt, const uint8_t ql[64], const uint8_t qc[64], int ri)
{
        *p++ = 0xFF, *p++ = 0xD8; /* SOI */
        if (fmt == FMT_RGB_444) { /* Adobe APP14, transform 0 = components are RGB (jpeg_reader.c understands it) */
                static const uint8_t adobe[] = { 0xFF, 0xEE, 0, 14, 'A', 'd', 'o', 'b', 'e', 0, 100, 0, 0, 0, 0, 0 };
                memcpy(p, adobe, sizeof adobe), p += sizeof adobe;
        } else { /* JFIF APP0 as jpeg_writer.c:232-247 */
                static const uint8_t jfif[] = { 0xFF, 0xE0, 0, 16, 'J', 'F', 'I', 'F', 0, 1, 1, 0, 0, 1, 0, 1, 0, 0 };
                memcpy(p, jfif, sizeof jfif), p += sizeof jfif;
        }
        for (int t = 0; t < 2; ++t) { /* DQT, zig-zag order */
                *p++ = 0xFF, *p++ = 0xDB;
                p = put16(p, 67);
                *p++ = (uint8_t) t;
                for (int k = 0; k < 64; ++k) {
                        *p++ = (t ? qc : ql)[ugb_jpeg_zigzag[k]];
                }
        }
        *p++ = 0xFF, *p++ = 0xC0; /* SOF0 */
        p = put16(p, 17);
        *p++ = 8;
        p = put16(p, h), p = put16(p, w);
        *p++ = 3;
        for (int c = 0; c < 3; ++c) {
                *p++ = (uint8_t) (c + 1);
                *p++ = (fmt == FMT_UYVY_422 && c == 0) ? 0x21 : 0x11;
                *p++ = c == 0 ? 0 : 1; /* component 0 uses table 0 (K.1), the others table 1 (K.2) */
        }
        p = put_dht(p, 0x00, ugb_jpeg_dc_luma_bits, ugb_jpeg_dc_vals, 12);
        p = put_dht(p, 0x10, ugb_jpeg_ac_luma_bits, ugb_jpeg_ac_luma_vals, 162);
        p = put_dht(p, 0x01, ugb_jpeg_dc_chroma_bits, ugb_jpeg_dc_vals, 12);
        p = put_dht(p, 0x11, ugb_jpeg_ac_chroma_bits, ugb_jpeg_ac_chroma_vals, 162);
        if (ri > 0) {
                *p++ = 0xFF, *p++ = 0xDD;
                p = put16(p, 4), p = put16(p, ri);
        }
        return p;
}
static uint8_t *write_sos(uint8_t *p, int first, int ncomp)
{
        *p++ = 0xFF, *p++ = 0xDA;
        p = put16(p, 6 + 2 * ncomp);
        *p++ = (uint8_t) ncomp;
        for (int c = first; c < first + ncomp; ++c) {
                *p++ = (uint8_t) (c + 1);
                *p++ = c == 0 ? 0x00 : 0x11;
        }
        *p++ = 0, *p++ = 63, *p++ = 0;
        return p;
}

static int clampi(int v, int hi) { return v < 0 ? 0 : v > hi ? hi : v; }

/* gather one 8x8 block of component `comp` at block coordinates (bx, by) of that component's grid; edges replicate */
static void gather(const uint8_t *src, long pitch, int w, int h, int fmt, int comp, int bx, int by, uint8_t px[64])
{
        for (int y = 0; y < 8; ++y) {
                const uint8_t *row = src + (long) clampi(by * 8 + y, h - 1) * pitch;
                for (int x = 0; x < 8; ++x) {
                        if (fmt == FMT_RGB_444) {
                                px[8 * y + x] = row[3 * clampi(bx * 8 + x, w - 1) + comp];
                        } else if (comp == 0) { /* UYVY: U Y0 V Y1 */
                                px[8 * y + x] = row[2 * clampi(bx * 8 + x, w - 1) + 1];
                        } else { /* chroma sample cx covers pixels 2cx, 2cx+1 */
                                const int cx = clampi(bx * 8 + x, (w + 1) / 2 - 1);
                                px[8 * y + x] = row[4 * cx + (comp == 1 ? 0 : 2)];
                        }
                }
        }
}

API int orc_jpeg_default_restart_interval(int fmt) { return fmt == FMT_RGB_444 ? 8 : 4; } /* gpujpeg.cpp:351 */

API size_t orc_jpeg_encode_ex(const uint8_t *src, long pitch, int w, int h, int fmt, int quality, int ri, int interleaved, uint8_t *out, size_t cap);
/* @returns number of bytes written (0 on error) */
API size_t orc_jpeg_encode(const uint8_t *src, long pitch, int w, int h, int fmt, int quality, int ri, uint8_t *out, size_t cap)
{
        return orc_jpeg_encode_ex(src, pitch, w, h, fmt, quality, ri, 0, out, cap);
}

/* interleaved (RGB input only): one scan whose MCU is the R, G and B block of an 8x8 area - what the reference module asks of GPUJPEG with its
 * `interleaved` option (gpujpeg.cpp:303,397-398); the default for RGB is one scan per component */
API size_t orc_jpeg_encode_ex(const uint8_t *src, long pitch, int w, int h, int fmt, int quality, int ri, int interleaved, uint8_t *out, size_t cap)
{
        /* worst case: 1658 bits per block, every byte stuffed (416 B), + RSTn per segment + headers */
        const size_t nblk = fmt == FMT_UYVY_422 ? (size_t) ((w + 15) / 16) * ((h + 7) / 8) * 4 : (size_t) ((w + 7) / 8) * ((h + 7) / 8) * 3;
        if (w <= 0 || h <= 0 || cap < 2048 + nblk * 418) {
                return 0;
        }
        if (ri <= 0) {
                ri = orc_jpeg_default_restart_interval(fmt);
        }
        uint8_t ql[64], qc[64];
        float ml[64], mc[64];
        ugb_jpeg_scaled_qtable(ugb_jpeg_q_luma, quality, ql);
        ugb_jpeg_scaled_qtable(ugb_jpeg_q_chroma, quality, qc);
        ugb_jpeg_quant_multipliers(ql, ml);
        ugb_jpeg_quant_multipliers(qc, mc);
        struct huff dcl, acl, dcc, acc;
        ugb_jpeg_build_codes(ugb_jpeg_dc_luma_bits, ugb_jpeg_dc_vals, 12, dcl.code, dcl.len);
        ugb_jpeg_build_codes(ugb_jpeg_ac_luma_bits, ugb_jpeg_ac_luma_vals, 162, acl.code, acl.len);
        ugb_jpeg_build_codes(ugb_jpeg_dc_chroma_bits, ugb_jpeg_dc_vals, 12, dcc.code, dcc.len);
        ugb_jpeg_build_codes(ugb_jpeg_ac_chroma_bits, ugb_jpeg_ac_chroma_vals, 162, acc.code, acc.len);

        uint8_t *p = write_headers(out, w, h, fmt, ql, qc, ri);
        uint8_t px[64];
        int16_t zz[64];
        if (fmt == FMT_UYVY_422) { /* one interleaved scan; MCU = 16x8 px = Y0 Y1 Cb Cr */
                p = write_sos(p, 0, 3);
                const int mw = (w + 15) / 16, mh = (h + 7) / 8, nm = mw * mh;
                struct bitw bw = { p, 0, 0 };
                int pred[3] = { 0, 0, 0 };
                for (int m = 0; m < nm; ++m) {
                        if (m > 0 && m % ri == 0) {
                                flush_bits(&bw);
                                *bw.p++ = 0xFF, *bw.p++ = (uint8_t) (0xD0 + ((m / ri - 1) & 7)); /* RSTn */
                                pred[0] = pred[1] = pred[2] = 0;
                        }
                        const int mx = m % mw, my = m / mw;
                        for (int k = 0; k < 4; ++k) {
                                const int comp = k < 2 ? 0 : k - 1;
                                gather(src, pitch, w, h, fmt, comp, comp == 0 ? mx * 2 + k : mx, my, px);
                                block_to_coeffs(px, comp == 0 ? ml : mc, zz);
                                encode_block(&bw, zz, &pred[comp], comp == 0 ? &dcl : &dcc, comp == 0 ? &acl : &acc);
                        }
                }
                flush_bits(&bw);
                p = bw.p;
        } else if (interleaved) { /* one scan; MCU = R, G, B block of the same 8x8 area */
                p = write_sos(p, 0, 3);
                const int bwid = (w + 7) / 8, bh = (h + 7) / 8, nm = bwid * bh;
                struct bitw bw = { p, 0, 0 };
                int pred[3] = { 0, 0, 0 };
                for (int m = 0; m < nm; ++m) {
                        if (m > 0 && m % ri == 0) {
                                flush_bits(&bw);
                                *bw.p++ = 0xFF, *bw.p++ = (uint8_t) (0xD0 + ((m / ri - 1) & 7));
                                pred[0] = pred[1] = pred[2] = 0;
                        }
                        for (int comp = 0; comp < 3; ++comp) {
                                gather(src, pitch, w, h, fmt, comp, m % bwid, m / bwid, px);
                                block_to_coeffs(px, comp == 0 ? ml : mc, zz);
                                encode_block(&bw, zz, &pred[comp], comp == 0 ? &dcl : &dcc, comp == 0 ? &acl : &acc);
                        }
                }
                flush_bits(&bw);
                p = bw.p;
        } else { /* three scans, one component each; MCU = one 8x8 block */
                const int bwid = (w + 7) / 8, bh = (h + 7) / 8, nb = bwid * bh;
                for (int comp = 0; comp < 3; ++comp) {
                        p = write_sos(p, comp, 1);
                        struct bitw bw = { p, 0, 0 };
                        int pred = 0;
                        for (int b = 0; b < nb; ++b) {
                                if (b > 0 && b % ri == 0) {
                                        flush_bits(&bw);
                                        *bw.p++ = 0xFF, *bw.p++ = (uint8_t) (0xD0 + ((b / ri - 1) & 7));
                                        pred = 0;
                                }
                                gather(src, pitch, w, h, fmt, comp, b % bwid, b / bwid, px);
                                block_to_coeffs(px, comp == 0 ? ml : mc, zz);
                                encode_block(&bw, zz, &pred, comp == 0 ? &dcl : &dcc, comp == 0 ? &acl : &acc);
                        }
                        flush_bits(&bw);
                        p = bw.p;
                }
        }
        *p++ = 0xFF, *p++ = 0xD9; /* EOI */
        return (size_t) (p - out);
}

/* The same stream produced by all host threads (bench.py's cpu_baseline for the JPEG workloads): restart segments are independent
 * (DC prediction restarts, the bit stream is byte aligned), so contiguous chunks of segments are coded by different threads into
 * private buffers and concatenated.  Byte-identical to orc_jpeg_encode (tests/test_jpeg.py). */
static size_t encode_segments(const uint8_t *src, long pitch, int w, int h, int fmt, int comp_of_scan, int ri, int seg0, int seg1, int nseg, int nmcu,
                              const float *ml, const float *mc, const struct huff *dcl, const struct huff *acl, const struct huff *dcc,
                              const struct huff *acc, uint8_t *dst)
{
        struct bitw bw = { dst, 0, 0 };
        uint8_t px[64];
        int16_t zz[64];
        const int mw = fmt == FMT_UYVY_422 ? (w + 15) / 16 : (w + 7) / 8;
        for (int s = seg0; s < seg1; ++s) {
                int pred[3] = { 0, 0, 0 };
                const int m1 = (s + 1) * ri < nmcu ? (s + 1) * ri : nmcu;
                for (int m = s * ri; m < m1; ++m) {
                        if (fmt == FMT_UYVY_422) {
                                for (int k = 0; k < 4; ++k) {
                                        const int comp = k < 2 ? 0 : k - 1;
                                        gather(src, pitch, w, h, fmt, comp, comp == 0 ? (m % mw) * 2 + k : m % mw, m / mw, px);
                                        block_to_coeffs(px, comp == 0 ? ml : mc, zz);
                                        encode_block(&bw, zz, &pred[comp], comp == 0 ? dcl : dcc, comp == 0 ? acl : acc);
                                }
                        } else {
                                gather(src, pitch, w, h, fmt, comp_of_scan, m % mw, m / mw, px);
                                block_to_coeffs(px, comp_of_scan == 0 ? ml : mc, zz);
                                encode_block(&bw, zz, &pred[0], comp_of_scan == 0 ? dcl : dcc, comp_of_scan == 0 ? acl : acc);
                        }
                }
                flush_bits(&bw);
                if (s != nseg - 1) {
                        *bw.p++ = 0xFF, *bw.p++ = (uint8_t) (0xD0 + (s & 7));
                }
        }
        return (size_t) (bw.p - dst);
}

API size_t orc_jpeg_encode_parallel(const uint8_t *src, long pitch, int w, int h, int fmt, int quality, int ri, uint8_t *out, size_t cap)
{
        const size_t nblk = fmt == FMT_UYVY_422 ? (size_t) ((w + 15) / 16) * ((h + 7) / 8) * 4 : (size_t) ((w + 7) / 8) * ((h + 7) / 8) * 3;
        if (w <= 0 || h <= 0 || cap < 2048 + nblk * 418) {
                return 0;
        }
        if (ri <= 0) {
                ri = orc_jpeg_default_restart_interval(fmt);
        }
        uint8_t ql[64], qc[64];
        float ml[64], mc[64];
        ugb_jpeg_scaled_qtable(ugb_jpeg_q_luma, quality, ql);
        ugb_jpeg_scaled_qtable(ugb_jpeg_q_chroma, quality, qc);
        ugb_jpeg_quant_multipliers(ql, ml);
        ugb_jpeg_quant_multipliers(qc, mc);
        struct huff dcl, acl, dcc, acc;
        ugb_jpeg_build_codes(ugb_jpeg_dc_luma_bits, ugb_jpeg_dc_vals, 12, dcl.code, dcl.len);
        ugb_jpeg_build_codes(ugb_jpeg_ac_luma_bits, ugb_jpeg_ac_luma_vals, 162, acl.code, acl.len);
        ugb_jpeg_build_codes(ugb_jpeg_dc_chroma_bits, ugb_jpeg_dc_vals, 12, dcc.code, dcc.len);
        ugb_jpeg_build_codes(ugb_jpeg_ac_chroma_bits, ugb_jpeg_ac_chroma_vals, 162, acc.code, acc.len);
        const int nscan = fmt == FMT_UYVY_422 ? 1 : 3, bpm = fmt == FMT_UYVY_422 ? 4 : 1;
        const int nmcu = fmt == FMT_UYVY_422 ? ((w + 15) / 16) * ((h + 7) / 8) : ((w + 7) / 8) * ((h + 7) / 8);
        const int nseg = (nmcu + ri - 1) / ri;
        int nchunk = omp_get_max_threads() * 4;
        nchunk = nchunk > nseg ? nseg : nchunk;
        const int per = (nseg + nchunk - 1) / nchunk;
        nchunk = (nseg + per - 1) / per;
        const size_t chunk_cap = (size_t) per * ((size_t) ri * bpm * 418 + 8);
        uint8_t *tmp = (uint8_t *) malloc(chunk_cap * nchunk);
        size_t *len = (size_t *) malloc(sizeof(size_t) * nchunk);
        if (!tmp || !len) {
                free(tmp), free(len);
                return 0;
        }
        uint8_t *p = write_headers(out, w, h, fmt, ql, qc, ri);
        for (int scan = 0; scan < nscan; ++scan) {
                p = write_sos(p, scan, fmt == FMT_UYVY_422 ? 3 : 1);
#pragma omp parallel for schedule(dynamic, 1)
                for (int c = 0; c < nchunk; ++c) {
                        const int s0 = c * per, s1 = s0 + per < nseg ? s0 + per : nseg;
                        len[c] = encode_segments(src, pitch, w, h, fmt, scan, ri, s0, s1, nseg, nmcu, ml, mc, &dcl, &acl, &dcc, &acc, tmp + chunk_cap * c);
                }
                size_t off = 0;
                for (int c = 0; c < nchunk; ++c) {
                        off += len[c];
                }
                size_t *start = len;  /* exclusive prefix in place, then a parallel copy */
                size_t run = 0;
                for (int c = 0; c < nchunk; ++c) {
                        const size_t l = len[c];
                        start[c] = run;
                        run += l;
                }
#pragma omp parallel for schedule(static)
                for (int c = 0; c < nchunk; ++c) {
                        const size_t l = (c + 1 < nchunk ? start[c + 1] : off) - start[c];
                        memcpy(p + start[c], tmp + chunk_cap * c, l);
                }
                p += off;
        }
        free(tmp), free(len);
        *p++ = 0xFF, *p++ = 0xD9;
        return (size_t) (p - out);
}

/* quantised zig-zag coefficients of every block in scan order (for stage-by-stage GPU parity) */
API void orc_jpeg_coefficients(const uint8_t *src, long pitch, int w, int h, int fmt, int quality, int16_t *out)
{
        uint8_t ql[64], qc[64], px[64];
        float ml[64], mc[64];
        ugb_jpeg_scaled_qtable(ugb_jpeg_q_luma, quality, ql);
        ugb_jpeg_scaled_qtable(ugb_jpeg_q_chroma, quality, qc);
        ugb_jpeg_quant_multipliers(ql, ml);
        ugb_jpeg_quant_multipliers(qc, mc);
        if (fmt == FMT_UYVY_422) {
                const int mw = (w + 15) / 16, mh = (h + 7) / 8;
                for (int m = 0; m < mw * mh; ++m) {
                        for (int k = 0; k < 4; ++k) {
                                const int comp = k < 2 ? 0 : k - 1;
                                gather(src, pitch, w, h, fmt, comp, comp == 0 ? (m % mw) * 2 + k : m % mw, m / mw, px);
                                block_to_coeffs(px, comp == 0 ? ml : mc, out + ((size_t) m * 4 + k) * 64);
                        }
                }
        } else {
                const int bw = (w + 7) / 8, bh = (h + 7) / 8;
                for (int comp = 0; comp < 3; ++comp) {
                        for (int b = 0; b < bw * bh; ++b) {
                                gather(src, pitch, w, h, fmt, comp, b % bw, b / bw, px);
                                block_to_coeffs(px, comp == 0 ? ml : mc, out + ((size_t) comp * bw * bh + b) * 64);
                        }
                }
        }
}

/* the quality-scaled Annex K table (natural order): which = 0 luminance (K.1), 1 chrominance (K.2) */
API void orc_jpeg_scaled_qtable(int which, int quality, uint8_t out[64]) { ugb_jpeg_scaled_qtable(which ? ugb_jpeg_q_chroma : ugb_jpeg_q_luma, quality, out); }
