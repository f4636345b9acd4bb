/* TEST INFRASTRUCTURE — thin C entry points over the UNMODIFIED reference objects (compiled from
 * /root/reference where they lie by oracle/Makefile into oracle/_ref/libugref.so).  Nothing in the
 * product links or loads this.  Used by tests/ and by bench.py's cpu_baseline / --impl reference arms.
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "color_space.h"
#include "pixfmt_conv.h"
#include "from_planar.h"
#include "to_planar.h"
#include "utils/parallel_conv.h"
#include "video_codec.h"

#define API __attribute__((visibility("default")))

/* row loop exactly as tools/convert.cpp:148-152 */
API int ref_convert(int in_codec, int out_codec, unsigned char *dst, long dst_pitch, const unsigned char *src,
                    long src_pitch, int dst_len, int height, int rshift, int gshift, int bshift)
{
        decoder_t dec = get_decoder_from_to((codec_t) in_codec, (codec_t) out_codec);
        if (dec == NULL) {
                return -4;
        }
        for (int y = 0; y < height; ++y) {
                dec(dst + y * dst_pitch, src + y * src_pitch, dst_len, rshift, gshift, bshift);
        }
        return 0;
}

/* all host cores through the reference's own parallel_pix_conv (src/utils/parallel_conv.c:64-85) */
API int ref_convert_parallel(int in_codec, int out_codec, unsigned char *dst, int dst_pitch, const unsigned char *src,
                             int src_pitch, int height, int threads)
{
        decoder_t dec = get_decoder_from_to((codec_t) in_codec, (codec_t) out_codec);
        if (dec == NULL) {
                return -4;
        }
        parallel_pix_conv(height, (char *) dst, dst_pitch, (const char *) src, src_pitch, dec, threads);
        return 0;
}

API int ref_has_decoder(int in_codec, int out_codec)
{
        return get_decoder_from_to((codec_t) in_codec, (codec_t) out_codec) != NULL;
}

API int ref_vc_get_linesize(unsigned width, int codec) { return vc_get_linesize(width, (codec_t) codec); }
API int ref_vc_get_size(unsigned width, int codec) { return vc_get_size(width, (codec_t) codec); }
API const char *ref_get_codec_name(int codec) { return get_codec_name((codec_t) codec); }

API void ref_get_color_coeffs(int cs, int depth, int out[14])
{
        const struct color_coeffs *c = get_color_coeffs((enum colorspace) cs, depth);
        const int v[14] = { c->y_r, c->y_g, c->y_b, c->cb_r, c->cb_g, c->cb_b, c->cr_r, c->cr_g, c->cr_b,
                            c->y_scale, c->r_cr, c->g_cb, c->g_cr, c->b_cb };
        memcpy(out, v, sizeof v);
}

API void ref_v210_to_p010le(int width, int height, unsigned char *out_y, unsigned ls_y, unsigned char *out_c, unsigned ls_c,
                            const unsigned char *in)
{
        struct to_planar_data d = { 0 };
        d.width = width, d.height = height;
        d.out_data[0] = out_y, d.out_data[1] = out_c;
        d.out_linesize[0] = ls_y, d.out_linesize[1] = ls_c;
        d.in_data = in;
        v210_to_p010le(d);
}

/* decode_to_planar_parallel (src/to_planar.c:496-522); threads 0 = all cores */
API void ref_v210_to_p010le_parallel(int width, int height, unsigned char *out_y, unsigned ls_y, unsigned char *out_c,
                                     unsigned ls_c, const unsigned char *in, int threads)
{
        struct to_planar_data d = { 0 };
        d.width = width, d.height = height;
        d.out_data[0] = out_y, d.out_data[1] = out_c;
        d.out_linesize[0] = ls_y, d.out_linesize[1] = ls_c;
        d.in_data = in;
        decode_to_planar_parallel(v210_to_p010le, d, vc_get_linesize(width, v210), threads);
}

/* yuv422p10le_to_v210 (src/from_planar.c:295-333): the inverse the v210 <-> planar identity tests of the lavc bridge run through */
API void ref_yuv422p10le_to_v210(int width, int height, unsigned char *out, unsigned out_pitch, const unsigned char *y, const unsigned char *cb,
                                 const unsigned char *cr, unsigned ls_y, unsigned ls_c)
{
        struct from_planar_data d = { 0 };
        d.width = width, d.height = height, d.out_data = out, d.out_pitch = out_pitch;
        d.in_data[0] = y, d.in_data[1] = cb, d.in_data[2] = cr;
        d.in_linesize[0] = ls_y, d.in_linesize[1] = ls_c, d.in_linesize[2] = ls_c;
        d.in_depth = 10;
        yuv422p10le_to_v210(d);
}

API int ref_get_best_decoder_from(int in_codec, const int *candidates, int count)
{
        codec_t cand[VIDEO_CODEC_END + 1];
        int n = 0;
        for (; n < count && n < VIDEO_CODEC_END; ++n) {
                cand[n] = (codec_t) candidates[n];
        }
        cand[n] = VIDEO_CODEC_NONE;
        codec_t out = VIDEO_CODEC_NONE;
        return get_best_decoder_from((codec_t) in_codec, cand, &out) != NULL ? (int) out : 0;
}

/* jpeg_read_info / jpeg_get_rtp_hdr_data (src/utils/jpeg_reader.c:860-1003, :1092-1160): what an unmodified UltraGrid receiver does with a
 * JPEG bitstream before RFC 2435 packetisation.  out[]: width, height, comp_count, color_spec, interleaved, restart_interval,
 * sampling h[0..2], v[0..2], quantisation-table map [0..2], offset of the entropy-coded data; qt = the two 64-byte tables it found */
#include "utils/jpeg_reader.h"
API int ref_jpeg_read_info(unsigned char *image, int len, int *out, unsigned char *qt, unsigned char *huff)
{
        static struct jpeg_info info;
        memset(&info, 0, sizeof info);
        const int rc = jpeg_read_info(image, len, &info);
        if (rc != 0) {
                return rc;
        }
        out[0] = info.width, out[1] = info.height, out[2] = info.comp_count, out[3] = info.color_spec, out[4] = info.interleaved, out[5] = info.restart_interval;
        for (int i = 0; i < 3; ++i) {
                out[6 + i] = info.sampling_factor_h[i], out[9 + i] = info.sampling_factor_v[i], out[12 + i] = info.comp_table_quantization_map[i];
        }
        out[15] = (int) (info.data - image);
        for (int t = 0; t < 2; ++t) {
                if (info.quantization_tables[t]) {
                        memcpy(qt + 64 * t, info.quantization_tables[t], 64);
                }
        }
        memcpy(huff, info.huff_lum_dc, 272), memcpy(huff + 272, info.huff_lum_ac, 272), memcpy(huff + 544, info.huff_chm_dc, 272), memcpy(huff + 816, info.huff_chm_ac, 272);
        return 0;
}
/* @returns 1 if the stream can be sent as RFC 2435 RTP/JPEG; out[]: width, height, type, q, restart_interval, data offset */
API int ref_jpeg_get_rtp_hdr_data(unsigned char *image, int len, int *out)
{
        struct jpeg_rtp_data d;
        memset(&d, 0, sizeof d);
        if (!jpeg_get_rtp_hdr_data(image, len, &d)) {
                return 0;
        }
        out[0] = d.width, out[1] = d.height, out[2] = d.type, out[3] = d.q, out[4] = d.restart_interval, out[5] = (int) (d.data - image);
        return 1;
}

/* the exported line converters that are not in decoders[] (pixfmt_conv.h:93-101), row loop as tools/convert.cpp:148-152 */
API int ref_copyline_named(int func, unsigned char *dst, long dst_pitch, const unsigned char *src, long src_pitch, int dst_len, int height, int rs, int gs, int bs)
{
        decoder_t f = func == 1 ? vc_copylineABGRtoRGB : func == 2 ? vc_copylineBGRAtoRGB : func == 3 ? vc_copylineToRGBA_inplace : func == 4 ? vc_copylineUYVYtoGrayscale : NULL;
        if (!f) {
                return -4;
        }
        for (int y = 0; y < height; ++y) {
                f(dst + (size_t) y * dst_pitch, src + (size_t) y * src_pitch, dst_len, rs, gs, bs);
        }
        return 0;
}

/* The RECEIVER side of RFC 2435 in UltraGrid: rtpdec_jpeg.c:150-200 (create_jpeg_frame) rebuilds a decodable JPEG around the scan data of the
 * RTP payload with the reference's own jpeg_writer.c (:215-382): headers from (type, width, height, restart interval), the two quantisation tables
 * of the payload, the scan data, EOI.  @returns the length of the rebuilt stream */
#include "utils/jpeg_writer.h"
API long ref_jpeg_writer_rebuild(int type, int width, int height, int restart_interval, unsigned char qt[2][64], const unsigned char *scan, long scan_len,
                                 unsigned char *out)
{
        struct jpeg_writer_data info = { 0 };
        info.subsampling = (enum gpujpeg_writer_subsasmpling) (type & 1);
        info.width = (unsigned) width, info.height = (unsigned) height, info.restart_interval = (unsigned) restart_interval;
        char *p = (char *) out;
        jpeg_writer_write_headers(&p, &info);
        jpeg_writer_fill_dqt(info.dqt_marker_start, qt);
        memcpy(p, scan, (size_t) scan_len);
        p += scan_len;
        jpeg_writer_write_eoi(&p);
        return (long) (p - (char *) out);
}
