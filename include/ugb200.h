/* B200-native additions to the UltraGrid hot-path C ABI: fused / asynchronous block-compression entry
 * points, whole-buffer pixel-format conversion (device form of decoder_t line functions) and
 * packed->planar conversion.  Plain C, device pointers + sizes only.
 *
 * All functions here are ASYNCHRONOUS on `stream` (no implicit synchronisation) and return
 *   0 ok, -1 bad arguments / alignment, -2 launch failure, -4 unsupported conversion.
 */
#ifndef UGB200_H
#define UGB200_H

#include "cuda_wrapper.h"

#ifdef __cplusplus
extern "C" {
#endif

/* codec_t values — numerically identical to UltraGrid's enum (src/types.h:62-112) */
enum ugb200_codec {
        UGB_VIDEO_CODEC_NONE = 0,
        UGB_RGBA, UGB_UYVY, UGB_YUYV, UGB_VUYA, UGB_R10k, UGB_R12L, UGB_v210, UGB_DVS10, UGB_DXT1, UGB_DXT1_YUV,
        UGB_DXT5, UGB_RGB, UGB_JPEG, UGB_JPEG_XS, UGB_RAW, UGB_H264, UGB_H265, UGB_VP8, UGB_VP9, UGB_BGR, UGB_J2K,
        UGB_J2KR, UGB_HW_VDPAU, UGB_HFYU, UGB_FFV1, UGB_CFHD, UGB_RG48, UGB_AV1, UGB_I420, UGB_Y216, UGB_Y416,
        UGB_PRORES, UGB_PRORES_4444, UGB_PRORES_4444_XQ, UGB_PRORES_422_HQ, UGB_PRORES_422, UGB_PRORES_422_PROXY,
        UGB_PRORES_422_LT, UGB_APV, UGB_PYROWAVE, UGB_DRM_PRIME,
        UGB_VIDEO_CODEC_COUNT
};

/* ---- block compression (see cuda_dxt.h for the synchronous reference-compatible entry points) ---- */

/* same as cuda_{rgb,yuv}_to_dxt{1,6} (cuda_dxt/cuda_dxt.h:30-88) minus the cudaStreamSynchronize */
UGB_API int ugb200_rgb_to_dxt1_async(const void *src, void *out, int size_x, int size_y, cuda_wrapper_stream_t stream);
UGB_API int ugb200_yuv_to_dxt1_async(const void *src, void *out, int size_x, int size_y, cuda_wrapper_stream_t stream);
UGB_API int ugb200_rgb_to_dxt6_async(const void *src, void *out, int size_x, int size_y, cuda_wrapper_stream_t stream);
UGB_API int ugb200_yuv_to_dxt6_async(const void *src, void *out, int size_x, int size_y, cuda_wrapper_stream_t stream);

/* Fused UYVY -> DXT: replaces the pair cuda_yuv422_to_yuv444 + cuda_yuv_to_dxt{1,6} that
 * src/video_compress/cuda_dxt.cpp:223-257 runs (results are bit-identical to that pair).
 * src: device UYVY, 8-byte aligned (16 for the fast path); src_pitch bytes per row (0 = size_x*2);
 * size_x, size_y multiples of 4; negative size_y mirrors vertically. */
UGB_API int ugb200_uyvy_to_dxt1_async(const void *src, void *out, int size_x, int size_y, long src_pitch,
                              cuda_wrapper_stream_t stream);
UGB_API int ugb200_uyvy_to_dxt6_async(const void *src, void *out, int size_x, int size_y, long src_pitch,
                              cuda_wrapper_stream_t stream);

/* ---- pixel-format line converters over a whole buffer ---------------------------------------------
 * Device form of   for (y < height) decoder(dst + y*dst_pitch, src + y*src_pitch, dst_len, rs, gs, bs);
 * with decoder = get_decoder_from_to(in, out)  (src/pixfmt_conv.h:62-63, src/pixfmt_conv.c:3110-3125,
 * row loop as tools/convert.cpp:148-152).  Per-row results are byte-identical to the reference decoder,
 * including how many bytes of dst_len each loop really writes.  src_size = readable bytes from src
 * (0 = src_pitch*height): some reference loops over-read a partial pixel group; reads past src_size give 0. */
UGB_API int ugb200_pixfmt_supported(int in_codec, int out_codec); /* get_decoder_from_to() != NULL */
UGB_API int ugb200_pixfmt_convert(int in_codec, int out_codec, void *dst, long dst_pitch, const void *src, long src_pitch,
                          int dst_len, int height, long src_size, int rshift, int gshift, int bshift,
                          cuda_wrapper_stream_t stream);

/* ---- packed -> planar (src/to_planar.h:53-59) ----------------------------------------------------- */
struct ugb200_to_planar_data { /* same fields as struct to_planar_data; pointers are DEVICE pointers */
        int            width;
        int            height;
        unsigned char *out_data[4];
        unsigned       out_linesize[4];
        const unsigned char *in_data;
};
/* v210_to_p010le (src/to_planar.c:64-155). in_linesize 0 = vc_get_linesize(width, v210). */
UGB_API int ugb200_v210_to_p010le(const struct ugb200_to_planar_data *d, long in_linesize, cuda_wrapper_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
