/* Plain-C driver of the C++ video_compress module layer (ultragrid_b200/csrc/host/video_compress.h), for tests,
 * bench.py and non-C++ hosts.  It stands where UltraGrid's sender calls compress_init / compress_frame / compress_pop
 * (src/video_compress.h:95-107, call sites src/rxtx.cpp:183-194,260-289).
 */
#ifndef UGB200_VCOMPRESS_H
#define UGB200_VCOMPRESS_H

#include <stddef.h>

#include "cuda_wrapper.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ugb200_compress ugb200_compress;

/* -D/--cuda-device a,b,... (src/main.cpp:402-431 -> cuda_devices[], src/host.cpp:177-179); up to 8 devices */
UGB_API int ugb200_set_cuda_devices(const int *devices, int count);

/* compress_init(parent, config): "cuda_dxt[:DXT1|:DXT5]" (src/video_compress/cuda_dxt.cpp:108-119; asynchronous, 3 frames in
 * flight; "cuda_dxt_sync" is the same module behind the reference's synchronous tile API) or
 * "GPUJPEG[:q=<1-100>][:restart=<n>][:lanes=<1-8>]" (src/video_compress/gpujpeg.cpp:371-424; lanes = frames in flight per CUDA device,
 * default 3, B200 addition; 1 on a single device = the reference's synchronous push).  NULL on error. */
UGB_API ugb200_compress *ugb200_compress_init(const char *config);

/* compress_frame(): hand one frame to the compressor.  data is a host pointer (mem_location 0 = CPU_MEM) or a device
 * pointer (1 = CUDA_MEM, src/types.h:295-298); it must stay valid until the frame has been popped (the reference holds the input shared_ptr the same way).  data == NULL passes the
 * poison pill that ends the stream (src/video_compress.h:143-147).  0 ok, <0 error. */
UGB_API int ugb200_compress_push(ugb200_compress *s, const void *data, int mem_location, int width, int height, int codec,
                                 double fps);

/* compress_pop(): blocks for the next compressed frame, in submission order.  Copies it to out (capacity cap).
 * 0 ok; 1 end of stream (poison pill came through); -1 error / buffer too small. */
UGB_API int ugb200_compress_pop(ugb200_compress *s, void *out, size_t cap, size_t *out_len, int *out_codec, unsigned *seq);

/* Same as ugb200_compress_pop without the copy: *data points into the pooled (pinned) output frame and stays valid until the
 * next pop / done on this handle — what UltraGrid's sender gets as shared_ptr<video_frame>. */
UGB_API int ugb200_compress_pop_ref(ugb200_compress *s, const void **data, size_t *len, int *out_codec, unsigned *seq);

UGB_API void ugb200_compress_done(ugb200_compress *s);

/* ---- decompress side (src/video_decompress.h:173-213; modules gpujpeg, gpujpeg_to_dxt, dxt_cuda) --------------------------------------
 * decompress_init_multi(compression, internal, to, &state, 1): the registered module with the best priority for compression
 * (UGB_JPEG, UGB_DXT1, UGB_DXT5) -> out_codec (UGB_RGB / RGBA / UYVY / DXT1 / DXT5, or UGB_VIDEO_CODEC_NONE to probe).  NULL if none. */
typedef struct ugb200_decompress ugb200_decompress;
UGB_API ugb200_decompress *ugb200_decompress_init(int compression, int out_codec);
UGB_API const char *ugb200_decompress_module(ugb200_decompress *d);  /* name of the module that was selected */
/* decompress_reconfigure(state, desc{width, height, compression}, rshift, gshift, bshift, pitch, out_codec): non-zero = ok */
UGB_API int ugb200_decompress_reconfigure(ugb200_decompress *d, int width, int height, int compression, int rshift, int gshift, int bshift, int pitch,
                                          int out_codec);
/* decompress_frame(): src and dst are HOST buffers.  Returns decompress_status: 0 no frame, 1 got frame, 2 got codec (probe: internal_props[3] =
 * depth, subsampling (4440/4220/4200), rgb), 3 unsupported pixel format. */
UGB_API int ugb200_decompress_frame(ugb200_decompress *d, void *dst, const void *src, unsigned src_len, int frame_seq, int *internal_props);
UGB_API void ugb200_decompress_done(ugb200_decompress *d);

/* get_best_decoder_from(in, candidates, &out) (src/pixfmt_conv.c:3148-3172): the codec_t a module converts `in` to when it
 * natively accepts `candidates`; 0 (VIDEO_CODEC_NONE) if no conversion exists. */
UGB_API int ugb200_get_best_decoder_from(int in_codec, const int *candidates, int count);

#ifdef __cplusplus
}
#endif
#endif
