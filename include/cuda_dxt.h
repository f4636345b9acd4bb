/* DXT block-compression C ABI — drop-in for UltraGrid's cuda_dxt/cuda_dxt.h:30-89.
 *
 * Contract kept from the reference (cuda_dxt/cuda_dxt.cu:735-824):
 *   - src and out are DEVICE pointers; src 16-byte aligned, out 8-byte aligned (16 for DXT5-YCoCg)
 *   - src is packed 3 bytes/pixel (R,G,B or Y,U,V), no row padding; size_x, size_y multiples of 4
 *   - negative size_y = read the image bottom-up (vertical mirror)
 *   - the call is SYNCHRONOUS: it synchronises `stream` before returning
 *   - returns 0 ok, -1 bad size/alignment, -3 synchronisation failed (-2: launch failed, new)
 *   - output: DXT1  -> one 8-byte  block {u32 palette, u32 indices} per 4x4, raster block order
 *             DXT5  -> one 16-byte block {x,y: Y "alpha" block, z: CoCg 565 endpoints, w: indices}
 *   - results are bit-identical to the reference kernels built with nvcc 12.9 for sm_100a
 */
#ifndef UGB200_CUDA_DXT_H
#define UGB200_CUDA_DXT_H

#ifndef UGB_API
#define UGB_API __attribute__((visibility("default")))
#endif

#include "cuda_wrapper.h"

#ifdef __cplusplus
extern "C" {
#endif

/* replaces cuda_dxt.h:30-37 */
UGB_API int cuda_rgb_to_dxt1(const void *src, void *out, int size_x, int size_y, cuda_wrapper_stream_t stream);
/* replaces cuda_dxt.h:53-60 — input samples are Y,U,V (BT.709 limited) and are converted to RGB first */
UGB_API int cuda_yuv_to_dxt1(const void *src, void *out, int size_x, int size_y, cuda_wrapper_stream_t stream);
/* replaces cuda_dxt.h:76-83 — "DXT6" = DXT5-YCoCg */
UGB_API int cuda_rgb_to_dxt6(const void *src, void *out, int size_x, int size_y, cuda_wrapper_stream_t stream);
/* replaces cuda_dxt.h:85-86 */
UGB_API int cuda_yuv_to_dxt6(const void *src, void *out, int size_x, int size_y, cuda_wrapper_stream_t stream);
/* replaces cuda_dxt.h:87-88 — UYVY -> packed Y,U,V 4:4:4, chroma replicated; pix_count % 4 == 0 */
UGB_API int cuda_yuv422_to_yuv444(const void *src, void *out, int pix_count, cuda_wrapper_stream_t str);

#ifdef __cplusplus
}
#endif
#endif
