/* The reference's small device-to-device pixel conversions (src/utils/cuda_pix_conv.h:7-37; used by video_capture/gpustitch.cpp:192-195,
 * 386-389), same C++ names and signatures so that a translation unit including this header links against libugb200 instead of
 * src/utils/cuda_pix_conv.cu.  All pointers are DEVICE pointers, pitches in bytes, the call is asynchronous on `stream`.
 * Results are bit-identical to the reference kernels as built by nvcc for sm_100a (the float matrix of UYVY->RGBA follows the
 * FMA contraction of that build: y = 1.164f * (Y - 16); R = fma(v, 1.793f, y); G = fma(v, -0.534f, y) - 0.213f * u; B = fma(u, 2.115f, y);
 * x > 0 ? (x < 255 ? trunc(x) : 255) : 0; alpha byte 0). */
#ifndef UGB200_CUDA_PIX_CONV_H
#define UGB200_CUDA_PIX_CONV_H
#ifdef __cplusplus
#include <stddef.h>

#include "cuda_wrapper.h"
struct CUstream_st;

UGB_API void cuda_RGB_to_RGBA(unsigned char *dst, size_t dstPitch, unsigned char *src, size_t srcPitch, size_t width, size_t height, struct CUstream_st *stream);
UGB_API void cuda_RGBA_to_RGB(unsigned char *dst, size_t dstPitch, unsigned char *src, size_t srcPitch, size_t width, size_t height, struct CUstream_st *stream);
UGB_API void cuda_RGBA_to_UYVY(unsigned char *dst, size_t dstPitch, unsigned char *src, size_t srcPitch, size_t width, size_t height, struct CUstream_st *stream);
UGB_API void cuda_UYVY_to_RGBA(unsigned char *dst, size_t dstPitch, unsigned char *src, size_t srcPitch, size_t width, size_t height, struct CUstream_st *stream);
#endif
#endif
