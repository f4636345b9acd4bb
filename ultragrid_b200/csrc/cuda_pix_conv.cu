// Device-to-device RGB / RGBA / UYVY conversions with the reference's C++ interface (include/cuda_pix_conv.h; src/utils/cuda_pix_conv.cu).
// The reference launches one thread per PIXEL in 32 x 32 CTAs (RGB: 3-byte loads, UYVY: each thread re-reads the shared chroma word).
// Here a thread owns 4 pixels (16-byte RGBA side, 12-byte RGB / 8-byte UYVY side), rows on grid.y; results are bit-identical.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/cuda_pix_conv.h"

namespace ugb {

__device__ __forceinline__ uint32_t ld32(const uint8_t *p, bool al)
{
        return al ? *(const uint32_t *) p : (uint32_t) p[0] | (uint32_t) p[1] << 8 | (uint32_t) p[2] << 16 | (uint32_t) p[3] << 24;
}
__device__ __forceinline__ void st32(uint8_t *p, uint32_t v, bool al)
{
        if (al) {
                *(uint32_t *) p = v;
        } else {
                p[0] = (uint8_t) v, p[1] = (uint8_t) (v >> 8), p[2] = (uint8_t) (v >> 16), p[3] = (uint8_t) (v >> 24);
        }
}

/// kern_RGBtoRGBA (cuda_pix_conv.cu:7-29): alpha byte 0
__global__ void __launch_bounds__(128) rgb_to_rgba_kernel(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch, int width, bool al)
{
        const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y;
        if (x0 >= width) {
                return;
        }
        const uint8_t *s = src + (size_t) y * spitch + 3 * (size_t) x0;
        uint8_t *d = dst + (size_t) y * dpitch + 4 * (size_t) x0;
        if (x0 + 4 <= width) {
                const uint32_t a = ld32(s, al), b = ld32(s + 4, al), c = ld32(s + 8, al);
                const uint4 o = make_uint4(a & 0x00ffffffu, __byte_perm(a, b, 0x4543) & 0x00ffffffu, __byte_perm(b, c, 0x4432) & 0x00ffffffu, c >> 8);
                if (al) {
                        *(uint4 *) d = o;
                } else {
                        st32(d, o.x, false), st32(d + 4, o.y, false), st32(d + 8, o.z, false), st32(d + 12, o.w, false);
                }
        } else {
                for (int k = 0; x0 + k < width; ++k) {
                        d[4 * k] = s[3 * k], d[4 * k + 1] = s[3 * k + 1], d[4 * k + 2] = s[3 * k + 2], d[4 * k + 3] = 0;
                }
        }
}

/// kern_RGBAtoRGB (:32-54)
__global__ void __launch_bounds__(128) rgba_to_rgb_kernel(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch, int width, bool al)
{
        const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y;
        if (x0 >= width) {
                return;
        }
        const uint8_t *s = src + (size_t) y * spitch + 4 * (size_t) x0;
        uint8_t *d = dst + (size_t) y * dpitch + 3 * (size_t) x0;
        if (x0 + 4 <= width) {
                uint32_t p[4];
                if (al) {
                        const uint4 v = *(const uint4 *) s;
                        p[0] = v.x, p[1] = v.y, p[2] = v.z, p[3] = v.w;
                } else {
                        p[0] = ld32(s, false), p[1] = ld32(s + 4, false), p[2] = ld32(s + 8, false), p[3] = ld32(s + 12, false);
                }
                st32(d, __byte_perm(p[0], p[1], 0x4210), al);      // r0 g0 b0 r1
                st32(d + 4, __byte_perm(p[1], p[2], 0x5421), al);  // g1 b1 r2 g2
                st32(d + 8, __byte_perm(p[2], p[3], 0x6542), al);  // b2 r3 g3 b3
        } else {
                for (int k = 0; x0 + k < width; ++k) {
                        d[3 * k] = s[4 * k], d[3 * k + 1] = s[4 * k + 1], d[3 * k + 2] = s[4 * k + 2];
                }
        }
}

/// kern_UYVYtoRGBA (:60-92), arithmetic as the reference build contracts it (see the header)
__device__ __forceinline__ uint32_t sat_trunc(float x) { return x > 0.0f ? (x < 255.0f ? (uint32_t) __float2int_rz(x) : 255u) : 0u; }
__global__ void __launch_bounds__(128) uyvy_to_rgba_kernel(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch, int width, bool al)
{
        const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y;
        if (x0 >= width) {
                return;
        }
        const uint8_t *s = src + (size_t) y * spitch + 2 * (size_t) x0;
        uint8_t *d = dst + (size_t) y * dpitch + 4 * (size_t) x0;
        uint32_t o[4];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
                if (x0 + 2 * h >= width) {
                        break;
                }
                const uint32_t w = ld32(s + 4 * h, al);  // the reference reads the whole 4-byte block for an odd last pixel too
                const float u = (float) ((int) (w & 0xff) - 128), v = (float) ((int) ((w >> 16) & 0xff) - 128);
                const float g2 = __fmul_rn(u, 0.213f);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                        const float yy = __fmul_rn((float) ((int) ((w >> (8 + 16 * k)) & 0xff) - 16), 1.164f);
                        const uint32_t r = sat_trunc(__fmaf_rn(v, 1.793f, yy)), g = sat_trunc(__fadd_rn(__fmaf_rn(v, -0.534f, yy), -g2)), b = sat_trunc(__fmaf_rn(u, 2.115f, yy));
                        o[2 * h + k] = r | g << 8 | b << 16;
                }
        }
        if (al && x0 + 4 <= width) {
                *(uint4 *) d = make_uint4(o[0], o[1], o[2], o[3]);
        } else {
                for (int k = 0; k < 4 && x0 + k < width; ++k) {
                        st32(d + 4 * k, o[k], al);
                }
        }
}

/// kern_RGBAtoUYVY (:95-133): integer BT.709 in 16.16, chroma = C division by 2 of the pair's sum; an odd last pixel is not converted
__global__ void __launch_bounds__(128) rgba_to_uyvy_kernel(uint8_t *dst, size_t dpitch, const uint8_t *src, size_t spitch, int width, bool al)
{
        const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x), y = blockIdx.y;
        if (x0 + 1 >= width) {
                return;
        }
        const uint8_t *s = src + (size_t) y * spitch + 4 * (size_t) x0;
        uint8_t *d = dst + (size_t) y * dpitch + 2 * (size_t) x0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
                if (x0 + 2 * h + 1 >= width) {
                        break;
                }
                const uint32_t p1 = ld32(s + 8 * h, al), p2 = ld32(s + 8 * h + 4, al);
                const int r1 = p1 & 0xff, g1 = (p1 >> 8) & 0xff, b1 = (p1 >> 16) & 0xff, r2 = p2 & 0xff, g2 = (p2 >> 8) & 0xff, b2 = (p2 >> 16) & 0xff;
                const int y1 = 11993 * r1 + 40239 * g1 + 4063 * b1 + (1 << 20), y2 = 11993 * r2 + 40239 * g2 + 4063 * b2 + (1 << 20);
                int u = (-6619 * r1 - 22151 * g1 + 28770 * b1) + (-6619 * r2 - 22151 * g2 + 28770 * b2);
                int v = (28770 * r1 - 26149 * g1 - 2621 * b1) + (28770 * r2 - 26149 * g2 - 2621 * b2);
                u = u / 2 + (1 << 23), v = v / 2 + (1 << 23);
                const int lim = (1 << 24) - 1;
                st32(d + 4 * h, (uint32_t) (min(max(u, 0), lim) >> 16) | (uint32_t) (min(max(y1, 0), lim) >> 16) << 8 | (uint32_t) (min(max(v, 0), lim) >> 16) << 16 |
                                    (uint32_t) (min(max(y2, 0), lim) >> 16) << 24,
                     al);
        }
}

template <class K>
static void launch(K kernel, unsigned char *dst, size_t dpitch, unsigned char *src, size_t spitch, size_t width, size_t height, cudaStream_t s)
{
        if (width == 0 || height == 0 || height > 65535) {
                return;
        }
        const bool al = !(15 & (size_t) dst) && !(15 & (size_t) src) && !(dpitch & 15) && !(spitch & 15);
        const dim3 grid((unsigned) (((width + 3) / 4 + 127) / 128), (unsigned) height);
        kernel<<<grid, 128, 0, s>>>(dst, dpitch, src, spitch, (int) width, al);
}

}  // namespace ugb

void cuda_RGB_to_RGBA(unsigned char *dst, size_t dstPitch, unsigned char *src, size_t srcPitch, size_t width, size_t height, struct CUstream_st *stream)
{
        ugb::launch(ugb::rgb_to_rgba_kernel, dst, dstPitch, src, srcPitch, width, height, (cudaStream_t) stream);
}
void cuda_RGBA_to_RGB(unsigned char *dst, size_t dstPitch, unsigned char *src, size_t srcPitch, size_t width, size_t height, struct CUstream_st *stream)
{
        ugb::launch(ugb::rgba_to_rgb_kernel, dst, dstPitch, src, srcPitch, width, height, (cudaStream_t) stream);
}
void cuda_RGBA_to_UYVY(unsigned char *dst, size_t dstPitch, unsigned char *src, size_t srcPitch, size_t width, size_t height, struct CUstream_st *stream)
{
        ugb::launch(ugb::rgba_to_uyvy_kernel, dst, dstPitch, src, srcPitch, width, height, (cudaStream_t) stream);
}
void cuda_UYVY_to_RGBA(unsigned char *dst, size_t dstPitch, unsigned char *src, size_t srcPitch, size_t width, size_t height, struct CUstream_st *stream)
{
        ugb::launch(ugb::uyvy_to_rgba_kernel, dst, dstPitch, src, srcPitch, width, height, (cudaStream_t) stream);
}
