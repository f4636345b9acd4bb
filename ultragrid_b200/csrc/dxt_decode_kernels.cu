// DXT1 and DXT5-YCoCg block DEcoders.  The reference decodes these textures only on the
// display side in OpenGL (src/video_decompress/dxt_glsl.c, dxt_compress/display_dxt5ycocg_fp.glsl); its one CPU decoder is the
// tool cuda_dxt/dxt62tga.c:24-108 (DXT5-YCoCg -> BGR, all arithmetic in double), and that is the contract here: the same
// operations in the same order with explicit-rounding FP64 intrinsics (no contraction), byte-identical output.
// DXT1 has no CPU decoder in the tree (parity unpinned): the S3TC rule is evaluated the same way dxt62tga.c does it for the
// colour block of DXT5 (endpoints / 31.0, / 63.0, thirds in double, (int)(255 * v + 0.5)), plus the 3-colour mode when
// color0 <= color1 (midpoint, index 3 = black).
//
// A thread decodes one 4x4 block: 8 or 16 bytes in, 4 rows x 12 bytes out.  1 B/px + 3 B/px, FP64 on the pixel path
// (DADD/DMUL run at half the FP32 rate on this part; the double -> int conversion is the slow instruction).
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>

#include "../../include/ugb200.h"

namespace ugb {

__device__ __forceinline__ uint32_t to_byte(double s)  // dxt62tga.c:14-21
{
        const int is = __double2int_rz(__dadd_rn(s, 0.5));
        return (uint32_t) min(max(is, 0), 255);
}
__device__ __forceinline__ double third(double a, double b)  // (2.0 * a + 1.0 * b) / 3.0, dxt62tga.c:76-81
{
        return __ddiv_rn(__dadd_rn(__dmul_rn(2.0, a), b), 3.0);
}

/// four pixels (24-bit colours p0..p3, R in the low byte) = 12 bytes of a row: three 32-bit stores when the row is 4-byte aligned
__device__ __forceinline__ void store_row(uint8_t *o, uint32_t p0, uint32_t p1, uint32_t p2, uint32_t p3, bool aligned)
{
        const uint32_t w0 = p0 | p1 << 24, w1 = (p1 >> 8) | p2 << 16, w2 = (p2 >> 16) | p3 << 8;
        if (aligned) {
                ((uint32_t *) o)[0] = w0, ((uint32_t *) o)[1] = w1, ((uint32_t *) o)[2] = w2;
        } else {
#pragma unroll
                for (int i = 0; i < 12; ++i) {
                        o[i] = (uint8_t) ((i < 4 ? w0 : i < 8 ? w1 : w2) >> (8 * (i & 3)));
                }
        }
}
template <int BGR>
__device__ __forceinline__ uint32_t pack_px(uint32_t r, uint32_t g, uint32_t b)
{
        return BGR ? b | g << 8 | r << 16 : r | g << 8 | b << 16;
}

// ---- DXT1 palette tables ---------------------------------------------------------------------------------------------------------------
// The four palette colours of a block are a function of its two 5:6:5 endpoints only, channel by channel.  Round 1 evaluated them per block in
// FP64 (24 software divisions + 12 double -> int conversions per block: 0.31 of the HBM roofline).  The same FP64 expressions are now evaluated
// ONCE per device for every endpoint pair by dxt1_tables_kernel - identical operations, so identical bytes - and a block reads the four bytes of
// a channel with one 32-bit load: [mode][q0][q1] -> entries 0..3, mode 0 = four-colour (c0 > c1), 1 = three-colour + black.
__device__ uint32_t g_dxt1_pal5[2][32][32], g_dxt1_pal6[2][64][64];
// DXT5-YCoCg: the FP64 quotients that depend on 5/6/8-bit codes only (dxt62tga.c:38-81), same expressions evaluated once per device:
//   g_q31[q] = q / 31.0, g_q63[q] = q / 63.0, g_q255[q] = q / 255.0;  g_t5[q0][q1] = (2 q0/31 + q1/31) / 3 (entry 2; entry 3 = [q1][q0]), g_t6 likewise;
//   g_s1[q] = 1 / (31.875 * q/31 + 1), g_s3[q0][q1] = 1 / (31.875 * g_t5[q0][q1] + 1)   (the scale of :24-27 per palette entry)
__device__ double g_q31[32], g_q63[64], g_q255[256], g_t5[32][32], g_t6[64][64], g_s1[32], g_s3[32][32];

__global__ void dxt1_tables_kernel()
{
        const int i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= 2 * 64 * 64) {
                return;
        }
        const int mode = i >> 12, q0 = (i >> 6) & 63, q1 = i & 63;
        if (mode == 0) {  // the double tables (one thread per pair)
                const double a6 = __ddiv_rn((double) q0, 63.0), b6 = __ddiv_rn((double) q1, 63.0);
                g_t6[q0][q1] = third(a6, b6);
                if (q1 == 0) {
                        g_q63[q0] = a6;
                }
                if (q0 < 32 && q1 < 32) {
                        const double a5 = __ddiv_rn((double) q0, 31.0), b5 = __ddiv_rn((double) q1, 31.0), t = third(a5, b5);
                        g_t5[q0][q1] = t;
                        g_s3[q0][q1] = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(31.875, t), 1.0));
                        if (q1 == 0) {
                                g_q31[q0] = a5;
                                g_s1[q0] = __ddiv_rn(1.0, __dadd_rn(__dmul_rn(31.875, a5), 1.0));
                        }
                }
                if (q0 < 4) {
                        const int c = q0 * 64 + q1;
                        g_q255[c] = __ddiv_rn((double) c, 255.0);
                }
        }
#pragma unroll
        for (int six = 0; six < 2; ++six) {
                if (!six && (q0 >= 32 || q1 >= 32)) {
                        continue;
                }
                const double den = six ? 63.0 : 31.0;
                double v[4];
                v[0] = __ddiv_rn((double) q0, den), v[1] = __ddiv_rn((double) q1, den);
                if (mode == 0) {
                        v[2] = third(v[0], v[1]);
                        v[3] = __ddiv_rn(__dadd_rn(v[0], __dmul_rn(2.0, v[1])), 3.0);
                } else {
                        v[2] = __dmul_rn(__dadd_rn(v[0], v[1]), 0.5);
                        v[3] = 0.0;
                }
                uint32_t w = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                        w |= to_byte(__dmul_rn(v[k], 255.0)) << (8 * k);
                }
                if (six) {
                        g_dxt1_pal6[mode][q0][q1] = w;
                } else {
                        g_dxt1_pal5[mode][q0][q1] = w;
                }
        }
}

template <int BGR>
__global__ void __launch_bounds__(128) dxt5ycocg_decode_kernel(const uint4 *__restrict__ in, uint8_t *__restrict__ out, int bw, int bh, long pitch, bool aligned)
{
        // per-thread palettes live in shared memory ([entry][thread]: conflict-free, dynamically indexable without local memory)
        __shared__ double s_ap[8][128], s_co[4][128], s_cg[4][128];
        const int tid = threadIdx.x, bx = blockIdx.x * blockDim.x + tid, by = blockIdx.y;
        if (bx >= bw) {
                return;
        }
        const uint4 blk = __ldg(in + (long) by * bw + bx);
        uint64_t alpha_code = (uint64_t) blk.x | (uint64_t) blk.y << 32, rgb_code = (uint64_t) blk.z | (uint64_t) blk.w << 32;
        // alpha palette (dxt62tga.c:38-62): 8 luma levels
        const double a0 = g_q255[alpha_code & 0xFF], a1 = g_q255[(alpha_code >> 8) & 0xFF];
        s_ap[0][tid] = a0, s_ap[1][tid] = a1;
        if (a0 > a1) {
#pragma unroll
                for (int k = 2; k < 8; ++k) {
                        s_ap[k][tid] = __ddiv_rn(__dadd_rn(__dmul_rn((double) (8 - k), a0), __dmul_rn((double) (k - 1), a1)), 7.0);
                }
        } else {
#pragma unroll
                for (int k = 2; k < 6; ++k) {
                        s_ap[k][tid] = __ddiv_rn(__dadd_rn(__dmul_rn((double) (6 - k), a0), __dmul_rn((double) (k - 1), a1)), 5.0);
                }
                s_ap[6][tid] = 0.0, s_ap[7][tid] = 1.0;
        }
        // colour palette (:68-81) and, per entry, the scaled Co / Cg (:24-27 depend on the entry only): every quotient is a function of the
        // 5 / 6-bit endpoint codes and comes from the per-device tables (the same FP64 expressions, evaluated by dxt1_tables_kernel)
        const int b0 = (int) (rgb_code & 0x1F), g0 = (int) ((rgb_code >> 5) & 0x3F), r0 = (int) ((rgb_code >> 11) & 0x1F);
        const int b1 = (int) ((rgb_code >> 16) & 0x1F), g1 = (int) ((rgb_code >> 21) & 0x3F), r1 = (int) ((rgb_code >> 27) & 0x1F);
        const double r[4] = { g_q31[r0], g_q31[r1], g_t5[r0][r1], g_t5[r1][r0] }, g[4] = { g_q63[g0], g_q63[g1], g_t6[g0][g1], g_t6[g1][g0] };
        const double scale[4] = { g_s1[b0], g_s1[b1], g_s3[b0][b1], g_s3[b1][b0] };
#pragma unroll
        for (int k = 0; k < 4; ++k) {
                s_co[k][tid] = __dmul_rn(__dadd_rn(r[k], -5.01960814E-01), scale[k]);
                s_cg[k][tid] = __dmul_rn(__dadd_rn(g[k], -5.01960814E-01), scale[k]);
        }
        alpha_code >>= 16, rgb_code >>= 32;
        uint8_t *o = out + (long) by * 4 * pitch + (long) bx * 12;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
                uint32_t px[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                        const double a = s_ap[alpha_code & 7][tid];
                        const int k = (int) (rgb_code & 3);
                        alpha_code >>= 3, rgb_code >>= 2;
                        const double co = s_co[k][tid], cg = s_cg[k][tid];
                        const double R = __dadd_rn(__dadd_rn(a, co), -cg), G = __dadd_rn(a, cg), B = __dadd_rn(__dadd_rn(a, -co), -cg);
                        px[x] = pack_px<BGR>(to_byte(__dmul_rn(R, 255.0)), to_byte(__dmul_rn(G, 255.0)), to_byte(__dmul_rn(B, 255.0)));
                }
                store_row(o + y * pitch, px[0], px[1], px[2], px[3], aligned);
        }
}

template <int BGR>
__global__ void __launch_bounds__(128) dxt1_decode_kernel(const uint2 *__restrict__ in, uint8_t *__restrict__ out, int bw, int bh, long pitch, bool aligned)
{
        const int bx = blockIdx.x * blockDim.x + threadIdx.x, by = blockIdx.y;
        if (bx >= bw) {
                return;
        }
        const uint2 blk = __ldg(in + (long) by * bw + bx);
        const uint32_t c0 = blk.x & 0xffff, c1 = blk.x >> 16;
        const int mode = c0 > c1 ? 0 : 1;
        const uint32_t pr = g_dxt1_pal5[mode][c0 >> 11][c1 >> 11], pg = g_dxt1_pal6[mode][(c0 >> 5) & 63][(c1 >> 5) & 63], pb = g_dxt1_pal5[mode][c0 & 31][c1 & 31];
        uint32_t pal[4];  // the four colours as 24-bit pixels, selected per pixel with two predicated moves
#pragma unroll
        for (int k = 0; k < 4; ++k) {
                pal[k] = pack_px<BGR>((pr >> (8 * k)) & 0xff, (pg >> (8 * k)) & 0xff, (pb >> (8 * k)) & 0xff);
        }
        uint32_t idx = blk.y;
        uint8_t *o = out + (long) by * 4 * pitch + (long) bx * 12;
#pragma unroll
        for (int y = 0; y < 4; ++y) {
                uint32_t px[4];
#pragma unroll
                for (int x = 0; x < 4; ++x) {
                        const uint32_t lo = (idx & 1) ? pal[1] : pal[0], hi = (idx & 1) ? pal[3] : pal[2];
                        px[x] = (idx & 2) ? hi : lo;
                        idx >>= 2;
                }
                store_row(o + y * pitch, px[0], px[1], px[2], px[3], aligned);
        }
}

}  // namespace ugb

using namespace ugb;

#define UGB_DECODE(name, KERNEL, T)                                                                                                        \
        extern "C" UGB_API int name(const void *src, void *out, int w, int h, long out_pitch, int bgr, cuda_wrapper_stream_t stream)       \
        {                                                                                                                                  \
                if (src == nullptr || out == nullptr || w <= 0 || h <= 0 || (w & 3) || (h & 3) || (sizeof(T) - 1 & (size_t) src)) {        \
                        return -1; /* the block codecs need multiples of 4 (cuda_dxt.cu:745) */                                           \
                }                                                                                                                          \
                if (out_pitch == 0) {                                                                                                      \
                        out_pitch = (long) w * 3;                                                                                          \
                }                                                                                                                          \
                UGB_DECODE_PRE_##KERNEL(stream)                                                                                              \
                dim3 grid((w / 4 + 127) / 128, h / 4);                                                                                     \
                const bool aligned = !(3 & (size_t) out) && !(out_pitch & 3);                                                              \
                if (bgr) {                                                                                                                 \
                        KERNEL<1><<<grid, 128, 0, (cudaStream_t) stream>>>((const T *) src, (uint8_t *) out, w / 4, h / 4, out_pitch, aligned);    \
                } else {                                                                                                                   \
                        KERNEL<0><<<grid, 128, 0, (cudaStream_t) stream>>>((const T *) src, (uint8_t *) out, w / 4, h / 4, out_pitch, aligned);    \
                }                                                                                                                          \
                return cudaGetLastError() == cudaSuccess ? 0 : -2;                                                                         \
        }
namespace ugb {
/// the DXT1 palette tables exist once per device; the first decode call on a device builds them (host-synchronised, a one-off of a few
/// microseconds - not inside a stream capture)
static int ensure_dxt1_tables(cudaStream_t stream)
{
        static std::mutex m;
        static bool ready[64] = {};
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
                return -2;
        }
        std::lock_guard<std::mutex> lk(m);
        if (!ready[dev]) {
                (void) stream;
                cudaStream_t own;  // a stream of its own: the tables are complete (host-synchronised) before any decode kernel is queued anywhere
                if (cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking) != cudaSuccess) {
                        return -2;
                }
                dxt1_tables_kernel<<<(2 * 64 * 64 + 255) / 256, 256, 0, own>>>();
                const bool ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(own) == cudaSuccess;
                cudaStreamDestroy(own);
                if (!ok) {
                        return -2;
                }
                ready[dev] = true;
        }
        return 0;
}
}  // namespace ugb
#define UGB_DECODE_PRE_dxt1_decode_kernel(stream) if (ensure_dxt1_tables((cudaStream_t) (stream)) != 0) return -2;
#define UGB_DECODE_PRE_dxt5ycocg_decode_kernel(stream) if (ensure_dxt1_tables((cudaStream_t) (stream)) != 0) return -2;
UGB_DECODE(ugb200_dxt1_to_rgb, dxt1_decode_kernel, uint2)
UGB_DECODE(ugb200_dxt5ycocg_to_rgb, dxt5ycocg_decode_kernel, uint4)
