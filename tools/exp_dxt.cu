// Experiment harness (not product code): launch-shape / occupancy / control-flow variants of the fused UYVY -> DXT kernels, timed with
// CUDA events on four rotating 8K frames (264 MB > L2) and compared byte for byte with the shipped kernel's output.
// Build: nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -o tools/exp_dxt tools/exp_dxt.cu
// Run:   tools/exp_dxt            all variants, one line each
//        tools/exp_dxt one NAME   that variant only, 3 launches (for ncu)
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../ultragrid_b200/csrc/dxt_kernels.cu"

namespace ugb {

template <int DXT_TYPE, int BPT, int TPB, int MINB, bool BRANCH>
__global__ void __launch_bounds__(TPB, MINB) exp_kernel(const uint8_t *__restrict__ src, void *__restrict__ out, int wb, int h, long pitch)
{
        typedef typename block_out<DXT_TYPE>::type out_t;
        const int gx = blockIdx.x * blockDim.x + threadIdx.x;
        const int by = blockIdx.y;
        if (gx >= wb / BPT) {
                return;
        }
        const uint8_t *p = src + (long) (by * 4) * pitch + gx * (8 * BPT);
        uint32_t w[4][2 * BPT];
#pragma unroll
        for (int y = 0; y < 4; ++y, p += pitch) {
                if (BPT == 2) {
                        const uint4 v = ld_stream_v4(p);
                        w[y][0] = v.x, w[y][1] = v.y, w[y][2] = v.z, w[y][3] = v.w;
                } else {
                        const uint2 v = ld_stream_v2(p);
                        w[y][0] = v.x, w[y][1] = v.y;
                }
        }
        out_t res[BPT];
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
                if constexpr (DXT_TYPE == 1) {
                        const uint32_t wk[4][2] = { { w[0][2 * k], w[0][2 * k + 1] }, { w[1][2 * k], w[1][2 * k + 1] },
                                                    { w[2][2 * k], w[2][2 * k + 1] }, { w[3][2 * k], w[3][2 * k + 1] } };
                        res[k] = dxt1_encode_uyvy_packed<BRANCH>(wk);
                } else {
                        float r[16], g[16], b[16];
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                                load_row_uyvy(w[y][2 * k], w[y][2 * k + 1], r + 4 * y, g + 4 * y, b + 4 * y);
                        }
                        res[k] = encode_block<DXT_TYPE>(r, g, b);
                }
        }
        out_t *o = (out_t *) out + ((long) by * wb + gx * BPT);
        if (DXT_TYPE == 1 && BPT == 2) {
                *(uint4 *) o = make_uint4(((uint2 *) res)[0].x, ((uint2 *) res)[0].y, ((uint2 *) res)[1].x, ((uint2 *) res)[1].y);
        } else {
#pragma unroll
                for (int k = 0; k < BPT; ++k) {
                        o[k] = res[k];
                }
        }
}


__device__ __forceinline__ void spin_cycles(long long n)
{
        const long long t0 = clock64();
        while (clock64() - t0 < n) {
        }
}

/// one-shot kernel with a start skew: the CTAs of the first wave wait a pseudo-random time in [0, SKEW) cycles -
/// the CTAs that follow inherit the offsets, so that the warps of an SM are spread over the phases of the block encode instead of marching
/// through them together (conversion = FMA + ALU, bounding box = ALU only, covariance / projection = FMA only)
template <int DXT_TYPE, int BPT, int TPB, int MINB, int SKEW>
__global__ void __launch_bounds__(TPB, MINB) exp_skew_kernel(const uint8_t *__restrict__ src, void *__restrict__ out, int wb, int h, long pitch)
{
        typedef typename block_out<DXT_TYPE>::type out_t;
        const int lin = blockIdx.y * gridDim.x + blockIdx.x;
        if (lin < 148 * MINB) {
                spin_cycles((long long) (((lin * 2654435761u) >> 16) * (unsigned) SKEW >> 16));  // uniform in [0, SKEW)
        }
        const int gx = blockIdx.x * blockDim.x + threadIdx.x;
        const int by = blockIdx.y;
        if (gx >= wb / BPT) {
                return;
        }
        const uint8_t *p = src + (long) (by * 4) * pitch + gx * (8 * BPT);
        uint32_t w[4][2 * BPT];
#pragma unroll
        for (int y = 0; y < 4; ++y, p += pitch) {
                if (BPT == 2) {
                        const uint4 v = ld_stream_v4(p);
                        w[y][0] = v.x, w[y][1] = v.y, w[y][2] = v.z, w[y][3] = v.w;
                } else {
                        const uint2 v = ld_stream_v2(p);
                        w[y][0] = v.x, w[y][1] = v.y;
                }
        }
        out_t res[BPT];
#pragma unroll
        for (int k = 0; k < BPT; ++k) {
                if constexpr (DXT_TYPE == 1) {
                        const uint32_t wk[4][2] = { { w[0][2 * k], w[0][2 * k + 1] }, { w[1][2 * k], w[1][2 * k + 1] },
                                                    { w[2][2 * k], w[2][2 * k + 1] }, { w[3][2 * k], w[3][2 * k + 1] } };
                        res[k] = dxt1_encode_uyvy_packed<true>(wk);
                } else {
                        float r[16], g[16], b[16];
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
                                load_row_uyvy(w[y][2 * k], w[y][2 * k + 1], r + 4 * y, g + 4 * y, b + 4 * y);
                        }
                        res[k] = encode_block<DXT_TYPE>(r, g, b);
                }
        }
        out_t *o = (out_t *) out + ((long) by * wb + gx * BPT);
        if (DXT_TYPE == 1 && BPT == 2) {
                *(uint4 *) o = make_uint4(((uint2 *) res)[0].x, ((uint2 *) res)[0].y, ((uint2 *) res)[1].x, ((uint2 *) res)[1].y);
        } else {
#pragma unroll
                for (int k = 0; k < BPT; ++k) {
                        o[k] = res[k];
                }
        }
}

/// persistent kernel: 148 * MINB CTAs; a warp takes items (32 * BPT horizontally adjacent blocks of one block row) round-robin
/// (DYN = false) or from an atomic counter (DYN = true), and fetches the next item's rows into registers before it encodes the current one.
/// SKEW: every warp starts after a pseudo-random delay in [0, SKEW) cycles.
template <int DXT_TYPE, int BPT, int TPB, int MINB, bool DYN, int SKEW, int PF = 1>
__global__ void __launch_bounds__(TPB, MINB) exp_persist_kernel(const uint8_t *__restrict__ src, void *__restrict__ out, int wb, int h, long pitch,
                                                                 unsigned *__restrict__ counter)
{
        typedef typename block_out<DXT_TYPE>::type out_t;
        constexpr int WPC = TPB / 32;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int groups_per_row = (wb / BPT + 31) / 32;
        const int nitems = groups_per_row * (h / 4);
        const int total_warps = gridDim.x * WPC;
        if (SKEW > 0) {  // uniform in [0, SKEW)
                spin_cycles((long long) ((((blockIdx.x * WPC + warp) * 2654435761u) >> 16) * (unsigned) SKEW >> 16));
        }
        auto next_item = [&](int cur) -> int {
                if (DYN) {
                        unsigned v = 0;
                        if (lane == 0) {
                                v = atomicAdd(counter, 1u);
                        }
                        return (int) __shfl_sync(0xffffffffu, v, 0);
                }
                return cur < 0 ? blockIdx.x * WPC + warp : cur + total_warps;
        };
        auto load_item = [&](int item, uint32_t (&w)[4][2 * BPT]) {
                const int by = item / groups_per_row, gx = (item - by * groups_per_row) * 32 + lane;
                const bool ok = gx < wb / BPT;
                const uint8_t *p = src + (long) (by * 4) * pitch + (ok ? gx : 0) * (8 * BPT);
#pragma unroll
                for (int y = 0; y < 4; ++y, p += pitch) {
                        if (BPT == 2) {
                                const uint4 v = ld_stream_v4(p);
                                w[y][0] = v.x, w[y][1] = v.y, w[y][2] = v.z, w[y][3] = v.w;
                        } else {
                                const uint2 v = ld_stream_v2(p);
                                w[y][0] = v.x, w[y][1] = v.y;
                        }
                }
        };
        auto prefetch_item = [&](int item) {  // PF == 2: pull the next item's rows into L2 while this one is encoded (no registers held)
                const int by = item / groups_per_row, gx = (item - by * groups_per_row) * 32 + lane;
                const uint8_t *p = src + (long) (by * 4) * pitch + (gx < wb / BPT ? gx : 0) * (8 * BPT);
#pragma unroll
                for (int y = 0; y < 4; ++y, p += pitch) {
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
        };
        int item = next_item(-1);
        uint32_t wn[4][2 * BPT];
        if (PF == 1 && item < nitems) {
                load_item(item, wn);
        }
        while (item < nitems) {
                uint32_t w[4][2 * BPT];
                if (PF == 1) {
#pragma unroll
                        for (int y = 0; y < 4; ++y) {
#pragma unroll
                                for (int k = 0; k < 2 * BPT; ++k) {
                                        w[y][k] = wn[y][k];
                                }
                        }
                } else {
                        load_item(item, w);
                }
                const int nxt = next_item(item);
                if (nxt < nitems) {
                        if (PF == 1) {
                                load_item(nxt, wn);
                        } else if (PF == 2) {
                                prefetch_item(nxt);
                        }
                }
                out_t res[BPT];
#pragma unroll
                for (int k = 0; k < BPT; ++k) {
                        if constexpr (DXT_TYPE == 1) {
                                const uint32_t wk[4][2] = { { w[0][2 * k], w[0][2 * k + 1] }, { w[1][2 * k], w[1][2 * k + 1] },
                                                            { w[2][2 * k], w[2][2 * k + 1] }, { w[3][2 * k], w[3][2 * k + 1] } };
                                res[k] = dxt1_encode_uyvy_packed<true>(wk);
                        } else {
                                float r[16], g[16], b[16];
#pragma unroll
                                for (int y = 0; y < 4; ++y) {
                                        load_row_uyvy(w[y][2 * k], w[y][2 * k + 1], r + 4 * y, g + 4 * y, b + 4 * y);
                                }
                                res[k] = encode_block<DXT_TYPE>(r, g, b);
                        }
                }
                const int by = item / groups_per_row, gx = (item - by * groups_per_row) * 32 + lane;
                if (gx < wb / BPT) {
                        out_t *o = (out_t *) out + ((long) by * wb + gx * BPT);
                        if (DXT_TYPE == 1 && BPT == 2) {
                                *(uint4 *) o = make_uint4(((uint2 *) res)[0].x, ((uint2 *) res)[0].y, ((uint2 *) res)[1].x, ((uint2 *) res)[1].y);
                        } else {
#pragma unroll
                                for (int k = 0; k < BPT; ++k) {
                                        o[k] = res[k];
                                }
                        }
                }
                item = nxt;
        }
}

__global__ void fill_kernel(uint32_t *p, long nwords, uint32_t seed, int w_words, int h)
{
        for (long i = (long) blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (long) gridDim.x * blockDim.x) {
                uint32_t x = (uint32_t) i * 2654435761u ^ seed;
                x ^= x << 13, x ^= x >> 17, x ^= x << 5;
                x *= 0x9E3779B1u;
                x ^= x >> 15;
                const int row = (int) (i / w_words);
                if (seed == 1u) {  // frame 0 also carries flat and smooth regions (flat-block path, near-flat blocks)
                        if (row < h / 16) {
                                x = 0x80808080u;
                        } else if (row < h / 8) {
                                const uint32_t v = (uint32_t) ((i % w_words) * 255 / w_words);
                                x = 0x80008000u | v << 8 | ((v + (row & 1)) & 0xff) << 24;
                        } else if (row < h / 4) {
                                x = (x & 0x03030303u) + 0x40804080u;  // low-amplitude noise
                        }
                }
                p[i] = x;
        }
}

}  // namespace ugb

struct variant {
        std::string name;
        int dxt, bpt, tpb, minb;  // minb > 0: persistent kernel with 148 * minb CTAs
        void (*kern)(const uint8_t *, void *, int, int, long);
        void (*pkern)(const uint8_t *, void *, int, int, long, unsigned *);
};

#define V(D, B, T, M, BR) { "d" #D "_b" #B "_t" #T "_m" #M "_" #BR, D, B, T, 0, ugb::exp_kernel<D, B, T, M, BR>, nullptr }
#define S(D, B, T, M, SK) { "d" #D "_b" #B "_t" #T "_m" #M "_skew" #SK, D, B, T, 0, ugb::exp_skew_kernel<D, B, T, M, SK>, nullptr }
#define P(D, B, T, M, DYN, SK) { "d" #D "_b" #B "_t" #T "_m" #M "_persist_" #DYN "_skew" #SK, D, B, T, M, nullptr, ugb::exp_persist_kernel<D, B, T, M, DYN, SK> }
#define Q(D, B, T, M, DYN, SK, PF) { "d" #D "_b" #B "_t" #T "_m" #M "_persist_" #DYN "_skew" #SK "_pf" #PF, D, B, T, M, nullptr, ugb::exp_persist_kernel<D, B, T, M, DYN, SK, PF> }

int main(int argc, char **argv)
{
        const int W = 7680, H = 4320;
        const long frame = (long) W * H * 2;
        std::vector<variant> vs = {
                // DXT1: shipped shape first (two blocks per thread, 64-thread CTAs), then the alternatives that were measured
                V(1, 2, 64, 12, true), V(1, 2, 128, 6, true), V(1, 2, 32, 24, true), V(1, 2, 256, 3, true), V(1, 2, 64, 12, false), V(1, 2, 64, 10, true),
                V(1, 1, 64, 14, true), V(1, 1, 128, 8, true), S(1, 2, 64, 12, 2000), S(1, 2, 64, 12, 8000),
                P(1, 2, 128, 5, false, 1400), P(1, 2, 128, 5, true, 0), Q(1, 2, 128, 6, true, 1400, 2), Q(1, 2, 128, 6, true, 1400, 0),
                // DXT5-YCoCg
                V(6, 1, 128, 6, true), V(6, 1, 64, 12, true), V(6, 1, 128, 5, true), V(6, 1, 128, 4, true), V(6, 1, 128, 7, true), V(6, 1, 256, 3, true),
                S(6, 1, 64, 12, 4000), P(6, 1, 128, 4, true, 1800), Q(6, 1, 128, 6, true, 1800, 2),
        };
        const char *only = argc > 2 && !strcmp(argv[1], "one") ? argv[2] : nullptr;

        uint8_t *src[4], *out, *ref[2];
        unsigned *counters;
        for (int i = 0; i < 4; ++i) {
                cudaMalloc(&src[i], frame + 256);
                ugb::fill_kernel<<<148 * 8, 256>>>((uint32_t *) src[i], frame / 4, (uint32_t) (i + 1), W * 2 / 4, H);
        }
        cudaMalloc(&out, (size_t) W * H);
        cudaMalloc(&ref[0], (size_t) W * H);
        cudaMalloc(&ref[1], (size_t) W * H);
        cudaMalloc(&counters, 4096 * sizeof(unsigned));
        std::vector<uint8_t> h_ref[2], h_out((size_t) W * H);
        // reference outputs of frame 0 from the shipped entry points
        ugb200_uyvy_to_dxt1_async(src[0], ref[0], W, H, 0, nullptr);
        ugb200_uyvy_to_dxt6_async(src[0], ref[1], W, H, 0, nullptr);
        if (cudaDeviceSynchronize() != cudaSuccess) {
                printf("setup failed: %s\n", cudaGetErrorString(cudaGetLastError()));
                return 1;
        }
        for (int k = 0; k < 2; ++k) {
                h_ref[k].resize((size_t) W * H / (k == 0 ? 2 : 1));
                cudaMemcpy(h_ref[k].data(), ref[k], h_ref[k].size(), cudaMemcpyDeviceToHost);
        }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        const int wb = W / 4, hb = H / 4;
        for (const variant &v : vs) {
                if (only && v.name != only) {
                        continue;
                }
                const int groups = wb / v.bpt;
                const dim3 grid((groups + v.tpb - 1) / v.tpb, hb);
                const void *fn = v.kern ? (const void *) v.kern : (const void *) v.pkern;
                int launch_no = 0;
                auto launch = [&](const uint8_t *s) {
                        if (v.kern) {
                                v.kern<<<grid, v.tpb>>>(s, out, wb, H, (long) W * 2);
                        } else {
                                v.pkern<<<148 * v.minb, v.tpb>>>(s, out, wb, H, (long) W * 2, counters + (launch_no++ & 4095));
                        }
                };
                cudaFuncAttributes fa;
                cudaFuncGetAttributes(&fa, fn);
                int occ = 0;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, v.tpb, 0);
                cudaMemset(counters, 0, 4096 * sizeof(unsigned));
                if (only) {
                        for (int i = 0; i < 3; ++i) {
                                launch(src[i]);
                        }
                        cudaDeviceSynchronize();
                        printf("%s ran\n", v.name.c_str());
                        continue;
                }
                const size_t out_bytes = (size_t) W * H / (v.dxt == 1 ? 2 : 1);
                cudaMemset(out, 0xEE, out_bytes);
                launch(src[0]);
                cudaMemcpy(h_out.data(), out, out_bytes, cudaMemcpyDeviceToHost);
                const bool same = memcmp(h_out.data(), h_ref[v.dxt == 1 ? 0 : 1].data(), out_bytes) == 0;
                float best = 1e9f, sum = 0;
                const int reps = 3, iters = v.dxt == 1 ? 40 : 20;
                for (int r = 0; r < reps; ++r) {
                        cudaMemset(counters, 0, 4096 * sizeof(unsigned));
                        launch_no = 0;
                        cudaDeviceSynchronize();
                        cudaEventRecord(e0);
                        for (int i = 0; i < iters; ++i) {
                                launch(src[i & 3]);
                        }
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        float ms;
                        cudaEventElapsedTime(&ms, e0, e1);
                        best = ms < best ? ms : best;
                        sum += ms;
                }
                const cudaError_t err = cudaGetLastError();
                printf("%-36s regs %3d spill %3zu B  CTAs/SM %2d warps/SM %2d  best %7.2f us  mean %7.2f us  %s%s\n", v.name.c_str(), fa.numRegs,
                       (size_t) fa.localSizeBytes, occ, occ * v.tpb / 32, best / iters * 1e3, sum / reps / iters * 1e3, same ? "bit-exact" : "MISMATCH",
                       err == cudaSuccess ? "" : cudaGetErrorString(err));
                fflush(stdout);
        }
        return 0;
}
