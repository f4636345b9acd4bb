// Instruction-throughput microbenchmark for the ops the DXT/JPEG encoders are made of (sm_100a).
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/ubench tools/ubench.cu ; run on the GPU box.
// Prints lane-ops per clock per SM for each op (128 = one warp instruction per SMSP per clock).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

#define KERNEL(name, decl, body, fin)                                                                                  \
        __global__ void name(uint32_t *out, float seed)                                                                \
        {                                                                                                              \
                decl;                                                                                                  \
                for (int it = 0; it < ITERS; ++it) {                                                                   \
                        _Pragma("unroll") for (int k = 0; k < ILP; ++k) { body; }                                      \
                }                                                                                                      \
                uint32_t acc = 0;                                                                                      \
                _Pragma("unroll") for (int k = 0; k < ILP; ++k) { fin; }                                               \
                if (acc == 0x12345678u) out[threadIdx.x] = acc;                                                        \
        }

KERNEL(k_ffma, float a[ILP]; float b = seed; float c = seed * 0.5f; for (int k = 0; k < ILP; ++k) a[k] = seed + k,
       asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b), "f"(c)), acc += __float_as_uint(a[k]))
KERNEL(k_ffma_imm, float a[ILP]; for (int k = 0; k < ILP; ++k) a[k] = seed + k,
       asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f3F000000;" : "+f"(a[k])), acc += __float_as_uint(a[k]))
KERNEL(k_ffma2, unsigned long long a[ILP]; unsigned long long b = ((unsigned long long) __float_as_uint(seed) << 32) | __float_as_uint(seed);
       for (int k = 0; k < ILP; ++k) a[k] = b + k,
       asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[k]) : "l"(b)), acc += (uint32_t) a[k])
KERNEL(k_fmul2, unsigned long long a[ILP]; unsigned long long b = ((unsigned long long) __float_as_uint(seed) << 32) | __float_as_uint(seed);
       for (int k = 0; k < ILP; ++k) a[k] = b + k,
       asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(a[k]) : "l"(b)), acc += (uint32_t) a[k])
KERNEL(k_fmnmx3, float a[ILP]; float b = seed; float c = seed * 0.5f; for (int k = 0; k < ILP; ++k) a[k] = seed + k,
       asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[k]) : "f"(b), "f"(c)), acc += __float_as_uint(a[k]))
KERNEL(k_fmnmx, float a[ILP]; float b = seed; for (int k = 0; k < ILP; ++k) a[k] = seed + k,
       asm volatile("min.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b)), acc += __float_as_uint(a[k]))
KERNEL(k_prmt, uint32_t a[ILP]; uint32_t b = __float_as_uint(seed); for (int k = 0; k < ILP; ++k) a[k] = b + k,
       asm volatile("prmt.b32 %0, %0, %1, 0x7541;" : "+r"(a[k]) : "r"(b)), acc += a[k])
KERNEL(k_i2f, uint32_t a[ILP]; for (int k = 0; k < ILP; ++k) a[k] = __float_as_uint(seed) + k,
       { float f; asm volatile("cvt.rn.f32.u32 %0, %1;" : "=f"(f) : "r"(a[k] & 0xff)); a[k] = __float_as_uint(f); }, acc += a[k])
KERNEL(k_f2i, float a[ILP]; for (int k = 0; k < ILP; ++k) a[k] = seed + k,
       { uint32_t u; asm volatile("cvt.rzi.u32.f32 %0, %1;" : "=r"(u) : "f"(a[k])); a[k] = __uint_as_float(u | 0x3f800000u); }, acc += __float_as_uint(a[k]))
KERNEL(k_fadd_rm, float a[ILP]; float b = seed; for (int k = 0; k < ILP; ++k) a[k] = seed + k,
       asm volatile("add.rm.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b)), acc += __float_as_uint(a[k]))
KERNEL(k_fadd_sat, float a[ILP]; float b = seed; for (int k = 0; k < ILP; ++k) a[k] = seed + k,
       asm volatile("add.rn.sat.f32 %0, %0, %1;" : "+f"(a[k]) : "f"(b)), acc += __float_as_uint(a[k]))
KERNEL(k_imad, uint32_t a[ILP]; uint32_t b = __float_as_uint(seed); for (int k = 0; k < ILP; ++k) a[k] = b + k,
       asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(a[k]) : "r"(b)), acc += a[k])
KERNEL(k_lea, uint32_t a[ILP]; uint32_t b = __float_as_uint(seed); for (int k = 0; k < ILP; ++k) a[k] = b + k,
       a[k] = (a[k] << 2) + b, acc += a[k])
KERNEL(k_lop3, uint32_t a[ILP]; uint32_t b = __float_as_uint(seed); for (int k = 0; k < ILP; ++k) a[k] = b + k,
       asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(a[k]) : "r"(b)), acc += a[k])
KERNEL(k_dadd, double a[ILP]; double b = seed; for (int k = 0; k < ILP; ++k) a[k] = seed + k,
       asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[k]) : "d"(b)), acc += (uint32_t) __double2loint(a[k]))
KERNEL(k_f2f64, float a[ILP]; for (int k = 0; k < ILP; ++k) a[k] = seed + k,
       { double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(a[k])); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(a[k]) : "d"(d)); }, acc += __float_as_uint(a[k]))
// mixes: one FMA-pipe op + one ALU-pipe op per slot
KERNEL(k_mix_ffma_prmt, float a[ILP]; uint32_t p[ILP]; float b = seed; for (int k = 0; k < ILP; ++k) { a[k] = seed + k; p[k] = k; },
       { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[k]) : "f"(b)); asm volatile("prmt.b32 %0, %0, %1, 0x7541;" : "+r"(p[k]) : "r"(p[(k + 1) % ILP])); },
       acc += __float_as_uint(a[k]) + p[k])
KERNEL(k_mix_ffma2_fmnmx3, unsigned long long a[ILP]; float m[ILP]; float b = seed;
       unsigned long long bb = ((unsigned long long) __float_as_uint(seed) << 32) | __float_as_uint(seed);
       for (int k = 0; k < ILP; ++k) { a[k] = bb + k; m[k] = seed + k; },
       { asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(a[k]) : "l"(bb)); asm volatile("min.f32 %0, %0, %1, %1;" : "+f"(m[k]) : "f"(b)); },
       acc += (uint32_t) a[k] + __float_as_uint(m[k]))

template <typename K>
static void run(const char *name, K kern, double ops_per_iter_slot, int sms, double ghz_hint)
{
        uint32_t *out;
        cudaMalloc(&out, 4096);
        const int blocks = sms * 4, threads = 256;
        kern<<<blocks, threads>>>(out, 1.0f);
        cudaDeviceSynchronize();
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) kern<<<blocks, threads>>>(out, 1.0f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        const double lane_ops = 5.0 * blocks * threads * (double) ITERS * ILP * ops_per_iter_slot;
        const double per_s = lane_ops / (ms * 1e-3);
        printf("%-20s %8.3f ms  %7.2f Tlane-op/s  = %6.1f lane-ops/clk/SM @%.3f GHz\n", name, ms / 5, per_s / 1e12, per_s / sms / (ghz_hint * 1e9), ghz_hint);
        cudaFree(out);
}

int main()
{
        cudaDeviceProp p;
        cudaGetDeviceProperties(&p, 0);
        int clk_khz = 0;
        cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
        const double ghz = clk_khz / 1e6;
        printf("%s, %d SMs, max clock %.3f GHz (lane-ops/clk computed against max clock; real clock may be lower)\n", p.name, p.multiProcessorCount, ghz);
        const int s = p.multiProcessorCount;
        run("ffma r,r,r", k_ffma, 1, s, ghz);
        run("ffma imm", k_ffma_imm, 1, s, ghz);
        run("ffma2 (per lane-pair)", k_ffma2, 1, s, ghz);
        run("fmul2", k_fmul2, 1, s, ghz);
        run("fmnmx3", k_fmnmx3, 1, s, ghz);
        run("fmnmx", k_fmnmx, 1, s, ghz);
        run("prmt", k_prmt, 1, s, ghz);
        run("i2f (+lop)", k_i2f, 1, s, ghz);
        run("f2i (+lop)", k_f2i, 1, s, ghz);
        run("fadd.rm", k_fadd_rm, 1, s, ghz);
        run("fadd.sat", k_fadd_sat, 1, s, ghz);
        run("imad", k_imad, 1, s, ghz);
        run("lea", k_lea, 1, s, ghz);
        run("lop3", k_lop3, 1, s, ghz);
        run("dadd", k_dadd, 1, s, ghz);
        run("f2f 32<->64 (pair)", k_f2f64, 1, s, ghz);
        run("mix ffma+prmt (pairs)", k_mix_ffma_prmt, 1, s, ghz);
        run("mix ffma2+fmnmx3 (pairs)", k_mix_ffma2_fmnmx3, 1, s, ghz);
        return 0;
}
