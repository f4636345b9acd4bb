// Experiment harness (not product code): the JPEG stream compaction step in isolation.  64 800 restart segments of ~100 bytes sit in
// worst-case slots (6 664 bytes apart, as jpeg_kernels.cu lays them out for an 8K UYVY frame); the kernel moves each to its byte offset in
// the stream.  Variants: the earlier form (a warp per segment, byte loads and byte stores) and the shipped one, aligned 32-bit stores fed by a funnel shift of
// two aligned source words (head and tail bytes of a segment as byte stores; interior words belong to exactly one segment, so there is no race).
// The per-segment routine (ultragrid_b200/csrc/jpeg_compact.cuh, shared with the product kernel) is __host__ __device__: `tools/exp_compact check`
// runs it on the CPU against memcpy for random sizes and alignments.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/exp_compact tools/exp_compact.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime.h>

#include "../ultragrid_b200/csrc/jpeg_compact.cuh"  // the product's per-segment routine

template <int LANES>
__global__ void __launch_bounds__(256) compact_aligned_kernel(const uint8_t *__restrict__ slots, const uint32_t *__restrict__ sizes,
                                                              const uint32_t *__restrict__ offs, int nseg, long slot, uint8_t *__restrict__ out)
{
        const int t = blockIdx.x * blockDim.x + threadIdx.x, s = t / LANES, lane = t % LANES;
        if (s >= nseg) {
                return;
        }
        ugb::compact_segment(lane, LANES, (const uint32_t *) (slots + (long) s * slot), sizes[s], out + offs[s]);
}

/// the form shipped until the end of round 1: a warp per segment, bytes
__global__ void __launch_bounds__(256) compact_bytes_kernel(const uint8_t *__restrict__ slots, const uint32_t *__restrict__ sizes,
                                                            const uint32_t *__restrict__ offs, int nseg, long slot, uint8_t *__restrict__ out)
{
        const int s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
        if (s >= nseg) {
                return;
        }
        const uint32_t n = sizes[s], off = offs[s];
        const uint8_t *src = slots + (long) s * slot;
        for (uint32_t i = lane; i < n; i += 32) {
                out[off + i] = src[i];
        }
}

static int cpu_check()
{
        srand(7);
        std::vector<uint32_t> slot(4096 / 4 + 4);
        std::vector<uint8_t> out(8192), want(8192);
        for (int trial = 0; trial < 200000; ++trial) {
                const uint32_t n = trial < 64 ? (uint32_t) trial : (uint32_t) (rand() % 700);
                const uint32_t off = 16 + (uint32_t) (rand() % 9);
                const int lanes = (trial & 1) ? 32 : 8;
                for (auto &w : slot) {
                        w = (uint32_t) rand() * 2654435761u;
                }
                memset(out.data(), 0xEE, out.size());
                memset(want.data(), 0xEE, want.size());
                memcpy(want.data() + off, slot.data(), n);
                for (int lane = 0; lane < lanes; ++lane) {
                        ugb::compact_segment(lane, lanes, slot.data(), n, out.data() + off);
                }
                if (memcmp(out.data(), want.data(), out.size()) != 0) {
                        printf("MISMATCH n=%u off=%u lanes=%d\n", n, off, lanes);
                        return 1;
                }
        }
        printf("cpu check ok: 200000 random segments, no byte outside [off, off + n) touched\n");
        return 0;
}

int main(int argc, char **argv)
{
        if (argc > 1 && !strcmp(argv[1], "check")) {
                return cpu_check();
        }
        const int nseg = 64800;
        const long slot = 4 * 4 * 416 + 8;
        std::vector<uint32_t> sizes(nseg), offs(nseg);
        srand(3);
        uint32_t total = 623;  // header
        for (int s = 0; s < nseg; ++s) {
                sizes[s] = 60 + rand() % 90;
                offs[s] = total;
                total += sizes[s];
        }
        uint8_t *d_slots, *d_out, *d_ref;
        uint32_t *d_sizes, *d_offs;
        cudaMalloc(&d_slots, (size_t) nseg * slot);
        cudaMalloc(&d_out, total + 64);
        cudaMalloc(&d_ref, total + 64);
        cudaMalloc(&d_sizes, nseg * 4);
        cudaMalloc(&d_offs, nseg * 4);
        {
                std::vector<uint8_t> h((size_t) nseg * slot);
                for (size_t i = 0; i < h.size(); i += 4) {
                        *(uint32_t *) &h[i] = (uint32_t) rand() * 2654435761u;
                }
                cudaMemcpy(d_slots, h.data(), h.size(), cudaMemcpyHostToDevice);
        }
        cudaMemcpy(d_sizes, sizes.data(), nseg * 4, cudaMemcpyHostToDevice);
        cudaMemcpy(d_offs, offs.data(), nseg * 4, cudaMemcpyHostToDevice);
        cudaMemset(d_ref, 0, total + 64);
        compact_bytes_kernel<<<(nseg * 32 + 255) / 256, 256>>>(d_slots, d_sizes, d_offs, nseg, slot, d_ref);
        std::vector<uint8_t> ref(total + 64), got(total + 64);
        cudaMemcpy(ref.data(), d_ref, total + 64, cudaMemcpyDeviceToHost);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0), cudaEventCreate(&e1);
        // a 200 MB scratch write between launches keeps the slots and the stream out of L2, as they are after the fused kernel of a real frame
        uint8_t *scratch;
        cudaMalloc(&scratch, 256u << 20);
        for (int v = 0; v < 4; ++v) {
                float sum = 0;
                const int iters = 10;
                bool same = true;
                for (int i = 0; i < iters; ++i) {
                        cudaMemsetAsync(scratch, i, 256u << 20);
                        cudaMemsetAsync(d_out, 0, total + 64);
                        cudaEventRecord(e0);
                        if (v == 0) {
                                compact_bytes_kernel<<<(nseg * 32 + 255) / 256, 256>>>(d_slots, d_sizes, d_offs, nseg, slot, d_out);
                        } else if (v == 1) {
                                compact_aligned_kernel<32><<<(nseg * 32 + 255) / 256, 256>>>(d_slots, d_sizes, d_offs, nseg, slot, d_out);
                        } else if (v == 2) {
                                compact_aligned_kernel<8><<<(nseg * 8 + 255) / 256, 256>>>(d_slots, d_sizes, d_offs, nseg, slot, d_out);
                        } else {
                                compact_aligned_kernel<16><<<(nseg * 16 + 255) / 256, 256>>>(d_slots, d_sizes, d_offs, nseg, slot, d_out);
                        }
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                        float ms;
                        cudaEventElapsedTime(&ms, e0, e1);
                        sum += ms;
                }
                cudaMemcpy(got.data(), d_out, total + 64, cudaMemcpyDeviceToHost);
                same = memcmp(got.data(), ref.data(), total + 64) == 0;
                const char *names[] = { "bytes, warp per segment (earlier)", "aligned words, 32 lanes", "aligned words, 8 lanes (shipped)", "aligned words, 16 lanes" };
                printf("%-36s %7.2f us  %s  (%u bytes, %d segments)\n", names[v], sum / iters * 1e3, same ? "identical" : "MISMATCH", total, nseg);
        }
        return 0;
}
