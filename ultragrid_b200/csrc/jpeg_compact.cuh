// The per-segment routine of the JPEG stream compaction (jpeg_compact_kernel in jpeg_kernels.cu), kept in a header of its own so that the very
// same code runs on the CPU in tests/test_device_identities.py (through tools/exp_compact.cu) and on the GPU in the product and in the harness.
#pragma once
#include <stdint.h>

namespace ugb {

__host__ __device__ inline uint32_t slot_byte(const uint32_t *src32, uint32_t i) { return (src32[i >> 2] >> (8 * (i & 3))) & 0xffu; }

/// Lane `lane` of `nlanes` moves its share of one restart segment: n bytes from the 4-byte aligned slot `src32` to `dst`, which has any
/// alignment.  The bytes in front of the first aligned word of the stream and behind the last one go out as bytes, everything between as
/// aligned 32-bit words assembled from two aligned slot words by a funnel shift.  Interior words belong to one segment only, so neighbouring
/// segments never write the same word; the second slot word read lies at most one word behind the last byte (inside the slot's spare bytes).
__host__ __device__ inline void compact_segment(int lane, int nlanes, const uint32_t *src32, uint32_t n, uint8_t *dst)
{
        const uint32_t mis = (uint32_t) ((size_t) dst & 3u);
        const uint32_t head = mis ? (4u - mis < n ? 4u - mis : n) : 0u;  // bytes in front of the first aligned word of the stream
        if ((uint32_t) lane < head) {
                dst[lane] = (uint8_t) slot_byte(src32, (uint32_t) lane);
        }
        const uint32_t body = (n - head) >> 2;  // whole aligned words
        uint32_t *dw = (uint32_t *) (dst + head);
        const uint32_t bs = 8u * (head & 3u);   // the slot runs `head` bytes ahead of a word boundary: the same shift for every word
        for (uint32_t j = (uint32_t) lane; j < body; j += (uint32_t) nlanes) {
                const uint32_t wi = (head + 4u * j) >> 2;
                const uint32_t lo = src32[wi];
#ifdef __CUDA_ARCH__
                dw[j] = bs ? __funnelshift_r(lo, src32[wi + 1], bs) : lo;
#else
                dw[j] = bs ? (lo >> bs) | (src32[wi + 1] << (32u - bs)) : lo;
#endif
        }
        const uint32_t done = head + 4u * body;
        if ((uint32_t) lane < n - done) {
                dst[done + lane] = (uint8_t) slot_byte(src32, done + (uint32_t) lane);
        }
}

}  // namespace ugb
