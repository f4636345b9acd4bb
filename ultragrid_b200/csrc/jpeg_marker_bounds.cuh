// The part of the JPEG decoder's device marker scan that decides about multi-scan streams, in a header of its own so that the very same code runs on the
// GPU (jpeg_marker_bounds_kernel / jpeg_marker_segments_multi_kernel in jpeg_decode_kernels.cu) and on the CPU in tests/test_jpeg_marker_bounds.py, where it
// is compared with the host parser (ugb200_jpeg_debug_segments) on intact, damaged and randomly mutated streams.
#pragma once
#include <stdint.h>

#ifndef __CUDACC__
#define UGB_HD
#else
#define UGB_HD __host__ __device__
#endif

namespace ugb {

// layout of the marker scan's `meta` words on the device (jpeg_marker_*_kernel)
constexpr int kMetaTotal = 0, kMetaFirstOther = 1, kMetaOtherCount = 2, kMetaOther = 3, kMaxOther = 8, kMetaError = 11, kMetaBounds = 12, kMetaWords = 32;

/// Streams with one scan PER COMPONENT (RGB as GPUJPEG stores it, gpujpeg.cpp:303-305): the SOS headers of scans 2 and 3 lie behind entropy-coded data, where
/// the host does not look any more.  One thread walks the few candidates that are not RSTn: each scan's data runs from behind its SOS header to the next such
/// marker; that marker must be the next scan's SOS (one component, the expected component id, baseline table selectors), the last scan must end in EOI or with
/// the stream.  bounds[j] = { first list index, RSTn count, data begin, terminator position, Td, Ta }.  Anything else - tables redefined between the scans,
/// another component order, more markers than expected - sets meta[kMetaError]: the host then repeats the frame with its own parser, which defines what
/// happens to irregular and damaged streams.  `list` = positions of all marker candidates behind begin0, ascending; meta[kMetaTotal] their number,
/// meta[kMetaOtherCount] / meta[kMetaOther ..] the number / list indices (any order) of those that are not RSTn.
UGB_HD inline void marker_bounds(const uint8_t *s, uint32_t len, const uint32_t *list, uint32_t *meta, uint32_t begin0, int nscans, uint32_t comp_ids)
{
        const uint32_t total = meta[kMetaTotal];
        uint32_t n = meta[kMetaOtherCount];
        if (n > (uint32_t) kMaxOther) {
                meta[kMetaError] = 1;
                return;
        }
        uint32_t idx[kMaxOther];
        for (uint32_t i = 0; i < n; ++i) {  // the atomics hand them out in any order
                uint32_t v = meta[kMetaOther + i], j = i;
                for (; j > 0 && idx[j - 1] > v; --j) {
                        idx[j] = idx[j - 1];
                }
                idx[j] = v;
        }
        uint32_t begin = begin0, k = 0, err = 0;
        for (int j = 0; j < nscans; ++j) {
                uint32_t lo = 0, hi = total;  // first candidate at or behind `begin`
                while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (list[mid] < begin) {
                                lo = mid + 1;
                        } else {
                                hi = mid;
                        }
                }
                while (k < n && idx[k] < lo) {
                        ++k;
                }
                const uint32_t term_idx = k < n ? idx[k] : total, term = k < n ? list[term_idx] : len;
                uint32_t *b = meta + kMetaBounds + 6 * j;
                b[0] = lo, b[1] = term_idx - lo, b[2] = begin, b[3] = term;
                if (j == nscans - 1) {
                        if (k < n && s[term + 1] != 0xD9) {
                                err = 1;  // something follows the last scan that is not EOI
                        }
                        break;
                }
                if (k >= n || term + 10 > len || s[term + 1] != 0xDA || s[term + 2] != 0 || s[term + 3] != 8 || s[term + 4] != 1 ||
                    s[term + 5] != ((comp_ids >> (8 * (j + 1))) & 0xff) || (s[term + 6] >> 4) > 1 || (s[term + 6] & 15) > 1) {
                        err = 1;
                        break;
                }
                b[6 + 4] = s[term + 6] >> 4, b[6 + 5] = s[term + 6] & 15;  // Td, Ta of scan j + 1
                begin = term + 10;
                ++k;
        }
        meta[kMetaError] = err;
}

/// restart segment `i` (0 <= i < nseg) of a multi-scan stream from the bounds above: [begin, end) of its entropy-coded bytes - per scan the rules of the host
/// parser's SOS branch (excess RSTn are ignored, the last segment runs to the scan's terminator, missing segments are empty at the terminator)
UGB_HD inline void marker_segment_multi(const uint32_t *list, const uint32_t *meta, int nscans, int seg1, int seg2, int nseg, int i, uint32_t *seg_b, uint32_t *seg_e)
{
        const int j = nscans > 2 && i >= seg2 ? 2 : nscans > 1 && i >= seg1 ? 1 : 0;
        const int first = j == 2 ? seg2 : j == 1 ? seg1 : 0, last = j == 0 ? (nscans > 1 ? seg1 : nseg) : j == 1 ? (nscans > 2 ? seg2 : nseg) : nseg;
        const int li = i - first, n = last - first;
        const uint32_t *b = meta + kMetaBounds + 6 * j;
        const uint32_t lo = b[0], stop = b[1], begin0 = b[2], term = b[3];
        const uint32_t pushed = stop < (uint32_t) (n - 1) ? stop : (uint32_t) (n - 1);
        if ((uint32_t) li < pushed) {
                *seg_b = li == 0 ? begin0 : list[lo + li - 1] + 2, *seg_e = list[lo + li];
        } else if ((uint32_t) li == pushed) {
                *seg_b = stop == 0 ? begin0 : list[lo + stop - 1] + 2, *seg_e = term;
        } else {
                *seg_b = *seg_e = term;
        }
}

}  // namespace ugb
