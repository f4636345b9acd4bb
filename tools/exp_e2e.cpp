// Experiment / demo harness (not product code): frames per second through the compress modules with PINNED HOST frames, plain C ABI only
// (include/ugb200_vcompress.h + cuda_wrapper.h) - what UltraGrid's sender does with compress_frame / compress_pop.
// Build: g++ -O2 -std=c++17 -o tools/exp_e2e tools/exp_e2e.cpp -Lultragrid_b200 -lugb200 -Wl,-rpath,'$ORIGIN/../ultragrid_b200' -pthread
// Run:   tools/exp_e2e [frames]
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../include/cuda_wrapper.h"
#include "../include/ugb200_jpeg.h"
#include "../include/ugb200_vcompress.h"

static double now_s()
{
        return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
        const int W = 7680, H = 4320, UYVY = 2;
        const int n = argc > 1 && atoi(argv[1]) > 0 ? atoi(argv[1]) : 48;
        const size_t frame = (size_t) W * H * 2;
        uint8_t *host[3];
        for (int k = 0; k < 3; ++k) {
                if (cuda_wrapper_malloc_host((void **) &host[k], frame) != CUDA_WRAPPER_SUCCESS) {
                        fprintf(stderr, "pinned allocation failed\n");
                        return 1;
                }
                uint32_t x = 12345u + k;
                for (int y = 0; y < H; ++y) {  // smooth ramps + a few levels of noise ("natural"): a q=90 stream of several MB
                        uint8_t *row = host[k] + (size_t) y * W * 2;
                        for (int i = 0; i < W * 2; i += 4) {
                                x ^= x << 13, x ^= x >> 17, x ^= x << 5;
                                const int px = i / 2;
                                row[i + 0] = (uint8_t) (128 + ((px * 64) / W - 32) + (x & 3));
                                row[i + 1] = (uint8_t) (16 + (px * 200) / W + ((x >> 8) % 13));
                                row[i + 2] = (uint8_t) (128 + ((y * 64) / H - 32) + ((x >> 4) & 3));
                                row[i + 3] = (uint8_t) (16 + (px * 200) / W + ((x >> 16) % 13));
                        }
                }
        }
        if (argc > 1 && !strcmp(argv[1], "jpeg")) {  // device-resident encode, kernels only (target for ncu)
                void *dev;
                cuda_wrapper_stream_t st;
                if (cuda_wrapper_malloc(&dev, frame + 64) != 0 || cuda_wrapper_memcpy(dev, host[0], frame, CUDA_WRAPPER_MEMCPY_HOST_TO_DEVICE) != 0 ||
                    cuda_wrapper_stream_create(&st) != 0) {
                        return 1;
                }
                ugb200_jpeg_encoder *e = ugb200_jpeg_encoder_create(st);
                struct ugb200_jpeg_params p;
                ugb200_jpeg_default_params(&p);
                p.quality = 90;
                size_t sz = 0;
                for (int rep = 0; rep < 3; ++rep) {
                        const double t0 = now_s();
                        const int iters = rep == 0 ? 1 : 20;
                        for (int i = 0; i < iters; ++i) {
                                if (ugb200_jpeg_encode_device(e, dev, 0, W, H, UYVY, &p) != 0) {
                                        return 2;
                                }
                        }
                        if (ugb200_jpeg_result_device(e, nullptr, &sz) != 0) {
                                return 3;
                        }
                        printf("jpeg device encode: %.1f us/frame, stream %zu bytes\n", (now_s() - t0) / iters * 1e6, sz);
                }
                ugb200_jpeg_encoder_destroy(e);
                return 0;
        }
        const struct {
                const char *cfg;
                int depth;
        } runs[] = { { "cuda_dxt:DXT1", 3 }, { "GPUJPEG:q=90:lanes=1", 1 }, { "GPUJPEG:q=90:lanes=2", 2 }, { "GPUJPEG:q=90", 3 }, { "GPUJPEG:q=90:lanes=4", 4 } };
        for (const auto &r : runs) {
                ugb200_compress *c = ugb200_compress_init(r.cfg);
                if (!c) {
                        printf("%s: init failed\n", r.cfg);
                        continue;
                }
                size_t len = 0, first_len = 0;
                uint64_t sum0 = 0, sum = 0;
                bool ok = true;
                auto pass = [&](int frames, bool check) {
                        int inflight = 0;
                        auto pop = [&]() {
                                const void *data;
                                int codec;
                                unsigned seq;
                                if (ugb200_compress_pop_ref(c, &data, &len, &codec, &seq) != 0) {
                                        ok = false;
                                        return;
                                }
                                if (check) {  // same input every third frame -> same stream
                                        sum = 0;
                                        for (size_t i = 0; i < len; i += 4096) {
                                                sum += ((const uint8_t *) data)[i];
                                        }
                                        if (seq % 3 == 0) {
                                                if (first_len == 0) {
                                                        first_len = len, sum0 = sum;
                                                } else if (len != first_len || sum != sum0) {
                                                        ok = false;
                                                }
                                        }
                                }
                                --inflight;
                        };
                        for (int i = 0; i < frames; ++i) {
                                if (ugb200_compress_push(c, host[i % 3], 0, W, H, UYVY, 60.0) != 0) {
                                        ok = false;
                                }
                                if (++inflight == r.depth) {
                                        pop();
                                }
                        }
                        while (inflight > 0 && ok) {
                                pop();
                        }
                };
                pass(9, true);  // buffers, pool frames, every lane
                const double t0 = now_s();
                pass(n, false);
                const double dt = now_s() - t0;
                printf("%-24s %7.1f frames/s  (%d frames, %.1f ms each, H2D %.1f GB/s, last stream %zu bytes) %s\n", r.cfg, n / dt, n, dt / n * 1e3,
                       frame * n / dt / 1e9, len, ok ? "ok" : "FAILED");
                fflush(stdout);
                ugb200_compress_done(c);
        }
        return 0;
}
