// video_decompress framework + the B200 decompress modules, host side (g++ only; everything through the C ABI).
//
//   framework        src/video_decompress.c:100-300        best module by priority, thin dispatch
//   gpujpeg          src/video_decompress/gpujpeg.c        JPEG -> RGB / RGBA / UYVY (probe with out_codec NONE)
//   gpujpeg_to_dxt   src/video_decompress/gpujpeg_to_dxt.cpp   JPEG -> RGB on the device -> cuda_rgb_to_dxt{1,6} with mirrored height
//   dxt_cuda         (the reference has only the OpenGL module src/video_decompress/dxt_glsl.c) DXT1 / DXT5-YCoCg -> RGB, RGBA, UYVY
#include "video_decompress.h"

#include <cstdio>
#include <cstring>
#include <string>

#include "../../../include/cuda_dxt.h"
#include "../../../include/ugb200_jpeg.h"
#include "../../../include/ugb200_vcompress.h"
#include "video_codec.h"

// ---- framework --------------------------------------------------------------------------------------------------------------
struct state_decompress {
        const video_decompress_info *functions;
        void *state;
        std::string name;
};

/// find_best_decompress, src/video_decompress.c:100-148: lowest priority value within [prio_min, prio_max]
static int find_best_decompress(codec_t compression, struct pixfmt_desc internal, codec_t to, int prio_min, int prio_max, const video_decompress_info **vdi,
                                std::string *name)
{
        const char *names[32];
        const void *infos[32];
        const int n = get_libraries_for_class(LIBRARY_CLASS_VIDEO_DECOMPRESS, VIDEO_DECOMPRESS_ABI_VERSION, names, infos, 32);
        int best = -1;
        for (int i = 0; i < n; ++i) {
                const auto *info = (const video_decompress_info *) infos[i];
                const int prio = info->get_decompress_priority(compression, internal, to);
                if (prio < 0 || prio < prio_min || prio > prio_max) {
                        continue;
                }
                if (best == -1 || prio < best) {
                        best = prio, *vdi = info, *name = names[i];
                }
        }
        return best;
}

/// decompress_init_multi, src/video_decompress.c:173-230 (count states of the same module, for tiled streams)
bool decompress_init_multi(codec_t compression, struct pixfmt_desc internal, codec_t to, struct state_decompress **out, int count)
{
        int prio_min = 0;
        const int prio_max = 1000;
        for (;;) {
                const video_decompress_info *vdi = nullptr;
                std::string name;
                const int prio = find_best_decompress(compression, internal, to, prio_min, prio_max, &vdi, &name);
                if (prio == -1) {
                        return false;
                }
                int ok = 0;
                for (; ok < count; ++ok) {
                        void *st = vdi->init();
                        if (!st) {
                                break;
                        }
                        out[ok] = new state_decompress{ vdi, st, name };
                }
                if (ok == count) {
                        return true;
                }
                while (ok-- > 0) {
                        decompress_done(out[ok]);
                }
                prio_min = prio + 1;  // failed: try the next best one
        }
}
int decompress_reconfigure(struct state_decompress *s, struct video_desc desc, int rshift, int gshift, int bshift, int pitch, codec_t out_codec)
{
        return s->functions->reconfigure(s->state, desc, rshift, gshift, bshift, pitch, out_codec);
}
decompress_status decompress_frame(struct state_decompress *s, unsigned char *dst, unsigned char *src, unsigned int src_len, int frame_seq,
                                   struct video_frame_callbacks *callbacks, struct pixfmt_desc *internal_prop)
{
        return s->functions->decompress(s->state, dst, src, src_len, frame_seq, callbacks, internal_prop);
}
int decompress_get_property(struct state_decompress *s, int property, void *val, size_t *len) { return s->functions->get_property(s->state, property, val, len); }
void decompress_done(struct state_decompress *s)
{
        s->functions->done(s->state);
        delete s;
}
const char *decompress_module_name(struct state_decompress *s) { return s->name.c_str(); }

static int no_corrupted_frames(void *, int property, void *val, size_t *len)  // gpujpeg.c:325-343 and the others alike
{
        if (property == DECOMPRESS_PROPERTY_ACCEPTS_CORRUPTED_FRAME && *len >= sizeof(int)) {
                *(int *) val = 0, *len = sizeof(int);
                return 1;
        }
        return 0;
}

// ---- gpujpeg ---------------------------------------------------------------------------------------------------------------------
namespace {
struct state_decompress_gpujpeg {  // gpujpeg.c:63-70
        ugb200_jpeg_decoder *decoder = nullptr;
        struct video_desc desc{};
        int rshift = 0, gshift = 0, bshift = 0, pitch = 0;
        codec_t out_codec = VIDEO_CODEC_NONE;
};
}  // namespace

static void *gpujpeg_decompress_init(void)
{
        if (cuda_wrapper_set_device((int) cuda_devices[0]) != CUDA_WRAPPER_SUCCESS) {  // gpujpeg_init_device, gpujpeg.c:163
                fprintf(stderr, "[GPUJPEG dec.] initializing CUDA device %u failed.\n", cuda_devices[0]);
                return nullptr;
        }
        return new state_decompress_gpujpeg();
}
static int gpujpeg_decompress_reconfigure(void *state, struct video_desc desc, int rshift, int gshift, int bshift, int pitch, codec_t out_codec)
{
        auto *s = (state_decompress_gpujpeg *) state;
        if (out_codec != RGB && out_codec != RGBA && out_codec != UYVY && out_codec != I420 && out_codec != VIDEO_CODEC_NONE) {
                return 0;  // the reference asserts this set, gpujpeg.c:181-182
        }
        s->desc = desc, s->rshift = rshift, s->gshift = gshift, s->bshift = bshift, s->pitch = pitch, s->out_codec = out_codec;
        if (!s->decoder) {
                s->decoder = ugb200_jpeg_decoder_create(nullptr);
        }
        // dst holds pitch * desc.height bytes: a stream that declares another size must not be decoded into it
        return s->decoder != nullptr && ugb200_jpeg_decoder_expect(s->decoder, (int) desc.width, (int) desc.height) == 0;
}
/// gpujpeg_probe_internal_codec, gpujpeg.c:205-262
static decompress_status gpujpeg_probe_internal_codec(unsigned char *buffer, size_t len, struct pixfmt_desc *internal_prop)
{
        struct ugb200_jpeg_image_info info;
        if (ugb200_jpeg_get_image_info(buffer, len, &info) != 0) {
                fprintf(stderr, "[GPUJPEG dec.] probe - cannot get image info!\n");
                return DECODER_NO_FRAME;
        }
        internal_prop->depth = 8;
        internal_prop->rgb = info.native_codec == RGB;
        internal_prop->subsampling = info.h_samp == 1 ? SUBS_444 : info.v_samp == 1 ? SUBS_422 : SUBS_420;
        return DECODER_GOT_CODEC;
}
static decompress_status gpujpeg_decompress(void *state, unsigned char *dst, unsigned char *buffer, unsigned int src_len, int, struct video_frame_callbacks *,
                                            struct pixfmt_desc *internal_prop)
{
        auto *s = (state_decompress_gpujpeg *) state;
        if (s->out_codec == VIDEO_CODEC_NONE) {
                return gpujpeg_probe_internal_codec(buffer, src_len, internal_prop);
        }
        cuda_wrapper_set_device((int) cuda_devices[0]);
        // the device path writes any pitch and any RGBA shifts directly (the reference needs a second CPU pass for those, gpujpeg.c:295-318)
        const int rc = ugb200_jpeg_decode(s->decoder, buffer, src_len, dst, 0, s->pitch, s->out_codec, s->rshift, s->gshift, s->bshift);
        return rc == 0 ? DECODER_GOT_FRAME : DECODER_NO_FRAME;
}
static void gpujpeg_decompress_done(void *state)
{
        auto *s = (state_decompress_gpujpeg *) state;
        ugb200_jpeg_decoder_destroy(s->decoder);
        delete s;
}
static int gpujpeg_decompress_get_priority(codec_t compression, struct pixfmt_desc, codec_t ugc)  // gpujpeg.c:355-367
{
        if (compression != JPEG) {
                return -1;
        }
        if (ugc == VIDEO_CODEC_NONE) {
                return VDEC_PRIO_PROBE_HI;
        }
        return ugc == I420 || ugc == RGB || ugc == RGBA || ugc == UYVY ? VDEC_PRIO_PREFERRED : VDEC_PRIO_NA;
}
static const struct video_decompress_info gpujpeg_dec_info = { gpujpeg_decompress_init, gpujpeg_decompress_reconfigure, gpujpeg_decompress, no_corrupted_frames,
                                                               gpujpeg_decompress_done, gpujpeg_decompress_get_priority };
REGISTER_MODULE(gpujpeg, &gpujpeg_dec_info, LIBRARY_CLASS_VIDEO_DECOMPRESS, VIDEO_DECOMPRESS_ABI_VERSION);

// ---- gpujpeg_to_dxt ------------------------------------------------------------------------------------------------------------------
namespace {
struct state_gpujpeg_to_dxt {
        ugb200_jpeg_decoder *decoder = nullptr;
        void *rgb = nullptr, *dxt = nullptr;  // device
        size_t rgb_cap = 0, dxt_cap = 0;
        struct video_desc desc{};
        codec_t out_codec = VIDEO_CODEC_NONE;
};
}  // namespace
static void *gpujpeg_to_dxt_init(void)
{
        if (cuda_wrapper_set_device((int) cuda_devices[0]) != CUDA_WRAPPER_SUCCESS) {
                return nullptr;
        }
        auto *s = new state_gpujpeg_to_dxt();
        s->decoder = ugb200_jpeg_decoder_create(nullptr);
        if (!s->decoder) {
                delete s;
                return nullptr;
        }
        return s;
}
static int gpujpeg_to_dxt_reconfigure(void *state, struct video_desc desc, int, int, int, int pitch, codec_t out_codec)
{
        auto *s = (state_gpujpeg_to_dxt *) state;
        if ((out_codec != DXT1 && out_codec != DXT5) || desc.width % 4 || desc.height % 4 || pitch != (int) vc_get_linesize(desc.width, out_codec)) {
                return 0;  // gpujpeg_to_dxt.cpp:230-236
        }
        const size_t rgb = (size_t) desc.width * desc.height * 3, dxt = (size_t) desc.width * desc.height / (out_codec == DXT1 ? 2 : 1);
        if (rgb > s->rgb_cap) {
                cuda_wrapper_free(s->rgb);
                if (cuda_wrapper_malloc(&s->rgb, rgb) != CUDA_WRAPPER_SUCCESS) {
                        return 0;
                }
                s->rgb_cap = rgb;
        }
        if (dxt > s->dxt_cap) {
                cuda_wrapper_free(s->dxt);
                if (cuda_wrapper_malloc(&s->dxt, dxt) != CUDA_WRAPPER_SUCCESS) {
                        return 0;
                }
                s->dxt_cap = dxt;
        }
        s->desc = desc, s->out_codec = out_codec;
        return ugb200_jpeg_decoder_expect(s->decoder, (int) desc.width, (int) desc.height) == 0;  // s->rgb holds width * height * 3 bytes
}
/// worker_thread, gpujpeg_to_dxt.cpp:134-166: decode to RGB on the device, encode with mirrored height, copy the blocks out
static decompress_status gpujpeg_to_dxt_decompress(void *state, unsigned char *dst, unsigned char *buffer, unsigned int src_len, int, struct video_frame_callbacks *,
                                                   struct pixfmt_desc *)
{
        auto *s = (state_gpujpeg_to_dxt *) state;
        cuda_wrapper_set_device((int) cuda_devices[0]);
        if (ugb200_jpeg_decode(s->decoder, buffer, src_len, s->rgb, 1, 0, RGB, 0, 8, 16) != 0) {
                return DECODER_NO_FRAME;
        }
        const int w = (int) s->desc.width, h = (int) s->desc.height;
        const int rc = s->out_codec == DXT1 ? cuda_rgb_to_dxt1(s->rgb, s->dxt, w, -h, nullptr) : cuda_rgb_to_dxt6(s->rgb, s->dxt, w, -h, nullptr);
        if (rc != 0 || cuda_wrapper_memcpy(dst, s->dxt, (size_t) w * h / (s->out_codec == DXT1 ? 2 : 1), CUDA_WRAPPER_MEMCPY_DEVICE_TO_HOST) != CUDA_WRAPPER_SUCCESS) {
                return DECODER_NO_FRAME;
        }
        return DECODER_GOT_FRAME;
}
static void gpujpeg_to_dxt_done(void *state)
{
        auto *s = (state_gpujpeg_to_dxt *) state;
        ugb200_jpeg_decoder_destroy(s->decoder);
        cuda_wrapper_free(s->rgb), cuda_wrapper_free(s->dxt);
        delete s;
}
static int gpujpeg_to_dxt_get_priority(codec_t compression, struct pixfmt_desc, codec_t ugc)  // gpujpeg_to_dxt.cpp:364-369
{
        return compression == JPEG && (ugc == DXT1 || ugc == DXT5) ? 900 : -1;
}
static const struct video_decompress_info gpujpeg_to_dxt_info = { gpujpeg_to_dxt_init, gpujpeg_to_dxt_reconfigure, gpujpeg_to_dxt_decompress, no_corrupted_frames,
                                                                  gpujpeg_to_dxt_done, gpujpeg_to_dxt_get_priority };
REGISTER_MODULE(gpujpeg_to_dxt, &gpujpeg_to_dxt_info, LIBRARY_CLASS_VIDEO_DECOMPRESS, VIDEO_DECOMPRESS_ABI_VERSION);

// ---- dxt_cuda: DXT1 / DXT5-YCoCg -> RGB / RGBA / UYVY ----------------------------------------------------------------------------------
namespace {
struct state_dxt_cuda {
        void *blocks = nullptr, *rgb = nullptr, *conv = nullptr;  // device
        size_t blocks_cap = 0, rgb_cap = 0, conv_cap = 0;
        struct video_desc desc{};
        int rshift = 0, gshift = 8, bshift = 16, pitch = 0;
        codec_t out_codec = VIDEO_CODEC_NONE;
};
bool dev_grow(void *&p, size_t &cap, size_t need)
{
        if (need <= cap) {
                return true;
        }
        cuda_wrapper_free(p);
        p = nullptr, cap = 0;
        if (cuda_wrapper_malloc(&p, need) != CUDA_WRAPPER_SUCCESS) {
                return false;
        }
        cap = need;
        return true;
}
}  // namespace
static void *dxt_cuda_init(void) { return cuda_wrapper_set_device((int) cuda_devices[0]) == CUDA_WRAPPER_SUCCESS ? new state_dxt_cuda() : nullptr; }
static int dxt_cuda_reconfigure(void *state, struct video_desc desc, int rshift, int gshift, int bshift, int pitch, codec_t out_codec)
{
        auto *s = (state_dxt_cuda *) state;
        if ((desc.color_spec != DXT1 && desc.color_spec != DXT5) || (out_codec != RGB && out_codec != RGBA && out_codec != UYVY) || desc.width % 4 || desc.height % 4) {
                return 0;
        }
        s->desc = desc, s->rshift = rshift, s->gshift = gshift, s->bshift = bshift, s->pitch = pitch, s->out_codec = out_codec;
        return dev_grow(s->blocks, s->blocks_cap, (size_t) desc.width * desc.height) && dev_grow(s->rgb, s->rgb_cap, (size_t) desc.width * desc.height * 3 + 64) &&
               dev_grow(s->conv, s->conv_cap, (size_t) desc.width * desc.height * 4 + 64);
}
static decompress_status dxt_cuda_decompress(void *state, unsigned char *dst, unsigned char *buffer, unsigned int src_len, int, struct video_frame_callbacks *,
                                             struct pixfmt_desc *)
{
        auto *s = (state_dxt_cuda *) state;
        const int w = (int) s->desc.width, h = (int) s->desc.height;
        const size_t need = (size_t) w * h / (s->desc.color_spec == DXT1 ? 2 : 1);
        if (src_len < need) {
                return DECODER_NO_FRAME;
        }
        cuda_wrapper_set_device((int) cuda_devices[0]);
        if (cuda_wrapper_memcpy(s->blocks, buffer, need, CUDA_WRAPPER_MEMCPY_HOST_TO_DEVICE) != CUDA_WRAPPER_SUCCESS) {
                return DECODER_NO_FRAME;
        }
        int rc = s->desc.color_spec == DXT1 ? ugb200_dxt1_to_rgb(s->blocks, s->rgb, w, h, 0, 0, nullptr) : ugb200_dxt5ycocg_to_rgb(s->blocks, s->rgb, w, h, 0, 0, nullptr);
        const void *res = s->rgb;
        const long ls = (long) vc_get_linesize((unsigned) w, s->out_codec);
        if (rc == 0 && s->out_codec != RGB) {
                rc = ugb200_pixfmt_convert(RGB, s->out_codec, s->conv, ls, s->rgb, (long) w * 3, (int) ls, h, 0, s->rshift, s->gshift, s->bshift, nullptr);
                res = s->conv;
        }
        if (rc != 0) {
                return DECODER_NO_FRAME;
        }
        const long pitch = s->pitch ? s->pitch : ls;
        if (pitch == ls) {
                rc = cuda_wrapper_memcpy(dst, res, (size_t) ls * h, CUDA_WRAPPER_MEMCPY_DEVICE_TO_HOST);
        } else {
                rc = cuda_wrapper_memcpy2d(dst, (size_t) pitch, res, (size_t) ls, (size_t) ls, (size_t) h, CUDA_WRAPPER_MEMCPY_DEVICE_TO_HOST);
        }
        return rc == CUDA_WRAPPER_SUCCESS ? DECODER_GOT_FRAME : DECODER_NO_FRAME;
}
static void dxt_cuda_done(void *state)
{
        auto *s = (state_dxt_cuda *) state;
        cuda_wrapper_free(s->blocks), cuda_wrapper_free(s->rgb), cuda_wrapper_free(s->conv);
        delete s;
}
static int dxt_cuda_get_priority(codec_t compression, struct pixfmt_desc, codec_t ugc)  // same contract as dxt_glsl.c:228-237 (+ RGB)
{
        if (compression != DXT1 && compression != DXT5) {
                return -1;
        }
        return ugc == RGBA || ugc == UYVY || ugc == RGB ? 500 : -1;
}
static const struct video_decompress_info dxt_cuda_info = { dxt_cuda_init, dxt_cuda_reconfigure, dxt_cuda_decompress, no_corrupted_frames, dxt_cuda_done,
                                                            dxt_cuda_get_priority };
REGISTER_MODULE(dxt_cuda, &dxt_cuda_info, LIBRARY_CLASS_VIDEO_DECOMPRESS, VIDEO_DECOMPRESS_ABI_VERSION);

// ---- plain-C driver (include/ugb200_vcompress.h) ---------------------------------------------------------------------------------------
struct ugb200_decompress {
        state_decompress *s;
};
extern "C" {
UGB_API ugb200_decompress *ugb200_decompress_init(int compression, int out_codec)
{
        state_decompress *st = nullptr;
        struct pixfmt_desc internal = { 0, SUBS_UNKNOWN, false };
        if (!decompress_init_multi((codec_t) compression, internal, (codec_t) out_codec, &st, 1)) {
                return nullptr;
        }
        return new ugb200_decompress{ st };
}
UGB_API const char *ugb200_decompress_module(ugb200_decompress *d) { return d ? decompress_module_name(d->s) : ""; }
UGB_API int ugb200_decompress_reconfigure(ugb200_decompress *d, int width, int height, int compression, int rshift, int gshift, int bshift, int pitch, int out_codec)
{
        if (!d) {
                return 0;
        }
        const struct video_desc desc = { (unsigned) width, (unsigned) height, (codec_t) compression, 30.0, 0, 1 };
        return decompress_reconfigure(d->s, desc, rshift, gshift, bshift, pitch, (codec_t) out_codec);
}
UGB_API int ugb200_decompress_frame(ugb200_decompress *d, void *dst, const void *src, unsigned src_len, int frame_seq, int *internal_props)
{
        if (!d) {
                return DECODER_NO_FRAME;
        }
        struct pixfmt_desc prop = { 0, SUBS_UNKNOWN, false };
        const decompress_status st = decompress_frame(d->s, (unsigned char *) dst, (unsigned char *) src, src_len, frame_seq, nullptr, &prop);
        if (internal_props) {
                internal_props[0] = prop.depth, internal_props[1] = (int) prop.subsampling, internal_props[2] = prop.rgb;
        }
        return (int) st;
}
UGB_API void ugb200_decompress_done(ugb200_decompress *d)
{
        if (d) {
                decompress_done(d->s);
                delete d;
        }
}
}
