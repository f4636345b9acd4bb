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


#include "decompress_modules.h"  // gpujpeg, gpujpeg_to_dxt, dxt_cuda: shared with the real-ABI build (module/ug_decompress_module.cpp)

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
