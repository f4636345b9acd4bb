/* TEST INFRASTRUCTURE - drives the UNMODIFIED UltraGrid compress framework (src/video_compress.cpp + src/lib_common.cpp, compiled from the
 * reference tree into oracle/_ref/libugframework.so by oracle/Makefile) the way UltraGrid's sender does: modules are dlopen()ed like
 * open_all() does (lib_common.cpp:186-204), looked up by name through load_library(), and fed real struct video_frame's through
 * compress_init / compress_frame / compress_pop (src/video_compress.h:84-96).  Compiled against the reference's own headers; supplies the few
 * globals that live in src/host.cpp.  Used by tests/test_real_module.py to show that ultragrid_b200/modules/ultragrid_vcompress_*.so load into an
 * unmodified UltraGrid.  The decompress side (fwd_dec_*) does the same with src/video_decompress.c: decompress_init_multi picks the module by the
 * priorities the modules report, decompress_reconfigure / decompress_frame / decompress_done drive it.  Never part of the product. */
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <memory>

#include "host.h"
#include "lib_common.h"
#include "module.h"
#include "video_codec.h"
#include "video_compress.h"
#include "video_decompress.h"
#include "video_frame.h"

/* src/host.cpp:177-179 */
unsigned int cuda_devices[MAX_CUDA_DEVICES] = { 0 };
unsigned int cuda_devices_count = 1;
bool cuda_devices_explicit = false;

struct fwd_state {
        struct module root;
        struct compress_state *cs = nullptr;
};

extern "C" {
#define API __attribute__((visibility("default")))

API int fwd_load_module(const char *path)
{
        void *h = dlopen(path, RTLD_NOW | RTLD_GLOBAL); /* as open_all(), lib_common.cpp:197 */
        if (!h) {
                fprintf(stderr, "fwd_load_module: %s\n", dlerror());
                return -1;
        }
        return 0;
}

API int fwd_has_module(const char *name)
{
        return load_library(name, LIBRARY_CLASS_VIDEO_COMPRESS, VIDEO_COMPRESS_ABI_VERSION) != nullptr;
}

API int fwd_has_decompress_module(const char *name)
{
        return load_library(name, LIBRARY_CLASS_VIDEO_DECOMPRESS, VIDEO_DECOMPRESS_ABI_VERSION) != nullptr;
}

API void fwd_set_device(int dev) { cuda_devices[0] = (unsigned) dev, cuda_devices_count = 1; }

API void *fwd_init(const char *cfg)
{
        auto *s = new fwd_state();
        module_init_default(&s->root); /* as init_root_module(), src/host.cpp:707-715 */
        s->root.cls = MODULE_CLASS_ROOT;
        module_register(&s->root, nullptr);
        if (compress_init(&s->root, cfg, &s->cs) != 0) {
                module_done(&s->root);
                delete s;
                return nullptr;
        }
        return s;
}

/* data == NULL: poison pill (src/video_compress.h:143-147) */
API void fwd_frame(void *st, void *data, int is_cuda, int width, int height, int codec, double fps)
{
        auto *s = (fwd_state *) st;
        if (!data) {
                compress_frame(s->cs, {});
                return;
        }
        struct video_desc d {};
        d.width = (unsigned) width, d.height = (unsigned) height, d.color_spec = (codec_t) codec, d.fps = fps, d.interlacing = PROGRESSIVE, d.tile_count = 1;
        std::shared_ptr<video_frame> f(vf_alloc_desc(d), vf_free); /* data stays the caller's: no data_deleter */
        f->tiles[0].data = (char *) data;
        f->tiles[0].data_len = (unsigned) vc_get_datalen((unsigned) width, (unsigned) height, (codec_t) codec);
        f->mem_location = is_cuda ? CUDA_MEM : CPU_MEM;
        compress_frame(s->cs, std::move(f));
}

/* 0 ok, 1 end of stream, -1 frame larger than cap */
API int fwd_pop(void *st, void *out, size_t cap, size_t *len, int *codec, unsigned *seq, unsigned *width, unsigned *height)
{
        auto *s = (fwd_state *) st;
        std::shared_ptr<video_frame> f = compress_pop(s->cs);
        if (!f) {
                return 1;
        }
        *len = f->tiles[0].data_len, *codec = (int) f->color_spec, *seq = f->seq, *width = f->tiles[0].width, *height = f->tiles[0].height;
        if (f->tiles[0].data_len > cap) {
                return -1;
        }
        memcpy(out, f->tiles[0].data, f->tiles[0].data_len);
        return 0;
}

API void fwd_done(void *st)
{
        auto *s = (fwd_state *) st;
        compress_done(s->cs);
        module_done(&s->root);
        delete s;
}

/* ---- decompress side: the receiver's calls (src/rtp/video_decoders.cpp: decompress_init_multi -> reconfigure -> frame) ---- */
API void *fwd_dec_init(int compression, int out_codec)
{
        struct state_decompress *st = nullptr;
        struct pixfmt_desc internal {};
        if (!decompress_init_multi((codec_t) compression, internal, (codec_t) out_codec, &st, 1)) {
                return nullptr;
        }
        return st;
}

API int fwd_dec_reconfigure(void *st, int width, int height, int compression, int rshift, int gshift, int bshift, int pitch, int out_codec)
{
        struct video_desc d {};
        d.width = (unsigned) width, d.height = (unsigned) height, d.color_spec = (codec_t) compression, d.fps = 30.0, d.interlacing = PROGRESSIVE, d.tile_count = 1;
        return decompress_reconfigure((struct state_decompress *) st, d, rshift, gshift, bshift, pitch, (codec_t) out_codec);
}

/* returns the decompress_status; props = { depth, subsampling, rgb } as filled for DECODER_GOT_CODEC */
API int fwd_dec_frame(void *st, void *dst, void *src, unsigned src_len, int frame_seq, int *props)
{
        struct pixfmt_desc p {};
        const decompress_status rc = decompress_frame((struct state_decompress *) st, (unsigned char *) dst, (unsigned char *) src, src_len, frame_seq, nullptr, &p);
        props[0] = p.depth, props[1] = (int) p.subsampling, props[2] = p.rgb;
        return (int) rc;
}

API int fwd_dec_accepts_corrupted(void *st)
{
        int v = -1;
        size_t len = sizeof v;
        return decompress_get_property((struct state_decompress *) st, DECOMPRESS_PROPERTY_ACCEPTS_CORRUPTED_FRAME, &v, &len) ? v : -1;
}

API void fwd_dec_done(void *st) { decompress_done((struct state_decompress *) st); }
}
