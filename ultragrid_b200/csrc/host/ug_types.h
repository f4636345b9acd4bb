// Host-side mirror of the UltraGrid types the compress path touches: codec_t (src/types.h:62-112), video_desc
// (:240-250), tile / video_frame (:281-343), mem_location_t (:295-298).  Field names are kept so that module code
// reads like the reference's; only what the hot path needs is present.  g++ only — no CUDA headers here.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <memory>

#include "../../../include/ugb200.h"

typedef enum ugb200_codec codec_t;
static const codec_t VIDEO_CODEC_NONE = UGB_VIDEO_CODEC_NONE, RGBA = UGB_RGBA, UYVY = UGB_UYVY, YUYV = UGB_YUYV, VUYA = UGB_VUYA,
                     R10k = UGB_R10k, R12L = UGB_R12L, v210 = UGB_v210, DXT1 = UGB_DXT1, DXT5 = UGB_DXT5, RGB = UGB_RGB, JPEG = UGB_JPEG,
                     BGR = UGB_BGR, RG48 = UGB_RG48, I420 = UGB_I420, Y216 = UGB_Y216, Y416 = UGB_Y416,
                     VIDEO_CODEC_END = UGB_VIDEO_CODEC_COUNT;

enum mem_location_t { CPU_MEM = 0, CUDA_MEM };  // types.h:295-298
enum subsampling { SUBS_UNKNOWN = 0, SUBS_420 = 4200, SUBS_422 = 4220, SUBS_444 = 4440, SUBS_4444 = 4444 };  // types.h:136-142
struct pixfmt_desc {  // types.h:144-149
        int depth;
        enum subsampling subsampling;
        bool rgb;
};

struct video_desc {  // types.h:240-250
        unsigned int width, height;
        codec_t color_spec;
        double fps;
        int interlacing;
        unsigned int tile_count;
};

struct tile {  // types.h:281-293
        unsigned int width, height;
        char *data;
        unsigned int data_len;
};

struct video_frame {  // types.h:303-343 (single tile; tiled frames are split by the framework, video_compress.cpp:441-490)
        codec_t color_spec;
        int interlacing;
        double fps;
        enum mem_location_t mem_location;
        uint32_t seq;
        uint64_t compress_start, compress_end;  // ns stamps, types.h:336-337
        unsigned int tile_count;
        struct tile tiles[1];
        void (*data_deleter)(void *);  // how tiles[0].data is released (pool frames: pinned memory)
};

enum { MAX_CUDA_DEVICES = 8 };  // src/host.h:97 has 4; raised for the 8 x B200 box
extern unsigned int cuda_devices[MAX_CUDA_DEVICES];  // src/host.cpp:177-179
extern unsigned int cuda_devices_count;

static inline struct video_desc video_desc_from_frame(const struct video_frame *f)
{
        return video_desc{ f->tiles[0].width, f->tiles[0].height, f->color_spec, f->fps, f->interlacing, f->tile_count };
}
static inline bool video_desc_eq(const video_desc &a, const video_desc &b)
{
        return a.width == b.width && a.height == b.height && a.color_spec == b.color_spec && a.fps == b.fps && a.interlacing == b.interlacing;
}
