// video_decompress module API — host mirror of src/video_decompress.h:42-213 (pure C in the reference) and of the selection logic of
// src/video_decompress.c:100-230 (best module by priority for a compression / internal format / output codec triple).
#pragma once
#include "ug_types.h"
#include "video_compress.h"  // registry

#define VIDEO_DECOMPRESS_ABI_VERSION 6                  // src/video_decompress.h:42
#define DECOMPRESS_PROPERTY_ACCEPTS_CORRUPTED_FRAME 1   // :69, int

struct video_frame_callbacks;  // used only by libavcodec in the reference

typedef void *(*decompress_init_t)(void);
typedef int (*decompress_reconfigure_t)(void *state, struct video_desc desc, int rshift, int gshift, int bshift, int pitch, codec_t out_codec);
typedef enum {  // :91-96
        DECODER_NO_FRAME = 0,
        DECODER_GOT_FRAME,
        DECODER_GOT_CODEC,
        DECODER_UNSUPP_PIXFMT,
} decompress_status;
typedef decompress_status (*decompress_decompress_t)(void *state, unsigned char *dst, unsigned char *buffer, unsigned int src_len, int frame_seq,
                                                     struct video_frame_callbacks *callbacks, struct pixfmt_desc *internal_prop);
typedef int (*decompress_get_property_t)(void *state, int property, void *val, size_t *len);
typedef void (*decompress_done_t)(void *);
enum vdec_priority {  // :139-155
        VDEC_PRIO_NA = -1,
        VDEC_PRIO_PROBE_HI = 50,
        VDEC_PRIO_PROBE_LO = 80,
        VDEC_PRIO_PREFERRED = 200,
        VDEC_PRIO_NORMAL = 500,
        VDEC_PRIO_NOT_PREFERRED = 800,
        VDEC_PRIO_LOW = 900,
};
typedef int (*decompress_get_priority_t)(codec_t compression, struct pixfmt_desc internal, codec_t ugc);

struct video_decompress_info {  // :164-171
        decompress_init_t init;
        decompress_reconfigure_t reconfigure;
        decompress_decompress_t decompress;
        decompress_get_property_t get_property;
        decompress_done_t done;
        decompress_get_priority_t get_decompress_priority;
};

// framework entry points, :173-213
struct state_decompress;
bool decompress_init_multi(codec_t compression, struct pixfmt_desc internal, codec_t to, struct state_decompress **out, int count);
int decompress_reconfigure(struct state_decompress *, struct video_desc, int rshift, int gshift, int bshift, int pitch, codec_t out_codec);
decompress_status decompress_frame(struct state_decompress *, unsigned char *dst, unsigned char *src, unsigned int src_len, int frame_seq,
                                   struct video_frame_callbacks *callbacks, struct pixfmt_desc *internal_prop);
int decompress_get_property(struct state_decompress *state, int property, void *val, size_t *len);
void decompress_done(struct state_decompress *);
const char *decompress_module_name(struct state_decompress *);  // which module was selected (the reference logs it)
