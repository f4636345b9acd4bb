// video_compress module API — host mirror of src/video_compress.h:71-236 and of the registry of src/lib_common.h:124-160.
// A module exposes exactly one of the four API shapes; the framework (video_compress.cpp here, src/video_compress.cpp
// there) drives it and hands results to compress_pop().
#pragma once
#include <memory>

#include "ug_types.h"

#define VIDEO_COMPRESS_ABI_VERSION 14  // src/video_compress.h:71
enum library_class { LIBRARY_CLASS_VIDEO_DECOMPRESS = 7, LIBRARY_CLASS_VIDEO_COMPRESS = 8 };  // positions in src/lib_common.h:73-86

struct module;  // parent in the module tree — unused by the hot path

typedef void *(*compress_init_t)(struct module *parent, const char *cfg);  // NULL = error
typedef void (*compress_done_t)(void *state);
typedef std::shared_ptr<video_frame> (*compress_frame_t)(void *state, std::shared_ptr<video_frame> frame);
typedef std::shared_ptr<video_frame> (*compress_tile_t)(void *state, std::shared_ptr<video_frame> in_frame);
typedef void (*compress_frame_async_push_t)(void *state, std::shared_ptr<video_frame> in_frame);  // empty ptr = poison pill
typedef std::shared_ptr<video_frame> (*compress_frame_async_pop_t)(void *state);
typedef void (*compress_tile_async_push_t)(void *state, std::shared_ptr<video_frame> in_frame);
typedef std::shared_ptr<video_frame> (*compress_tile_async_pop_t)(void *state);

struct video_compress_info {  // src/video_compress.h:221-236
        compress_init_t init_func;
        compress_done_t done;
        compress_frame_t compress_frame_func;
        compress_tile_t compress_tile_func;
        compress_frame_async_push_t compress_frame_async_push_func;
        compress_frame_async_pop_t compress_frame_async_pop_func;
        compress_tile_async_push_t compress_tile_async_push_func;
        compress_tile_async_pop_t compress_tile_async_pop_func;
        void *get_module_info;
};

void register_library(const char *name, const void *info, enum library_class cls, int abi_version);
const void *load_library(const char *name, enum library_class cls, int abi_version);
/// every registered module of a class (what list_modules / get_libraries_for_class of src/lib_common.h:96-113 provide)
int get_libraries_for_class(enum library_class cls, int abi_version, const char **names, const void **infos, int max);
#define REGISTER_MODULE(name, info, lclass, abi)                                                                                           \
        static struct ugb_reg_##name {                                                                                                     \
                ugb_reg_##name() { register_library(#name, info, lclass, abi); }                                                           \
        } ugb_reg_instance_##name

// framework entry points, src/video_compress.h:95-107
struct compress_state;
int compress_init(struct module *parent, const char *config_string, struct compress_state **state);
void compress_frame(struct compress_state *, std::shared_ptr<video_frame>);
std::shared_ptr<video_frame> compress_pop(struct compress_state *);
void compress_done(struct compress_state *);

// pooled output frames in pinned host memory (the role of video_frame_pool + cuda_buffer_data_allocator,
// src/video_compress/cuda_dxt.cpp:68-83)
std::shared_ptr<video_frame> pinned_pool_get(size_t bytes, int device);
