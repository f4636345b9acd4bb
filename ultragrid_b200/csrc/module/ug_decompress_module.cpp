// The decompress modules in UltraGrid's REAL ABI.
//
// Compiled against the reference's own headers (-I$(REF)/src, `make -C ultragrid_b200/csrc module`; no header is copied): struct video_desc / pixfmt_desc /
// codec_t (src/types.h), video_decompress_info and VIDEO_DECOMPRESS_ABI_VERSION (src/video_decompress.h:42,164-171), REGISTER_MODULE
// (src/lib_common.h:124-160), cuda_devices (src/host.h), vc_get_linesize (src/video_codec.h).  The results
//     ultragrid_b200/modules/ultragrid_vdecompress_gpujpeg.so          replaces the module built from src/video_decompress/gpujpeg.c
//     ultragrid_b200/modules/ultragrid_vdecompress_gpujpeg_to_dxt.so   replaces the one built from src/video_decompress/gpujpeg_to_dxt.cpp
//     ultragrid_b200/modules/ultragrid_vdecompress_dxt_cuda.so         CUDA counterpart of the OpenGL module src/video_decompress/dxt_glsl.c
// are what an unmodified UltraGrid dlopen()s from lib/ultragrid/ (lib_common.cpp:186-204) and selects by priority (src/video_decompress.c:100-230).
// The module bodies are ../host/decompress_modules.h - the same text the mirror build (libugb200.so) compiles; -DUGB_DECOMPRESS_MODULES=<bit> picks one.
// tests/test_real_module.py loads all three through the reference's own lib_common.cpp + video_decompress.c (oracle/_ref/libugframework.so).
#include "host.h"              // cuda_devices
#include "lib_common.h"        // REGISTER_MODULE, LIBRARY_CLASS_VIDEO_DECOMPRESS
#include "types.h"             // codec_t, video_desc, pixfmt_desc
#include "video_codec.h"       // vc_get_linesize
#include "video_decompress.h"  // video_decompress_info, decompress_status, vdec_priority

#include "../host/decompress_modules.h"
