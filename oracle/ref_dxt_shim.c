/* TEST INFRASTRUCTURE.  Exposes the static decoder of the reference's tool cuda_dxt/dxt62tga.c (DXT5-YCoCg -> BGR, :24-108) by
 * compiling that file in place (no copy) with its main() renamed. */
#define main dxt62tga_tool_main
#include "cuda_dxt/dxt62tga.c"
#undef main

__attribute__((visibility("default"))) void ref_dxt5ycocg_to_bgr(const void *in, unsigned char *out, int size_x, int size_y)
{
        dxt5ycocg_to_bgr((const u64 *) in, out, size_x, size_y);
}
