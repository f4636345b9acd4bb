/* The reference's CUDA pre/post-processing callbacks for the Comprimato J2K codec (src/cuda_wrapper/kernels.hpp:46-75, kernels.cu:261-306,452-489;
 * bound at src/video_decompress/cmpto_j2k.cpp:433 and src/video_compress/cmpto_j2k.cpp:349), same C++ names and signatures, so that a translation
 * unit including this header links against libugb200 instead of src/cuda_wrapper/kernels.cu.  All pointers are DEVICE pointers; rows are tightly
 * packed as in the reference: RG48 size_x * 6 bytes, R12L ceil(size_x / 8) * 36 bytes.  Asynchronous on `stream`; returns cudaGetLastError() (0 = ok).
 * The reference times every call with two events and an event synchronize (MEASURE_KERNEL_DURATION_*); that host stall is not reproduced.
 *   postprocess_rg48_to_r12l   16-bit RGB -> R12L (sample >> 4), the last group of a row whose width is not a multiple of 8 is written whole
 *                              (the reference fills it from an uninitialised temporary: bytes that depend on samples beyond size_x are indeterminate
 *                              there; here they come from whatever follows the row in the input, zero behind the last row)
 *   preprocess_r12l_to_rg48    R12L -> 16-bit RGB (sample << 4), exactly size_x * 6 bytes per row
 * The unused parameters (codec handles, component formats, buffer sizes, temp buffer) are accepted and ignored, as in the reference. */
#ifndef UGB200_CUDA_WRAPPER_KERNELS_HPP
#define UGB200_CUDA_WRAPPER_KERNELS_HPP
#include <cstddef>

#include "cuda_wrapper.h"

struct cmpto_j2k_dec_comp_format;
struct cmpto_j2k_enc_comp_format;
UGB_API int postprocess_rg48_to_r12l(void *postprocessor, void *img_custom_data, size_t img_custom_data_size, int size_x, int size_y,
                                     struct cmpto_j2k_dec_comp_format *comp_formats, int comp_count, void *input_samples, size_t input_samples_size,
                                     void *temp_buffer, size_t temp_buffer_size, void *output_buffer, size_t output_buffer_size, void *stream);
UGB_API int preprocess_r12l_to_rg48(void *preprocessor, void *img_custom_data, size_t img_custom_data_size, int size_x, int size_y,
                                    struct cmpto_j2k_enc_comp_format *comp_formats, int comp_count, void *input_samples, size_t input_samples_size,
                                    void *output_samples, size_t output_samples_size, void *stream);
#endif
