/* Device-management C ABI — same symbols and semantics as UltraGrid's src/cuda_wrapper.h:61-73
 * (implementation: src/cuda_wrapper.cu:82-181).  Host modules compiled with gcc/g++ include only
 * this header and never see CUDA headers.  All functions return 0 (CUDA_WRAPPER_SUCCESS) or the raw
 * cudaError_t value as int; cuda_wrapper_memcpy aborts on an unknown `kind` like the reference does
 * (src/cuda_wrapper.cu:62-80).
 */
#ifndef UGB200_CUDA_WRAPPER_H
#define UGB200_CUDA_WRAPPER_H

#ifndef UGB_API
#define UGB_API __attribute__((visibility("default")))
#endif

#ifndef __cplusplus
#include <stdbool.h>
#endif
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CUDA_WRAPPER_SUCCESS 0

#define CUDA_WRAPPER_MEMCPY_HOST_TO_DEVICE 0
#define CUDA_WRAPPER_MEMCPY_DEVICE_TO_HOST 1

typedef void *cuda_wrapper_stream_t; /* a cudaStream_t; NULL = default stream */

UGB_API int cuda_wrapper_free(void *buffer);
UGB_API int cuda_wrapper_free_host(void *buffer);
UGB_API int cuda_wrapper_host_alloc(void **pHost, size_t size, unsigned int flags);
UGB_API int cuda_wrapper_malloc(void **buffer, size_t data_len);
UGB_API int cuda_wrapper_malloc_host(void **buffer, size_t data_len);
UGB_API int cuda_wrapper_memcpy(void *dst, const void *src, size_t count, int kind);
UGB_API const char *cuda_wrapper_last_error_string(void);
UGB_API int cuda_wrapper_set_device(int index);
UGB_API int cuda_wrapper_get_last_error(void);
UGB_API const char *cuda_wrapper_get_error_string(int error);
UGB_API void cuda_wrapper_print_devices_info(bool full);
UGB_API void cuda_wrapper_device_reset(void);

/* ---- B200 additions (not in the reference): what an asynchronous, multi-GPU host module needs ---- */
UGB_API int cuda_wrapper_get_device_count(int *count);
UGB_API int cuda_wrapper_stream_create(cuda_wrapper_stream_t *stream);   /* non-blocking stream */
UGB_API int cuda_wrapper_stream_destroy(cuda_wrapper_stream_t stream);
UGB_API int cuda_wrapper_stream_synchronize(cuda_wrapper_stream_t stream);
UGB_API int cuda_wrapper_memcpy_async(void *dst, const void *src, size_t count, int kind, cuda_wrapper_stream_t stream);
/* NUMA placement (two-socket 8-GPU boxes): node of a device (-1 unknown); move the calling thread (affinity + preferred memory node) next to
 * a device, returns the node or -1; pinned allocation whose pages live on the device's node (falls back to cudaMallocHost; freed with
 * cuda_wrapper_free_host like any other) */
UGB_API int cuda_wrapper_device_numa_node(int device);
UGB_API int cuda_wrapper_bind_thread_to_device(int device);
UGB_API int cuda_wrapper_malloc_host_near(void **buffer, size_t data_len, int device);
/* pitched copy (cudaMemcpy2D), synchronous like cuda_wrapper_memcpy */
UGB_API int cuda_wrapper_memcpy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, int kind);

#ifdef __cplusplus
}
#endif
#endif
