// cuda_wrapper C ABI (include/cuda_wrapper.h) — behaviour of UltraGrid's src/cuda_wrapper.cu:82-181.
#include <cstdio>
#include <cstdlib>

#include <cuda_runtime.h>

#include "../../include/cuda_wrapper.h"

extern "C" {

int cuda_wrapper_free(void *buffer) { return (int) cudaFree(buffer); }
int cuda_wrapper_free_host(void *buffer) { return (int) cudaFreeHost(buffer); }
int cuda_wrapper_host_alloc(void **pHost, size_t size, unsigned int flags) { return (int) cudaHostAlloc(pHost, size, flags); }
int cuda_wrapper_malloc(void **buffer, size_t data_len) { return (int) cudaMalloc(buffer, data_len); }
int cuda_wrapper_malloc_host(void **buffer, size_t data_len) { return (int) cudaMallocHost(buffer, data_len); }

static cudaMemcpyKind kind_of(int kind)
{
        switch (kind) {
        case CUDA_WRAPPER_MEMCPY_HOST_TO_DEVICE:
                return cudaMemcpyHostToDevice;
        case CUDA_WRAPPER_MEMCPY_DEVICE_TO_HOST:
                return cudaMemcpyDeviceToHost;
        }
        abort();  // src/cuda_wrapper.cu:79
}

int cuda_wrapper_memcpy(void *dst, const void *src, size_t count, int kind)
{
        return (int) cudaMemcpy(dst, src, count, kind_of(kind));
}
const char *cuda_wrapper_last_error_string(void) { return cudaGetErrorString(cudaGetLastError()); }
int cuda_wrapper_get_last_error(void) { return (int) cudaGetLastError(); }
const char *cuda_wrapper_get_error_string(int error) { return cudaGetErrorString((cudaError_t) error); }
int cuda_wrapper_set_device(int index) { return (int) cudaSetDevice(index); }

void cuda_wrapper_print_devices_info(bool full)
{
        int n = 0;
        if (cudaGetDeviceCount(&n) != cudaSuccess) {
                fprintf(stderr, "Cannot get number of CUDA devices: %s\n", cudaGetErrorString(cudaGetLastError()));
                return;
        }
        if (n == 0) {
                fprintf(stderr, "There is no device supporting CUDA.\n");
                return;
        }
        printf("There %s %d devices supporting CUDA:\n", n == 1 ? "is" : "are", n);
        for (int i = 0; i < n; ++i) {
                cudaDeviceProp p;
                if (cudaGetDeviceProperties(&p, i) != cudaSuccess) {
                        fprintf(stderr, "Cannot get CUDA device #%d properties: %s\n", i, cudaGetErrorString(cudaGetLastError()));
                        continue;
                }
                printf("%sDevice #%d: %s\n", full ? "\n" : "", i, p.name);
                if (!full) {
                        continue;
                }
                printf("  Compute capability: %d.%d\n", p.major, p.minor);
                printf("  Total amount of global memory: %zu KiB\n", p.totalGlobalMem / 1024);
                printf("  Total amount of shared memory per block: %zu KiB\n", p.sharedMemPerBlock / 1024);
                printf("  Total number of registers available per block: %d\n", p.regsPerBlock);
                printf("  Multiprocessors: %d\n", p.multiProcessorCount);
                printf("  L2 cache: %d KiB\n", p.l2CacheSize / 1024);
        }
}

void cuda_wrapper_device_reset(void)
{
        if (cudaDeviceReset() != cudaSuccess) {
                fprintf(stderr, "cudaDeviceReset failed!\n");
        }
}

int cuda_wrapper_get_device_count(int *count) { return (int) cudaGetDeviceCount(count); }
int cuda_wrapper_stream_create(cuda_wrapper_stream_t *stream)
{
        return (int) cudaStreamCreateWithFlags((cudaStream_t *) stream, cudaStreamNonBlocking);
}
int cuda_wrapper_stream_destroy(cuda_wrapper_stream_t stream) { return (int) cudaStreamDestroy((cudaStream_t) stream); }
int cuda_wrapper_stream_synchronize(cuda_wrapper_stream_t stream) { return (int) cudaStreamSynchronize((cudaStream_t) stream); }
int cuda_wrapper_memcpy_async(void *dst, const void *src, size_t count, int kind, cuda_wrapper_stream_t stream)
{
        return (int) cudaMemcpyAsync(dst, src, count, kind_of(kind), (cudaStream_t) stream);
}

int cuda_wrapper_memcpy2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, int kind)
{
        return (int) cudaMemcpy2D(dst, dpitch, src, spitch, width, height, kind_of(kind));
}

}  // extern "C"
