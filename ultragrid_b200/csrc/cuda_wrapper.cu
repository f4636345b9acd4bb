// cuda_wrapper C ABI (include/cuda_wrapper.h) — behaviour of UltraGrid's src/cuda_wrapper.cu:82-181.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>

#include <sched.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <cuda_runtime.h>

#include "../../include/cuda_wrapper.h"

// ---- NUMA placement of pinned host memory (B200 addition) ---------------------------------------------------------------------------
// An 8-GPU HGX box has two sockets with four GPUs each.  A frame that is pinned on the other socket crosses the inter-socket link on
// its way to the GPU; with eight streams of 53 GB/s each that link, not PCIe, bounds the end-to-end rate (round 1: 35 GB/s per GPU at
// N = 8 against 53.6 GB/s at N = 1).  cuda_wrapper_malloc_host_near() therefore places the pages on the GPU's own node (mmap + mbind +
// first touch) and pins them with cudaHostRegister; cuda_wrapper_bind_thread_to_device() moves a worker thread next to its GPU.
namespace {
std::mutex g_near_lock;
std::map<void *, size_t> g_near;  // buffers of cuda_wrapper_malloc_host_near: base -> mapped length

int read_int_file(const char *path)
{
        FILE *f = fopen(path, "r");
        if (!f) {
                return -1;
        }
        int v = -1;
        if (fscanf(f, "%d", &v) != 1) {
                v = -1;
        }
        fclose(f);
        return v;
}
long sys_mbind(void *addr, unsigned long len, int mode, const unsigned long *mask, unsigned long maxnode, unsigned flags)
{
#ifdef SYS_mbind
        return syscall(SYS_mbind, addr, len, mode, mask, maxnode, flags);
#else
        return -1;
#endif
}
long sys_set_mempolicy(int mode, const unsigned long *mask, unsigned long maxnode)
{
#ifdef SYS_set_mempolicy
        return syscall(SYS_set_mempolicy, mode, mask, maxnode);
#else
        return -1;
#endif
}
constexpr int kMpolPreferred = 1;
}  // namespace

extern "C" int cuda_wrapper_device_numa_node(int device)
{
        char bus[32] = { 0 }, path[128];
        if (cudaDeviceGetPCIBusId(bus, sizeof bus, device) != cudaSuccess) {
                cudaGetLastError();
                return -1;
        }
        for (char *c = bus; *c; ++c) {  // sysfs spells the address in lower case
                if (*c >= 'A' && *c <= 'F') {
                        *c = (char) (*c - 'A' + 'a');
                }
        }
        snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", bus);
        return read_int_file(path);  // -1: unknown or a single-node machine
}

extern "C" int cuda_wrapper_bind_thread_to_device(int device)
{
        if (device < 0) {  // undo the memory preference (the affinity mask is the caller's to restore)
                sys_set_mempolicy(0 /* MPOL_DEFAULT */, nullptr, 0);
                return -1;
        }
        const int node = cuda_wrapper_device_numa_node(device);
        if (node < 0 || node >= 64) {
                return -1;
        }
        char path[96];
        snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
        FILE *f = fopen(path, "r");
        if (f) {  // "0-31,64-95": the node's CPUs, intersected with what this thread may use today (cpuset / taskset of the launcher)
                cpu_set_t now, want;
                CPU_ZERO(&want);
                sched_getaffinity(0, sizeof now, &now);
                int a, b, any = 0;
                while (fscanf(f, "%d", &a) == 1) {
                        b = a;
                        int c = fgetc(f);
                        if (c == '-') {
                                if (fscanf(f, "%d", &b) != 1) {
                                        break;
                                }
                                c = fgetc(f);
                        }
                        for (int i = a; i <= b && i < CPU_SETSIZE; ++i) {
                                if (CPU_ISSET(i, &now)) {
                                        CPU_SET(i, &want);
                                        any = 1;
                                }
                        }
                        if (c != ',') {
                                break;
                        }
                }
                fclose(f);
                if (any) {
                        sched_setaffinity(0, sizeof want, &want);
                }
        }
        const unsigned long mask = 1ul << node;
        sys_set_mempolicy(kMpolPreferred, &mask, 65);  // later allocations of this thread (malloc arenas, cudaMallocHost) prefer the node
        return node;
}

extern "C" int cuda_wrapper_malloc_host_near(void **buffer, size_t data_len, int device)
{
        const int node = cuda_wrapper_device_numa_node(device);
        if (node < 0 || node >= 64 || data_len == 0) {
                return (int) cudaMallocHost(buffer, data_len);
        }
        const size_t len = (data_len + (2u << 20) - 1) & ~(size_t) ((2u << 20) - 1);
        void *p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p == MAP_FAILED) {
                return (int) cudaMallocHost(buffer, data_len);
        }
        const unsigned long mask = 1ul << node;
        sys_mbind(p, len, kMpolPreferred, &mask, 65, 0);  // best effort: without it the first touch below decides
        memset(p, 0, len);                                // the pages exist (on the node) before they are pinned
        const cudaError_t rc = cudaHostRegister(p, len, cudaHostRegisterPortable);
        if (rc != cudaSuccess) {
                munmap(p, len);
                cudaGetLastError();
                return (int) cudaMallocHost(buffer, data_len);
        }
        {
                std::lock_guard<std::mutex> lk(g_near_lock);
                g_near[p] = len;
        }
        *buffer = p;
        return 0;
}

extern "C" {

int cuda_wrapper_free(void *buffer) { return (int) cudaFree(buffer); }
int cuda_wrapper_free_host(void *buffer)
{
        size_t len = 0;
        {
                std::lock_guard<std::mutex> lk(g_near_lock);
                auto it = g_near.find(buffer);
                if (it != g_near.end()) {
                        len = it->second;
                        g_near.erase(it);
                }
        }
        if (len) {
                const cudaError_t rc = cudaHostUnregister(buffer);
                munmap(buffer, len);
                return (int) rc;
        }
        return (int) cudaFreeHost(buffer);
}
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
