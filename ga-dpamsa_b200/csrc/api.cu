// libgadpamsa: error channel + device plumbing for hosts without torch.
#include "gad_common.cuh"

#include <string.h>

namespace {
thread_local char g_err[512] = "";
}

void gad_set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int gad_version(void) { return 100; }   // 0.1.0

extern "C" const char *gad_last_error(void) { return g_err; }

extern "C" int gad_device_count(int *count)
{
    GAD_REQUIRE(count, "gad_device_count: null pointer");
    *count = 0;
    GAD_CUDA(cudaGetDeviceCount(count));
    return GAD_OK;
}

extern "C" int gad_set_device(int device)
{
    GAD_CUDA(cudaSetDevice(device));
    return GAD_OK;
}

extern "C" int gad_malloc(void **d_ptr, size_t bytes)
{
    GAD_REQUIRE(d_ptr, "gad_malloc: null pointer");
    GAD_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return GAD_OK;
}

extern "C" int gad_free(void *d_ptr)
{
    GAD_CUDA(cudaFree(d_ptr));
    return GAD_OK;
}

extern "C" int gad_malloc_host(void **h_ptr, size_t bytes)
{
    GAD_REQUIRE(h_ptr, "gad_malloc_host: null pointer");
    GAD_CUDA(cudaMallocHost(h_ptr, bytes ? bytes : 1));
    return GAD_OK;
}

extern "C" int gad_free_host(void *h_ptr)
{
    GAD_CUDA(cudaFreeHost(h_ptr));
    return GAD_OK;
}

extern "C" int gad_memcpy_h2d(void *d_dst, const void *h_src, size_t bytes, void *stream)
{
    GAD_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, gad_stream(stream)));
    return GAD_OK;
}

extern "C" int gad_memcpy_d2h(void *h_dst, const void *d_src, size_t bytes, void *stream)
{
    GAD_CUDA(cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, gad_stream(stream)));
    return GAD_OK;
}

extern "C" int gad_memset(void *d_dst, int value, size_t bytes, void *stream)
{
    GAD_CUDA(cudaMemsetAsync(d_dst, value, bytes, gad_stream(stream)));
    return GAD_OK;
}

extern "C" int gad_stream_sync(void *stream)
{
    GAD_CUDA(cudaStreamSynchronize(gad_stream(stream)));
    return GAD_OK;
}

// Lets kernels launched on the current device dereference pointers into `peer_device`'s memory (NVLink / PCIe P2P).
// Used by scripts/peer_crossover_probe.py to run gad_crossover_peer in ONE process on two GPUs (so that it can be
// profiled); the sharded loop gets its peer pointers from symmetric-memory allocations instead.
extern "C" int gad_enable_peer_access(int peer_device)
{
    int can = 0, cur = 0;
    GAD_CUDA(cudaGetDevice(&cur));
    GAD_CUDA(cudaDeviceCanAccessPeer(&can, cur, peer_device));
    GAD_REQUIRE(can, "device %d cannot access device %d", cur, peer_device);
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return GAD_OK;
    }
    GAD_CUDA(e);
    return GAD_OK;
}
