// TEST INFRASTRUCTURE ONLY -- a host-memory stand-in for the CUDA runtime calls the library's host code makes, so that a translation unit of the
// product (its <<< >>> launches rewritten to emu::launch by tests/native_build.py) can run on the CPU execution model of ../cuda_emu.hpp.
// "Device" memory is host memory, streams are immediate, events carry no time.  Used by tests only; never on any product path.
#pragma once
#include <cstdlib>
#include <cstring>
#include <map>
#include <sys/mman.h>
#include <unistd.h>

// PLVS_EMU_GUARD=1: every "device" / pinned allocation ends flush against an inaccessible page and starts behind one, so a kernel or a copy that runs
// past the end of a buffer (or before its start, for 4 KiB-multiple sizes) faults at the offending access -- the CPU model's stand-in for
// compute-sanitizer's memcheck.  Sizes are rounded up to 16 bytes (vector loads), so overruns inside that slack go unseen.
namespace fake_cuda {
inline bool guard_mode() { static const bool on = std::getenv("PLVS_EMU_GUARD") != nullptr; return on; }
inline std::map<void*, std::pair<void*, size_t>>& guarded() { static std::map<void*, std::pair<void*, size_t>> m; return m; }
inline void* guarded_alloc(size_t bytes)
{
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t body = (bytes + 15) / 16 * 16;
    const size_t span = (body + page - 1) / page * page;
    char* base = (char*)mmap(nullptr, span + 2 * page, PROT_NONE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == (char*)MAP_FAILED) return nullptr;
    mprotect(base + page, span, PROT_READ | PROT_WRITE);
    char* p = base + page + (span - body);
    guarded()[p] = std::make_pair((void*)base, span + 2 * page);
    return p;
}
inline void guarded_free(void* p)
{
    auto it = guarded().find(p);
    if (it == guarded().end()) return;
    munmap(it->second.first, it->second.second);
    guarded().erase(it);
}
}  // namespace fake_cuda

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1, cudaErrorPeerAccessAlreadyEnabled = 704 };
typedef struct fake_stream* cudaStream_t;
typedef struct fake_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaHostAllocMapped = 2, cudaEventDisableTiming = 2 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };

inline const char* cudaGetErrorString(cudaError_t) { return "fake runtime error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaErrorInvalidValue; }
inline cudaError_t cudaMalloc(void** p, size_t bytes)
{
    *p = fake_cuda::guard_mode() ? fake_cuda::guarded_alloc(bytes) : std::aligned_alloc(256, (bytes + 255) / 256 * 256 + 256);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
inline cudaError_t cudaFree(void* p) { if (fake_cuda::guard_mode()) fake_cuda::guarded_free(p); else std::free(p); return cudaSuccess; }
inline cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
inline cudaError_t cudaHostGetDevicePointer(void** d, void* h, unsigned) { *d = h; return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) std::memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { if (n) std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { return cudaMemset(d, v, n); }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t)std::malloc(8); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t)std::malloc(8); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }      // a small "device": persistent kernels get 4 x occupancy CTAs
template <class F> inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) { *n = 2; return cudaSuccess; }
template <class S> inline cudaError_t cudaMemcpyToSymbolAsync(S& sym, const void* src, size_t n, size_t off = 0, cudaMemcpyKind = cudaMemcpyHostToDevice, cudaStream_t = nullptr)
{ std::memcpy((char*)&sym + off, src, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dpitch, const void* s, size_t spitch, size_t width, size_t height, cudaMemcpyKind, cudaStream_t = nullptr)
{ for (size_t y = 0; y < height; ++y) std::memmove((char*)d + y * dpitch, (const char*)s + y * spitch, width); return cudaSuccess; }
inline cudaError_t cudaMemcpy2D(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind k) { return cudaMemcpy2DAsync(d, dp, s, sp, w, h, k); }
