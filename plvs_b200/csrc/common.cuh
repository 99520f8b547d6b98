// Shared host/device helpers for libplvs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#ifndef PLVS_CUDA_EMU
#include <cuda.h>            // CUtensorMap + the cuTensorMapEncodeTiled prototype (the function is fetched with cudaGetDriverEntryPoint: no libcuda link)
#endif
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "../../include/plvs_b200.h"

#include <atomic>

namespace plvs {

void set_error(const char* fmt, ...);

// Bytes that crossed the bus, counted where they are copied (bench.py's e2e block reports them): every cudaMemcpyAsync of the library goes
// through counted_memcpy_async (the macro below), kernels that write results straight into mapped host memory add theirs with count_d2h.
extern std::atomic<long long> g_io_bytes[2];       // [0] host->device, [1] device->host
inline void count_h2d(size_t n) { g_io_bytes[0].fetch_add((long long)n, std::memory_order_relaxed); }
inline void count_d2h(size_t n) { g_io_bytes[1].fetch_add((long long)n, std::memory_order_relaxed); }
inline cudaError_t counted_memcpy_async(void* dst, const void* src, size_t n, cudaMemcpyKind kind, cudaStream_t st = 0)
{
    if (kind == cudaMemcpyHostToDevice) count_h2d(n); else if (kind == cudaMemcpyDeviceToHost) count_d2h(n);
    return cudaMemcpyAsync(dst, src, n, kind, st);
}
#define cudaMemcpyAsync(...) ::plvs::counted_memcpy_async(__VA_ARGS__)

#define PLVS_CUDA(expr)                                                                       \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::plvs::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return PLVS_ENODEV;                                                               \
        }                                                                                     \
    } while (0)

// dynamic shared memory of a kernel (the CPU execution model of tests/native/cuda_emu.hpp defines its own before this header is seen)
#ifndef PLVS_DYN_SMEM
#define PLVS_DYN_SMEM(T, name) extern __shared__ T name[]
#define PLVS_DYN_SMEM_ALIGNED(T, name, a) extern __shared__ __align__(a) T name[]
#endif

#ifdef PLVS_CUDA_EMU        // tests/native/cuda_emu.hpp: bulk copies complete at issue on the CPU model
inline uint32_t tma_smem_u32(const void* p) { return emu::smem_handle(p); }
inline void tma_mbar_init(uint32_t, uint32_t) {}
inline void tma_mbar_expect_tx(uint32_t, uint32_t) {}
inline void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t) { std::memcpy(emu::smem_pointer(dst), src, bytes); }
inline void tma_mbar_wait(uint32_t, uint32_t) {}
#elif defined(__CUDACC__)
// 1-D bulk asynchronous copies (the TMA engine; SASS UBLKCP) completed through an mbarrier.  Addresses and sizes must be
// multiples of 16 bytes.
__device__ __forceinline__ uint32_t tma_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile("{\n.reg .pred P1;\nTMA_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra TMA_DONE;\nbra TMA_WAIT;\nTMA_DONE:\n}" ::"r"(bar), "r"(parity) : "memory");
}
#endif

// Stream of a handle.  The three kinds of handle run concurrently in a SLAM process (tracking thread, local mapping, dense mapping) and
// share the SMs.  The matcher's kernels are short and latency-critical (a thread is blocked on every search): highest priority.  The TSDF
// scan is the longest serial kernel chain of a frame and the extractor's wide grids (thousands of CTAs) are what delays it most, so the TSDF
// streams come second and the extractor last (`rank` 2 = matcher, 1 = extractor, 0 = TSDF; PLVS_STREAM_ORDER=orb puts the extractor before the
// TSDF as in round 1, PLVS_STREAM_PRIORITIES=0 gives every stream the same priority).
inline cudaError_t create_handle_stream(cudaStream_t* st, int rank)
{
    const char* e = std::getenv("PLVS_STREAM_PRIORITIES");
    if (e && e[0] == '0') return cudaStreamCreateWithFlags(st, cudaStreamNonBlocking);
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return cudaStreamCreateWithFlags(st, cudaStreamNonBlocking);
    const char* o = std::getenv("PLVS_STREAM_ORDER");
    const bool orb_first = o && o[0] == 'o';
    // numerically lower = higher priority; spread the three ranks over the available range
    const int mid = (least + greatest) / 2;
    const bool tsdf_first = o && o[0] == 't';          // PLVS_STREAM_ORDER=tsdf: TSDF > matcher > extractor (A/B aid)
    int prio;
    if (tsdf_first) prio = rank == 0 ? greatest : rank == 2 ? mid : least;
    else if (rank == 2) prio = greatest;
    else if (rank == 1) prio = orb_first ? mid : least;
    else prio = orb_first ? least : mid;
    return cudaStreamCreateWithPriority(st, cudaStreamNonBlocking, prio);
}

inline int div_up(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// Per-frame buffers are sized by data-dependent counts (keypoints, queries, chunk ranges).  Growing one means
// cudaFree + cudaMalloc (device-wide synchronisation) or cudaHostAlloc (milliseconds), so small buffers grow with
// slack: 25 % on top of the request, at least 1.5x the old size.  Buffers above 64 MiB are allocated exactly.
inline size_t grow_capacity(size_t count, size_t old_n, size_t elem)
{
#ifdef PLVS_CUDA_EMU
    if (std::getenv("PLVS_EMU_GUARD")) return count;       // guard-page runs of the CPU model: no slack that would hide an overrun
#endif
    if (count * elem > ((size_t)64 << 20)) return count;
    size_t c = count + count / 4 + 64;
    if (c < old_n + old_n / 2) c = old_n + old_n / 2;
    return c;
}

// Device memory that is released with the owning handle.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    int alloc(size_t count) {
        if (count <= n && p) return PLVS_OK;
        count = grow_capacity(count, n, sizeof(T));
        release();
        if (cudaMalloc((void**)&p, count * sizeof(T)) != cudaSuccess) {
            p = nullptr; n = 0;
            set_error("cudaMalloc of %zu bytes failed", count * sizeof(T));
            return PLVS_ENOMEM;
        }
        n = count;
        return PLVS_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    ~DevBuf() { release(); }
};

// Pinned, device-mapped host memory: kernels read/write it directly over PCIe (posted writes),
// so small control/result records need no separate memcpy launches.
template <typename T>
struct PinBuf {
    T* h = nullptr;   // host view
    T* d = nullptr;   // device view of the same memory
    size_t n = 0;
    int alloc(size_t count) {
        if (count <= n && h) return PLVS_OK;
        count = grow_capacity(count, n, sizeof(T));
        release();
        if (cudaHostAlloc((void**)&h, count * sizeof(T), cudaHostAllocMapped) != cudaSuccess) {
            h = nullptr; n = 0;
            set_error("cudaHostAlloc of %zu bytes failed", count * sizeof(T));
            return PLVS_ENOMEM;
        }
        if (cudaHostGetDevicePointer((void**)&d, h, 0) != cudaSuccess) { cudaFreeHost(h); h = nullptr; return PLVS_ENODEV; }
        n = count;
        return PLVS_OK;
    }
    void release() { if (h) cudaFreeHost(h); h = nullptr; d = nullptr; n = 0; }
    ~PinBuf() { release(); }
};

// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may become resident while its
// predecessor in the stream still runs; it must not read the predecessor's output before PLVS_GRID_DEP_WAIT().  The predecessor lets it in
// early with PLVS_GRID_DEP_LAUNCH().  Both are no-ops for a plain launch and on the CPU execution model.
#if defined(PLVS_CUDA_EMU)
#define PLVS_GRID_DEP_WAIT() ((void)0)
#define PLVS_GRID_DEP_LAUNCH() ((void)0)
#else
#define PLVS_GRID_DEP_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define PLVS_GRID_DEP_LAUNCH() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#endif

// Optional per-kernel timing with CUDA events on the handle's own stream (bench.py's roofline leg).
// Off by default: when off, begin()/end() are a load and a branch.
extern int g_profiling;
struct KernelTimer {
    static constexpr int kSlots = 12;
    std::vector<cudaEvent_t> pool;
    std::vector<int> pending;      // slot of pair i (events 2i, 2i+1)
    float ms[kSlots] = {0};
    int count[kSlots] = {0};
    int component = 0;             // bit of g_profiling that enables this timer (1 orb, 2 match, 4 tsdf)
    int only_slot = -1;            // g_profiling bit 8 set => only this slot is timed (the roofline kernel)
    bool on(int slot) const { return (g_profiling & component) && (!(g_profiling & 8) || slot == only_slot); }
    void begin(int slot, cudaStream_t st) {
        armed = on(slot);
        if (!armed) return;
        const size_t i = pending.size();
        while (pool.size() < 2 * (i + 1)) { cudaEvent_t e; cudaEventCreate(&e); pool.push_back(e); }
        pending.push_back(slot);
        cudaEventRecord(pool[2 * i], st);
    }
    bool armed = false;
    void end(cudaStream_t st) { if (!armed || pending.empty()) return; cudaEventRecord(pool[2 * (pending.size() - 1) + 1], st); armed = false; }
    void collect() {           // call after the stream has been synchronised
        for (size_t i = 0; i < pending.size(); ++i) {
            float t = 0.f;
            if (cudaEventElapsedTime(&t, pool[2 * i], pool[2 * i + 1]) == cudaSuccess) { ms[pending[i]] += t; ++count[pending[i]]; }
        }
        pending.clear();
    }
    void reset() { for (int i = 0; i < kSlots; ++i) { ms[i] = 0.f; count[i] = 0; } pending.clear(); }
    ~KernelTimer() { for (cudaEvent_t e : pool) cudaEventDestroy(e); }
};

}  // namespace plvs
