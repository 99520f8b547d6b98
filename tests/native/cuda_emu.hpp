// TEST INFRASTRUCTURE ONLY -- a small CPU execution model for CUDA device code, so that kernels which have not met a GPU yet can be
// run in this container: tests/native/emu_kernels.cpp includes the product's device-only headers (plvs_b200/csrc/*.cuh) after this
// file and drives them through emu::launch.  Nothing here is linked into libplvs_b200.so, and nothing in the product falls back to it.
//
// Model: one CTA at a time; every CUDA thread is a fiber (own stack, hand-written switch) on the calling OS thread (so "atomics" are plain operations and
// __shared__ variables are function-level statics), scheduled round-robin.  __syncthreads / __syncwarp / __shfl_*_sync / __ballot_sync
// are rendezvous points: a fiber that reaches one yields until every live thread of the CTA (or every live lane named by the mask) has
// arrived.  A rendezvous that can never complete (divergent barrier, a shuffle some lane skips) is reported as a deadlock instead of
// hanging.  Floating point: x86-64 SSE arithmetic is IEEE like the device's with --fmad=false; compile with -ffp-contract=off.
// Clusters: the CTAs of one cluster are resident together (launch_cluster) and meet at emu::cluster_barrier.
// Scheduling: round-robin by default; PLVS_EMU_SCHED_SEED=n randomises the order in which thread segments run (a race detector of sorts).
// Not modelled: distributed shared memory, TMA / mbarrier, tensor cores, memory ordering weaker than sequential consistency, scheduling races.
#pragma once
#define PLVS_CUDA_EMU 1

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <vector>

struct dim3 { unsigned x, y, z; dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {} };
struct uint4 { uint32_t x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct int2 { int x, y; };
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float3 make_float3(float x, float y, float z) { return float3{x, y, z}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }

namespace emu {

// Context switch: callee-saved registers pushed on the old stack, stack pointers exchanged (glibc's swapcontext would add two signal-mask system
// calls per switch, which dominated the run time).  x86-64 System V only; the symbol is weak because every emulated unit carries a copy.
#if !defined(__x86_64__)
#error "tests/native/cuda_emu.hpp: the fiber switch is written for x86-64"
#endif
extern "C" void plvs_emu_switch(void** save_sp, void* load_sp);
asm(".text\n.weak plvs_emu_switch\n.type plvs_emu_switch,@function\nplvs_emu_switch:\n"
    "pushq %rbp\npushq %rbx\npushq %r12\npushq %r13\npushq %r14\npushq %r15\n"
    "movq %rsp, (%rdi)\nmovq %rsi, %rsp\n"
    "popq %r15\npopq %r14\npopq %r13\npopq %r12\npopq %rbx\npopq %rbp\nret\n");

struct Fiber {
    void* sp = nullptr;            // saved stack pointer while the fiber is switched out
    char* stack = nullptr;         // borrowed from the pool below: allocating and releasing 128 KiB per thread and launch dominated the run time
    dim3 tid;
    dim3 bid;
    int linear = 0;                // thread index inside its CTA
    int cta = 0;                   // CTA index inside the resident group (0 unless a cluster is being run)
    bool done = false;
};

struct Warp { uint64_t vals[32]; uint32_t preds = 0; int arrived = 0; unsigned gen = 0; uint32_t alive = 0; };

struct State {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    Fiber* cur = nullptr;
    void* sched_sp = nullptr;
    dim3 bdim, gdim;
    int alive = 0;                                  // live fibers of the resident group
    std::vector<int> cta_alive, bar_arrived;        // per resident CTA
    std::vector<unsigned> bar_gen;
    int cluster = 1, cl_arrived = 0;                // CTAs resident together; cluster barrier state
    unsigned cl_gen = 0;
    int warps_per_cta = 0;
    size_t dyn_stride = 0;
    long idle = 0;                 // consecutive yields without progress: deadlock detector
    std::function<void()>* body = nullptr;
    std::vector<char> dyn;
    char* dyn_smem = nullptr;
    const char* failure = nullptr;
};
inline State& st() { static State s; return s; }
inline char* pooled_stack(size_t index, size_t bytes)
{
    static std::vector<std::unique_ptr<char[]>> pool;
    if (pool.size() <= index) pool.resize(index + 1);
    if (!pool[index]) pool[index].reset(new char[bytes]);
    return pool[index].get();
}

inline void yield()
{
    State& g = st();
    if (++g.idle > 64L * (long)g.fibers.size() + 4096) { g.failure = "deadlock: a barrier / warp rendezvous can never complete"; }
    plvs_emu_switch(&g.cur->sp, g.sched_sp);
}

inline void trampoline()
{
    State& g = st();
    (*g.body)();
    Fiber* f = g.cur;
    f->done = true; g.idle = 0;
    --g.alive; --g.cta_alive[f->cta];
    g.warps[f->cta * g.warps_per_cta + (f->linear >> 5)].alive &= ~(1u << (f->linear & 31));
    plvs_emu_switch(&f->sp, g.sched_sp);
    std::abort();                  // a finished fiber is never resumed
}

// `cluster` CTAs (consecutive in x) are resident together and may meet at cluster_barrier(); 1 = ordinary launch.  Limits of the cluster mode:
// function-level __shared__ variables are one copy per launch (a cluster kernel must keep per-CTA state in dynamic shared memory, or touch its
// static shared variables from one CTA only), and distributed shared memory is not modelled.
template <class F>
void launch_cluster(int cluster, dim3 grid, dim3 block, size_t smem_bytes, F&& kernel_call)
{
    State& g = st();
    std::function<void()> body = kernel_call;
    g.body = &body; g.bdim = block; g.gdim = grid; g.cluster = cluster;
    const int nthreads = (int)(block.x * block.y * block.z);
    const size_t stack_bytes = 128 * 1024;
    g.warps_per_cta = (nthreads + 31) / 32;
    g.dyn_stride = (smem_bytes + 16 + 127) / 128 * 128;
    g.dyn.assign(g.dyn_stride * cluster + 256, 0);
    g.dyn_smem = g.dyn.data() + (256 - (reinterpret_cast<uintptr_t>(g.dyn.data()) & 255)) % 256;          // 256-byte aligned like the device's window
    if (grid.x % cluster) throw std::runtime_error("grid.x is not a multiple of the cluster size");
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; bx += cluster) {
        const int total = nthreads * cluster;
        g.fibers.clear(); g.fibers.resize(total);
        g.warps.assign((size_t)g.warps_per_cta * cluster, Warp());
        g.alive = total; g.idle = 0; g.failure = nullptr; g.cl_arrived = 0;
        g.cta_alive.assign(cluster, nthreads); g.bar_arrived.assign(cluster, 0); g.bar_gen.assign(cluster, 0u);
        for (int i = 0; i < total; ++i) {
            Fiber& f = g.fibers[i];
            const int t = i % nthreads;
            f.cta = i / nthreads; f.bid = dim3(bx + f.cta, by, bz);
            f.linear = t; f.tid = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y)); f.done = false;
            f.stack = pooled_stack((size_t)i, stack_bytes);
            // first switch-in: six zeroed callee-saved registers, then `ret` into trampoline with the stack as a call would leave it (rsp = 8 mod 16)
            uintptr_t top = (reinterpret_cast<uintptr_t>(f.stack) + stack_bytes) & ~(uintptr_t)15;
            void** sp = reinterpret_cast<void**>(top - 64);
            for (int k = 0; k < 6; ++k) sp[k] = nullptr;
            sp[6] = reinterpret_cast<void*>(&trampoline);
            sp[7] = nullptr;
            f.sp = sp;
            g.warps[f.cta * g.warps_per_cta + (t >> 5)].alive |= 1u << (t & 31);
        }
        // PLVS_EMU_SCHED_SEED=n: instead of round-robin, every sweep visits the live threads in a fresh pseudo-random order and leaves a random
        // quarter of them out -- between two rendezvous points any interleaving of whole thread segments can then occur, so a missing barrier
        // (a shared-memory producer / consumer pair that only works in thread order) shows up as a wrong result under some seed
        static const char* seed_env = std::getenv("PLVS_EMU_SCHED_SEED");
        uint64_t rng = seed_env ? 0x9E3779B97F4A7C15ull * (uint64_t)(std::atoll(seed_env) + 1) : 0;
        std::vector<int> order(total);
        for (int t = 0; t < total; ++t) order[t] = t;
        while (g.alive > 0) {
            if (seed_env) for (int t = total - 1; t > 0; --t) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; std::swap(order[t], order[(size_t)(rng % (uint64_t)(t + 1))]); }
            for (int k = 0; k < total && g.alive > 0; ++k) {
                Fiber& f = g.fibers[order[k]];
                if (f.done) continue;
                if (seed_env) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; if ((rng & 3u) == 0) continue; }
                g.cur = &f;
                plvs_emu_switch(&g.sched_sp, f.sp);
                if (g.failure) { const char* why = g.failure; g.failure = nullptr; g.fibers.clear(); throw std::runtime_error(why); }
            }
        }
        g.fibers.clear();
    }
    g.cluster = 1;
}

template <class F>
void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& kernel_call) { launch_cluster(1, grid, block, smem_bytes, kernel_call); }

inline void cta_barrier()
{
    State& g = st();
    const int c = g.cur->cta;
    const unsigned my = g.bar_gen[c];
    ++g.bar_arrived[c];
    while (g.bar_gen[c] == my) {
        if (g.bar_arrived[c] >= g.cta_alive[c]) { g.bar_arrived[c] = 0; ++g.bar_gen[c]; g.idle = 0; break; }
        yield();
    }
}

// barrier.cluster.arrive + wait over every live thread of the resident CTAs
inline void cluster_barrier()
{
    State& g = st();
    const unsigned my = g.cl_gen;
    ++g.cl_arrived;
    while (g.cl_gen == my) {
        if (g.cl_arrived >= g.alive) { g.cl_arrived = 0; ++g.cl_gen; g.idle = 0; break; }
        yield();
    }
}
inline unsigned cluster_rank() { return (unsigned)st().cur->cta; }

// 32-bit "shared-memory addresses" (what __cvta_generic_to_shared yields on the device) for code that does address arithmetic on them:
// region id in the top byte, byte offset below; a region is a 16 MiB window starting at the first pointer seen in it
inline std::vector<const char*>& smem_regions() { static std::vector<const char*> r; return r; }
inline uint32_t smem_handle(const void* p)
{
    auto& r = smem_regions();
    const char* c = (const char*)p;
    for (size_t i = 0; i < r.size(); ++i) if (c >= r[i] && c < r[i] + (1u << 24)) return (uint32_t)((i + 1) << 24) | (uint32_t)(c - r[i]);
    if (r.size() >= 254) r.clear();
    r.push_back(c);
    return (uint32_t)(r.size() << 24);
}
inline void* smem_pointer(uint32_t h) { return (void*)(smem_regions()[(h >> 24) - 1] + (h & 0xffffffu)); }

// __syncthreads_or: the barrier plus an OR over the CTA
inline int cta_barrier_or(int pred)
{
    static int acc[64];
    const int c = st().cur->cta;
    if (pred) acc[c] = 1;
    cta_barrier();                 // every contribution is in
    const int r = acc[c];
    cta_barrier();                 // every thread has read it
    acc[c] = 0;
    cta_barrier();                 // cleared before anybody can contribute to the next one
    return r;
}

inline Warp& my_warp() { State& g = st(); return g.warps[g.cur->cta * g.warps_per_cta + (g.cur->linear >> 5)]; }
inline int lane_id() { return st().cur->linear & 31; }

inline void warp_barrier(uint32_t mask)
{
    Warp& w = my_warp();
    const unsigned my = w.gen;
    ++w.arrived;
    while (w.gen == my) {
        if (w.arrived >= __builtin_popcount(mask & w.alive)) { w.arrived = 0; ++w.gen; st().idle = 0; break; }
        yield();
    }
}

template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle payload"); std::memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; std::memcpy(&v, &b, sizeof(T)); return v; }

template <class T> inline T shfl(uint32_t mask, T v, int src)
{
    Warp& w = my_warp();
    w.vals[lane_id()] = to_bits(v);
    warp_barrier(mask);
    const T r = (src >= 0 && src < 32 && ((w.alive >> src) & 1u)) ? from_bits<T>(w.vals[src]) : v;
    warp_barrier(mask);
    return r;
}

}  // namespace emu

// ---- the CUDA surface the device-only headers use -------------------------------------------------------------------------------
#define __grid_constant__
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))
#define __cluster_dims__(...)
#define PLVS_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(emu::st().dyn_smem + emu::st().dyn_stride * emu::st().cur->cta)
#define PLVS_DYN_SMEM_ALIGNED(T, name, a) PLVS_DYN_SMEM(T, name)
#define threadIdx (emu::st().cur->tid)
#define blockIdx (emu::st().cur->bid)
#define blockDim (emu::st().bdim)
#define gridDim (emu::st().gdim)

using std::max;
using std::min;
using std::isnan;
using std::isinf;

inline void __syncthreads() { emu::cta_barrier(); }
inline int __syncthreads_or(int pred) { return emu::cta_barrier_or(pred); }
inline long long clock64() { static long long t = 0; return ++t; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline float __fsqrt_rn(float a) { return std::sqrt(a); }
inline int __float2int_rd(float a) { return (int)std::floor(a); }
inline int __float2int_rn(float a) { return (int)std::nearbyint(a); }
inline int __float2int_rz(float a) { return (int)a; }
inline float __int2float_rn(int a) { return (float)a; }
inline uint32_t __float_as_uint(float a) { uint32_t u; std::memcpy(&u, &a, 4); return u; }
inline float __uint_as_float(uint32_t u) { float a; std::memcpy(&a, &u, 4); return a; }
inline int __float_as_int(float a) { int u; std::memcpy(&u, &a, 4); return u; }
inline float __int_as_float(int u) { float a; std::memcpy(&a, &u, 4); return a; }
inline float __saturatef(float a) { return a < 0.f ? 0.f : a > 1.f ? 1.f : a; }
inline void __syncwarp(uint32_t mask = 0xffffffffu) { emu::warp_barrier(mask); }
inline void __threadfence() {}
template <class T> inline T __shfl_sync(uint32_t m, T v, int src) { return emu::shfl(m, v, src); }
template <class T> inline T __shfl_xor_sync(uint32_t m, T v, int x) { return emu::shfl(m, v, emu::lane_id() ^ x); }
template <class T> inline T __shfl_up_sync(uint32_t m, T v, unsigned d) { const int l = emu::lane_id(); return emu::shfl(m, v, l >= (int)d ? l - (int)d : l); }
template <class T> inline T __shfl_down_sync(uint32_t m, T v, unsigned d) { const int l = emu::lane_id(); return emu::shfl(m, v, l + (int)d < 32 ? l + (int)d : l); }
inline uint32_t __ballot_sync(uint32_t mask, bool pred)
{
    emu::Warp& w = emu::my_warp();
    emu::warp_barrier(mask);                      // the previous ballot's readers are done
    if (pred) w.preds |= 1u << emu::lane_id(); else w.preds &= ~(1u << emu::lane_id());
    emu::warp_barrier(mask);
    const uint32_t r = w.preds & mask & w.alive;
    emu::warp_barrier(mask);
    return r;
}
template <class T> inline T __reduce_min_sync(uint32_t mask, T v)
{
    T r = v;
    for (int o = 16; o; o >>= 1) { const T w = emu::shfl(mask, r, emu::lane_id() ^ o); if (w < r) r = w; }
    return r;
}
template <class T> inline T __reduce_max_sync(uint32_t mask, T v)
{
    T r = v;
    for (int o = 16; o; o >>= 1) { const T w = emu::shfl(mask, r, emu::lane_id() ^ o); if (w > r) r = w; }
    return r;
}
template <class T> inline T __reduce_add_sync(uint32_t mask, T v)
{
    T r = v;
    for (int o = 16; o; o >>= 1) r += emu::shfl(mask, r, emu::lane_id() ^ o);
    return r;
}
template <class T> inline uint32_t __match_any_sync(uint32_t mask, T v)
{
    uint32_t r = 0;
    for (int j = 0; j < 32; ++j) { const T o = emu::shfl(mask, v, j); if (((mask >> j) & 1u) && o == v) r |= 1u << j; }
    return r & emu::my_warp().alive;
}
inline int __any_sync(uint32_t mask, int pred) { return __ballot_sync(mask, pred != 0) != 0; }
inline int __all_sync(uint32_t mask, int pred) { emu::Warp& w = emu::my_warp(); const uint32_t b = __ballot_sync(mask, pred != 0); return b == (mask & w.alive); }
template <class T> inline T __reduce_or_sync(uint32_t mask, T v)
{
    T r = v;
    for (int o = 16; o; o >>= 1) r |= emu::shfl(mask, r, emu::lane_id() ^ o);
    return r;
}
template <class T> inline T __reduce_and_sync(uint32_t mask, T v)
{
    T r = v;
    for (int o = 16; o; o >>= 1) r &= emu::shfl(mask, r, emu::lane_id() ^ o);
    return r;
}
inline int __popc(uint32_t v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <class T> inline T atomicAdd(T* p, T v) { const T o = *p; *p = o + v; return o; }
template <class T> inline T atomicSub(T* p, T v) { const T o = *p; *p = o - v; return o; }
template <class T> inline T atomicExch(T* p, T v) { const T o = *p; *p = v; return o; }
template <class T> inline T atomicCAS(T* p, T cmp, T v) { const T o = *p; if (o == cmp) *p = v; return o; }
template <class T> inline T atomicMax(T* p, T v) { const T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicOr(T* p, T v) { const T o = *p; *p = o | v; return o; }
template <class T> inline T atomicAnd(T* p, T v) { const T o = *p; *p = o & v; return o; }
template <class T> inline T atomicMin(T* p, T v) { const T o = *p; if (v < o) *p = v; return o; }
