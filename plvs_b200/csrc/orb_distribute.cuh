// Device keypoint distributor: PLVS2::ORBextractor::DistributeOctTree (src/ORBextractor.cc:611-865,
// DivideNode :536-592, compareNodes :594-609) as ONE CTA per (pyramid level, frame).
//
// The reference is a sequential std::list algorithm whose iteration order is observable (it is the order of the
// returned keypoints).  What it computes, restated so that it parallelises:
//   * A node = (UL.x, UL.y, UR.x, BR.y) + a contiguous range of a permutation of the candidates; a split is a STABLE
//     4-way partition of that range (so "first maximum response wins" sees the reference's order inside a node).
//   * Full sweeps: every node with >1 point splits; children are pushed to the list FRONT, the parent is erased.
//     With the list stored back-to-front a sweep is  new_list = kept-leaves (old order) ++ children (creation order).
//     All splits of a sweep are ONE segmented stable partition: two 64-bit prefix scans over the permutation
//     (quadrant counters packed 2 x 32 bit) + a scatter; child slots come from prefix scans over the node list.
//   * Partial phase (entered when another full sweep would overshoot the quota): the expandable nodes of the last
//     sweep are sorted by (size, UL.x) with std::sort -- the comparator is not a total order, so ties land wherever
//     libstdc++'s introsort puts them: one thread runs an exact emulation (stdsort_emul.cuh) -- and split largest
//     first until the node count reaches the quota.  How many get split is a prefix-sum question (a split adds
//     #non-empty children - 1 nodes), so the splits themselves run in parallel again.
//   * Result: per surviving node, in list order, the first maximum-response candidate.
// Checked bit-for-bit against the host implementation (orb_distribute.hpp) and the list-based oracle.
#pragma once
#include "orb_kernels.cuh"
#include "stdsort_emul.cuh"

namespace plvs {
namespace orb {

constexpr int kDistThreads = 1024;

struct DNode {
    short ulx, uly, urx, bry;
    int begin, count;
    int kid[4];        // child node ids of the split (-1 = empty quadrant)
    int leaf;          // bNoMore
    int dead;          // erased from the list
    int tag;           // == current split tag while this node is being split
};

struct DistLevel {                 // per (frame, level) scratch, all device pointers
    int* perm[2];                  // candidate permutation (ping-pong), n_max
    int* node_of[2];               // node id owning each permutation slot, n_max
    unsigned long long* scan_a;    // packed quadrant counters / prefix sums, n_max + 1
    unsigned long long* scan_b;
    DNode* nodes;                  // node pool, ncap
    int* order[2];                 // node list stored back-to-front, ncap
    int* plist;                    // parents of the current split in traversal order, ncap
    int* nkids;                    // per parent / scratch, ncap
    int* nexp;                     // per parent / scratch, ncap
    int* flag;                     // per list entry scratch, ncap
    unsigned long long* expand[2]; // expandable nodes in creation order: ((size<<12 | ulx) << 32) | node id, ncap
    uint32_t* stage;               // picked candidates of this level, ncap
    int ncap;
};

// ---- block-wide exclusive scans, in place over global arrays (n may exceed the block size) -------------------
__device__ inline unsigned long long block_scan_u64(unsigned long long* a, int n, unsigned long long* s_part /*32*/)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    unsigned long long running = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        const unsigned long long v = i < n ? a[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)x, o), hi = __shfl_up_sync(0xffffffffu, (uint32_t)(x >> 32), o);
            if (lane >= o) x += ((unsigned long long)hi << 32) | lo;
        }
        if (lane == 31) s_part[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned long long p = lane < nw ? s_part[lane] : 0ull;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t lo = __shfl_up_sync(0xffffffffu, (uint32_t)p, o), hi = __shfl_up_sync(0xffffffffu, (uint32_t)(p >> 32), o);
                if (lane >= o) p += ((unsigned long long)hi << 32) | lo;
            }
            s_part[lane] = p;
        }
        __syncthreads();
        const unsigned long long wbase = wid ? s_part[wid - 1] : 0ull;
        if (i < n) a[i] = running + wbase + x - v;
        const unsigned long long chunk = s_part[nw - 1];
        __syncthreads();
        running += chunk;
    }
    return running;
}

__device__ inline int block_scan_i32(int* a, int n, int* s_part /*32*/)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    int running = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        const int v = i < n ? a[i] : 0;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_part[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int p = lane < nw ? s_part[lane] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, p, o); if (lane >= o) p += y; }
            s_part[lane] = p;
        }
        __syncthreads();
        const int wbase = wid ? s_part[wid - 1] : 0;
        if (i < n) a[i] = running + wbase + x - v;
        const int chunk = s_part[nw - 1];
        __syncthreads();
        running += chunk;
    }
    return running;
}

__device__ __forceinline__ int quadrant_of(const DNode& nd, int x, int y)
{
    const int mx = nd.ulx + ((nd.urx - nd.ulx + 1) >> 1);      // UL.x + ceil((UR.x-UL.x)/2)
    const int my = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
    return (x < mx ? 0 : 1) + (y < my ? 0 : 2);
}

// the same from the node's first eight bytes (ulx, uly | urx, bry as two 32-bit words)
__device__ __forceinline__ int quadrant_of_packed(uint2 b, int x, int y)
{
    const int ulx = (short)(b.x & 0xffffu), uly = (short)(b.x >> 16), urx = (short)(b.y & 0xffffu), bry = (short)(b.y >> 16);
    const int mx = ulx + ((urx - ulx + 1) >> 1), my = uly + ((bry - uly + 1) >> 1);
    return (x < mx ? 0 : 1) + (y < my ? 0 : 2);
}

__device__ __forceinline__ void child_counts(const DistLevel& D, const DNode& nd, int c[4])
{
    const unsigned long long da = D.scan_a[nd.begin + nd.count] - D.scan_a[nd.begin], db = D.scan_b[nd.begin + nd.count] - D.scan_b[nd.begin];
    c[0] = (int)(uint32_t)da; c[1] = (int)(da >> 32); c[2] = (int)(uint32_t)db; c[3] = (int)(db >> 32);
}

// Exclusive prefix sums of TWO packed-counter arrays at once, in place, over n elements.  Every warp owns one contiguous segment and carries
// its running sums in registers (32 elements per step, shuffles only); the 32 segment totals are scanned once by warp 0 and added back in a
// second coalesced pass: two CTA barriers in all, however long the arrays are.
__device__ inline void block_scan2_u64(unsigned long long* __restrict__ a, unsigned long long* __restrict__ b, int n, unsigned long long* s_part /*64*/)
{
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    const int seg = ((n + nw - 1) / nw + 31) & ~31;                 // elements per warp, a multiple of 32
    const int beg = wid * seg, end = min(beg + seg, n);
    unsigned long long ra = 0, rb = 0;
    for (int base = beg; base < end; base += 32) {
        const int i = base + lane;
        const unsigned long long va = i < end ? a[i] : 0ull, vb = i < end ? b[i] : 0ull;
        unsigned long long xa = va, xb = vb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t al = __shfl_up_sync(0xffffffffu, (uint32_t)xa, o), ah = __shfl_up_sync(0xffffffffu, (uint32_t)(xa >> 32), o);
            const uint32_t bl = __shfl_up_sync(0xffffffffu, (uint32_t)xb, o), bh = __shfl_up_sync(0xffffffffu, (uint32_t)(xb >> 32), o);
            if (lane >= o) { xa += ((unsigned long long)ah << 32) | al; xb += ((unsigned long long)bh << 32) | bl; }
        }
        if (i < end) { a[i] = ra + xa - va; b[i] = rb + xb - vb; }
        const uint32_t tal = __shfl_sync(0xffffffffu, (uint32_t)xa, 31), tah = __shfl_sync(0xffffffffu, (uint32_t)(xa >> 32), 31);
        const uint32_t tbl = __shfl_sync(0xffffffffu, (uint32_t)xb, 31), tbh = __shfl_sync(0xffffffffu, (uint32_t)(xb >> 32), 31);
        ra += ((unsigned long long)tah << 32) | tal; rb += ((unsigned long long)tbh << 32) | tbl;
    }
    if (lane == 0) { s_part[wid] = ra; s_part[32 + wid] = rb; }
    __syncthreads();
    if (wid == 0) {
        const unsigned long long va = lane < nw ? s_part[lane] : 0ull, vb = lane < nw ? s_part[32 + lane] : 0ull;
        unsigned long long xa = va, xb = vb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t al = __shfl_up_sync(0xffffffffu, (uint32_t)xa, o), ah = __shfl_up_sync(0xffffffffu, (uint32_t)(xa >> 32), o);
            const uint32_t bl = __shfl_up_sync(0xffffffffu, (uint32_t)xb, o), bh = __shfl_up_sync(0xffffffffu, (uint32_t)(xb >> 32), o);
            if (lane >= o) { xa += ((unsigned long long)ah << 32) | al; xb += ((unsigned long long)bh << 32) | bl; }
        }
        s_part[lane] = xa - va; s_part[32 + lane] = xb - vb;           // exclusive segment bases
    }
    __syncthreads();
    const unsigned long long ba = s_part[wid], bb = s_part[32 + wid];
    if (wid > 0 && (ba | bb))
        for (int i = beg + lane; i < end; i += 32) { a[i] += ba; b[i] += bb; }
    __syncthreads();
}

// quadrant prefix sums over all permutation slots whose node carries `tag`
__device__ inline void quadrant_scans(const uint32_t* __restrict__ cand, int n, const DistLevel& D, int cur, int tag, unsigned long long* s_u64)
{
    const int tid = threadIdx.x, T = blockDim.x;
    const int* __restrict__ perm = D.perm[cur]; const int* __restrict__ nof = D.node_of[cur];
    const DNode* __restrict__ nodes = D.nodes;
    unsigned long long* __restrict__ sa = D.scan_a; unsigned long long* __restrict__ sb = D.scan_b;
    // four slots per thread and step: the three dependent loads (owner -> node, slot -> candidate) of the four are in flight together
    for (int p0 = tid; p0 <= n; p0 += 4 * T) {
        int k[4], ci[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int p = p0 + u * T; k[u] = p < n ? nof[p] : -1; ci[u] = p < n ? perm[p] : 0; }
        uint2 bd[4]; int tg[4]; uint32_t c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            bd[u] = make_uint2(0u, 0u); tg[u] = tag - 1; c[u] = 0u;
            if (k[u] >= 0) { { const uint32_t* nw32 = reinterpret_cast<const uint32_t*>(&nodes[k[u]]); bd[u] = make_uint2(nw32[0], nw32[1]); }      /* DNode is 44 bytes: 4-byte aligned only */ tg[u] = nodes[k[u]].tag; c[u] = cand[ci[u]]; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = p0 + u * T;
            if (p > n) continue;
            unsigned long long a = 0, b = 0;
            if (tg[u] == tag) {
                const int q = quadrant_of_packed(bd[u], unpack_x(c[u]) - kRoiMargin, unpack_y(c[u]) - kRoiMargin);
                if (q == 0) a = 1ull; else if (q == 1) a = 1ull << 32; else if (q == 2) b = 1ull; else b = 1ull << 32;
            }
            sa[p] = a; sb[p] = b;
        }
    }
    __syncthreads();
    block_scan2_u64(D.scan_a, D.scan_b, n + 1, s_u64);
}

// Splits plist[0..nparents) (all carrying `tag`, quadrant_scans already done) at once.  Children are created in the
// order of plist x quadrant (n1,n2,n3,n4), ids from *node_count; their ids go to child_order[], the ones with more
// than one point to expand_out[] (same order).  perm/node_of are rewritten into the other buffer for ALL slots.
__device__ inline void split_nodes(const uint32_t* __restrict__ cand, int n, DistLevel& D, int cur, int tag, int nparents, int* node_count,
                                   int* s_i32, int* total_kids_out, int* total_exp_out, unsigned long long* expand_out, int* child_order)
{
    const int tid = threadIdx.x, T = blockDim.x;
    for (int j = tid; j < nparents; j += T) {
        int c[4];
        child_counts(D, D.nodes[D.plist[j]], c);
        D.nkids[j] = (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0);
        D.nexp[j] = (c[0] > 1) + (c[1] > 1) + (c[2] > 1) + (c[3] > 1);
    }
    __syncthreads();
    const int total_kids = block_scan_i32(D.nkids, nparents, s_i32);
    __syncthreads();
    const int total_exp = block_scan_i32(D.nexp, nparents, s_i32);
    __syncthreads();
    const int base_id = *node_count;
    for (int j = tid; j < nparents; j += T) {
        const int pid = D.plist[j];
        DNode nd = D.nodes[pid];
        int c[4];
        child_counts(D, nd, c);
        const int mx = nd.ulx + ((nd.urx - nd.ulx + 1) >> 1), my = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
        int k = D.nkids[j], e = D.nexp[j], b = nd.begin;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (c[q] == 0) { nd.kid[q] = -1; continue; }
            DNode ch;
            ch.ulx = (short)((q & 1) ? mx : nd.ulx); ch.urx = (short)((q & 1) ? nd.urx : mx);
            ch.uly = (short)((q & 2) ? my : nd.uly); ch.bry = (short)((q & 2) ? nd.bry : my);
            ch.begin = b; ch.count = c[q]; ch.leaf = c[q] == 1; ch.dead = 0; ch.tag = 0;
            ch.kid[0] = ch.kid[1] = ch.kid[2] = ch.kid[3] = -1;
            const int id = base_id + k;
            D.nodes[id] = ch;
            nd.kid[q] = id;
            child_order[k] = id;
            if (c[q] > 1) { expand_out[e] = ((unsigned long long)(((unsigned)min(c[q], (1 << 20) - 1) << 12) | (unsigned)ch.ulx) << 32) | (unsigned)id; ++e; }
            ++k; b += c[q];
        }
        nd.dead = 1;
        D.nodes[pid] = nd;
    }
    __syncthreads();
    const int* __restrict__ perm = D.perm[cur]; const int* __restrict__ nof = D.node_of[cur];
    int* __restrict__ perm2 = D.perm[cur ^ 1]; int* __restrict__ nof2 = D.node_of[cur ^ 1];
    const DNode* __restrict__ nodes = D.nodes;
    const unsigned long long* __restrict__ sa = D.scan_a; const unsigned long long* __restrict__ sb = D.scan_b;
    for (int p0 = tid; p0 < n; p0 += 4 * T) {           // four slots per thread and step, their dependent loads in flight together
        int k[4], ci[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int p = p0 + u * T; k[u] = p < n ? nof[p] : -1; ci[u] = p < n ? perm[p] : 0; }
        uint2 bd[4]; int tg[4], nb[4]; uint32_t c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            bd[u] = make_uint2(0u, 0u); tg[u] = tag - 1; nb[u] = 0; c[u] = 0u;
            if (k[u] >= 0) { { const uint32_t* nw32 = reinterpret_cast<const uint32_t*>(&nodes[k[u]]); bd[u] = make_uint2(nw32[0], nw32[1]); }      /* DNode is 44 bytes: 4-byte aligned only */ tg[u] = nodes[k[u]].tag; nb[u] = nodes[k[u]].begin; c[u] = cand[ci[u]]; }
        }
        int q[4], cid[4], r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            q[u] = -1; cid[u] = 0; r[u] = 0;
            if (tg[u] != tag) continue;
            q[u] = quadrant_of_packed(bd[u], unpack_x(c[u]) - kRoiMargin, unpack_y(c[u]) - kRoiMargin);
            const int p = p0 + u * T;
            const unsigned long long sdf = q[u] < 2 ? sa[p] - sa[nb[u]] : sb[p] - sb[nb[u]];
            r[u] = (q[u] & 1) ? (int)(sdf >> 32) : (int)(uint32_t)sdf;
            cid[u] = nodes[k[u]].kid[q[u]];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) if (q[u] >= 0) r[u] += nodes[cid[u]].begin;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = p0 + u * T;
            if (k[u] < 0) continue;
            if (q[u] < 0) { perm2[p] = ci[u]; nof2[p] = k[u]; continue; }
            perm2[r[u]] = ci[u]; nof2[r[u]] = cid[u];
        }
    }
    __syncthreads();
    if (tid == 0) { *node_count = base_id + total_kids; *total_kids_out = total_kids; *total_exp_out = total_exp; }
    __syncthreads();
}

struct DistArgs {
    const uint32_t* cand;          // compacted candidates, frame stride = slots_per_frame
    const int* cand_count;         // [frame][level]
    long long slots_per_frame;
    const LevelGeom* levels;
    int nlevels;
    const int* quota;              // per level
    const DistLevel* scratch;      // [frame][level]
    uint32_t* sel;                 // out: selected candidates, frame stride = sel_cap, level-major
    int* sel_count;                // out: [frame][level]
    int sel_cap;
    int* error;                    // != 0 on node-pool / keypoint-capacity overflow
};

__global__ void __launch_bounds__(kDistThreads)
k_distribute(DistArgs A)
{
    PLVS_DYN_SMEM(unsigned long long, s_sort);           // max quota + 8 elements for the std::sort emulation
    __shared__ int s_i32[32];
    __shared__ unsigned long long s_u64[64];
    __shared__ int s_nc, s_nk, s_ne, s_live, s_flag;
    const int level = blockIdx.x, frame = blockIdx.y, tid = threadIdx.x, T = blockDim.x;
    const LevelGeom g = A.levels[level];
    const int n = A.cand_count[frame * A.nlevels + level];
    const int N = A.quota[level];
    const uint32_t* cand = A.cand + (long long)frame * A.slots_per_frame + g.slot_begin;
    DistLevel D = A.scratch[frame * A.nlevels + level];
    int* out_count = &A.sel_count[frame * A.nlevels + level];
    const int minX = kRoiMargin, maxX = g.w - kEdge + 3, minY = kRoiMargin, maxY = g.h - kEdge + 3;
    const int nIni = (int)roundf((float)(maxX - minX) / (float)(maxY - minY));
    if (n == 0 || nIni == 0 || nIni > D.ncap) { if (tid == 0) *out_count = 0; return; }
    const float hX = (float)(maxX - minX) / (float)nIni;

    // ---- roots (src/ORBextractor.cc:626-664): stable bucketing by root index, one scan per root (nIni is 1-3)
    int cur = 0;
    for (int r = tid; r < nIni; r += T) {
        DNode nd;
        nd.ulx = (short)(int)(hX * (float)r); nd.uly = 0; nd.urx = (short)(int)(hX * (float)(r + 1)); nd.bry = (short)(maxY - minY);
        nd.begin = 0; nd.count = 0; nd.leaf = 0; nd.dead = 0; nd.tag = 0; nd.kid[0] = nd.kid[1] = nd.kid[2] = nd.kid[3] = -1;
        D.nodes[r] = nd;
    }
    __syncthreads();
    {
        int placed = 0;
        for (int r = 0; r < nIni; ++r) {
            for (int p = tid; p <= n; p += T)
                D.scan_a[p] = (p < n && (int)((float)(unpack_x(cand[p]) - kRoiMargin) / hX) == r) ? 1ull : 0ull;
            __syncthreads();
            const int cnt = (int)block_scan_u64(D.scan_a, n + 1, s_u64);
            __syncthreads();
            for (int p = tid; p < n; p += T)
                if ((int)((float)(unpack_x(cand[p]) - kRoiMargin) / hX) == r) { const int pos = placed + (int)D.scan_a[p]; D.perm[cur][pos] = p; D.node_of[cur][pos] = r; }
            if (tid == 0) { D.nodes[r].begin = placed; D.nodes[r].count = cnt; D.nodes[r].leaf = cnt == 1; }
            placed += cnt;
            __syncthreads();
        }
    }
    if (tid == 0) {      // list = roots 0..nIni-1 front to back, empty ones erased; stored back-to-front
        int L = 0;
        for (int r = nIni - 1; r >= 0; --r) if (D.nodes[r].count > 0) D.order[0][L++] = r;
        s_live = L; s_nc = nIni;
    }
    __syncthreads();
    int ocur = 0, ecur = 0, live = s_live, tag = 0;
    bool done = (live == 0);

    while (!done) {
        const int prev = live;
        // ---- full sweep (src/ORBextractor.cc:683-753)
        int* ord = D.order[ocur]; int* ord2 = D.order[ocur ^ 1];
        for (int i = tid; i < live; i += T) D.flag[i] = D.nodes[ord[i]].leaf ? 1 : 0;
        __syncthreads();
        const int n_leaves = block_scan_i32(D.flag, live, s_i32);        // flag[i] = number of leaves stored before i
        __syncthreads();
        const int nparents = live - n_leaves;
        if (nparents == 0) break;
        if (s_nc + 4 * nparents > D.ncap) { if (tid == 0) atomicExch(A.error, 1); break; }
        ++tag;
        for (int i = tid; i < live; i += T) {
            const int id = ord[i];
            if (D.nodes[id].leaf) ord2[D.flag[i]] = id;                       // kept leaves, old relative order
            else { D.plist[nparents - 1 - (i - D.flag[i])] = id; D.nodes[id].tag = tag; }   // traversal = storage back -> front
        }
        __syncthreads();
        quadrant_scans(cand, n, D, cur, tag, s_u64);
        split_nodes(cand, n, D, cur, tag, nparents, &s_nc, s_i32, &s_nk, &s_ne, D.expand[ecur], ord2 + n_leaves);
        cur ^= 1; ocur ^= 1;
        const int nToExpand = s_ne;
        live = n_leaves + s_nk;
        __syncthreads();
        if (live >= N || live == prev) break;
        if (live + nToExpand * 3 <= N) continue;

        // ---- partial phase (src/ORBextractor.cc:760-838)
        int nexp = nToExpand;
        for (;;) {
            const int prev2 = live;
            ord = D.order[ocur];
            for (int i = tid; i < nexp; i += T) s_sort[i] = D.expand[ecur][i];
            __syncthreads();
            if (tid == 0) stdsort::sort(s_sort, nexp);
            __syncthreads();
            ++tag;
            for (int j = tid; j < nexp; j += T) { const int id = (int)(uint32_t)s_sort[nexp - 1 - j]; D.plist[j] = id; D.nodes[id].tag = tag; }   // largest first
            __syncthreads();
            quadrant_scans(cand, n, D, cur, tag, s_u64);
            // a split adds (#non-empty children - 1) nodes; the loop stops right after the split that reaches the quota
            for (int j = tid; j < nexp; j += T) {
                int c[4];
                child_counts(D, D.nodes[D.plist[j]], c);
                D.flag[j] = (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0) - 1;
            }
            __syncthreads();
            for (int j = tid; j < nexp; j += T) D.nkids[j] = D.flag[j];
            __syncthreads();
            block_scan_i32(D.flag, nexp, s_i32);                 // flag[j] = nodes gained before processing j
            if (tid == 0) s_flag = nexp;
            __syncthreads();
            for (int j = tid; j < nexp; j += T) if (live + D.flag[j] + D.nkids[j] >= N) atomicMin(&s_flag, j + 1);
            __syncthreads();
            const int nproc = s_flag;
            for (int j = nproc + tid; j < nexp; j += T) D.nodes[D.plist[j]].tag = 0;          // not reached: stay as they are
            __syncthreads();
            if (s_nc + 4 * nproc > D.ncap || live + 4 * nproc > D.ncap) { if (tid == 0) atomicExch(A.error, 1); done = true; break; }
            // children are pushed to the list front == appended to the storage, in processing order
            split_nodes(cand, n, D, cur, tag, nproc, &s_nc, s_i32, &s_nk, &s_ne, D.expand[ecur ^ 1], ord + live);
            cur ^= 1; ecur ^= 1;
            const int stored = live + s_nk;
            // drop the erased parents so the list stays dense
            int* ord2b = D.order[ocur ^ 1];
            for (int i = tid; i < stored; i += T) D.flag[i] = D.nodes[ord[i]].dead ? 0 : 1;
            __syncthreads();
            const int kept = block_scan_i32(D.flag, stored, s_i32);
            __syncthreads();
            for (int i = tid; i < stored; i += T) { const int id = ord[i]; if (!D.nodes[id].dead) ord2b[D.flag[i]] = id; }
            __syncthreads();
            ocur ^= 1;
            live = kept;
            nexp = s_ne;
            if (live >= N || live == prev2) { done = true; break; }
        }
        if (done) break;
    }
    __syncthreads();
    // ---- per surviving node, in list order (storage back -> front), the first maximum response (:842-862)
    {
        const int* ord = D.order[ocur];
        const int* perm = D.perm[cur];
        const int m = min(live, D.ncap);
        for (int i = tid; i < m; i += T) {
            const DNode nd = D.nodes[ord[live - 1 - i]];
            int best = perm[nd.begin];
            int br = unpack_s(cand[best]);
            for (int k = 1; k < nd.count; ++k) { const int c = perm[nd.begin + k]; const int r = unpack_s(cand[c]); if (r > br) { best = c; br = r; } }
            D.stage[i] = cand[best];
        }
        if (tid == 0) *out_count = m;
    }
}

// packs the per-level picks of a frame into its keypoint list (level-major) and publishes the level offsets
__global__ void __launch_bounds__(256)
k_pack_selected(DistArgs A, int* __restrict__ sel_level_off /* (nlevels+1) per frame */, int* __restrict__ n_kp_host /* mapped host, per frame */)
{
    const int frame = blockIdx.x, tid = threadIdx.x;
    __shared__ int s_off[PLVS_MAX_LEVELS + 1];
    if (tid == 0) {
        int k = 0;
        for (int l = 0; l < A.nlevels; ++l) { s_off[l] = k; k += A.sel_count[frame * A.nlevels + l]; }
        s_off[A.nlevels] = k;
        if (k > A.sel_cap) atomicExch(A.error, 2);
        for (int l = 0; l <= A.nlevels; ++l) sel_level_off[frame * (A.nlevels + 1) + l] = min(s_off[l], A.sel_cap);
        n_kp_host[frame] = min(k, A.sel_cap);
    }
    __syncthreads();
    for (int l = 0; l < A.nlevels; ++l) {
        const uint32_t* stage = A.scratch[frame * A.nlevels + l].stage;
        const int cnt = A.sel_count[frame * A.nlevels + l], base = s_off[l];
        for (int i = tid; i < cnt; i += 256) if (base + i < A.sel_cap) A.sel[(long long)frame * A.sel_cap + base + i] = stage[i];
    }
}

}  // namespace orb
}  // namespace plvs
