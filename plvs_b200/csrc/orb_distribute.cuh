// Device keypoint distributor: PLVS2::ORBextractor::DistributeOctTree (src/ORBextractor.cc:611-865,
// DivideNode :536-592, compareNodes :594-609) as ONE CTA per (pyramid level, frame).
//
// The reference is a sequential std::list algorithm whose iteration order is observable (it is the order of the
// returned keypoints).  What it computes, restated so that it parallelises:
//   * A node = (UL.x, UL.y, UR.x, BR.y) + the set of candidates inside it.  The reference keeps every node's keys in a vector whose order is
//     preserved by every split (stable), so inside a node the keys are always in candidate-index order and "first maximum response wins"
//     (:842-862) is simply max over (response, lowest index).  Nothing else reads that order: the candidates need not be moved at all -- each
//     one only carries the id of the node it is in (node_of[p]); a split counts its candidates per quadrant and relabels them.
//   * Full sweeps: every node with >1 point splits; children are pushed to the list FRONT, the parent is erased.
//     With the list stored back-to-front a sweep is  new_list = kept-leaves (old order) ++ children (creation order).
//     All splits of a sweep are two passes over the candidates: count per (parent, quadrant) -- in shared memory while the sweep has few
//     parents, where all candidates hit a handful of counters -- and relabel; child slots come from prefix scans over the node list.
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
    int slot;          // index in plist while tagged: its quadrant counters are qcnt[4 * slot .. 4 * slot + 3]
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

constexpr int kQcntShared = 1024;          // parents whose quadrant counters live in shared memory (4 ints each)

__device__ __forceinline__ int* qcnt_of(const DistLevel& D) { return reinterpret_cast<int*>(D.scan_a); }     // 4 ints per parent: 2 (n + 1) ints are there, parents hold >= 2 candidates

__device__ __forceinline__ void child_counts(const DistLevel& D, int slot, int c[4])
{
    const int* q = qcnt_of(D) + 4 * slot;
    c[0] = q[0]; c[1] = q[1]; c[2] = q[2]; c[3] = q[3];
}

// candidates per quadrant of every node that carries `tag` (slots 0..nparents-1) -> qcnt
__device__ inline void quadrant_counts(const uint32_t* __restrict__ cand, int n, const DistLevel& D, int tag, int nparents, int* s_q /*4 * kQcntShared*/)
{
    const int tid = threadIdx.x, T = blockDim.x;
    int* qc = qcnt_of(D);
    const bool in_smem = nparents <= kQcntShared;
    int* cnt = in_smem ? s_q : qc;
    for (int i = tid; i < 4 * nparents; i += T) cnt[i] = 0;
    __syncthreads();
    const int* __restrict__ nof = D.node_of[0];
    const DNode* __restrict__ nodes = D.nodes;
    // With few parents thousands of candidates hit a handful of counters: the lanes of a warp that hit the same one are found with MATCH.ANY and
    // their leader adds the group's size (same-address shared-memory atomics serialise)
    const bool aggregate = nparents <= 64;
    for (int pb = 0; pb < n; pb += 4 * T) {              // four candidates per thread and step: their dependent loads are in flight together
        const int p0 = pb + tid;
        int k[4]; uint32_t c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int p = p0 + u * T; k[u] = p < n ? nof[p] : -1; c[u] = p < n ? cand[p] : 0u; }
        uint2 bd[4]; int tg[4], sl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            bd[u] = make_uint2(0u, 0u); tg[u] = tag - 1; sl[u] = 0;
            if (k[u] >= 0) { const uint32_t* nw32 = reinterpret_cast<const uint32_t*>(&nodes[k[u]]); bd[u] = make_uint2(nw32[0], nw32[1]); tg[u] = nodes[k[u]].tag; sl[u] = nodes[k[u]].slot; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int key = tg[u] == tag ? 4 * sl[u] + quadrant_of_packed(bd[u], unpack_x(c[u]) - kRoiMargin, unpack_y(c[u]) - kRoiMargin) : -1;
            if (aggregate) {
                const uint32_t grp = __match_any_sync(0xffffffffu, key);
                if (key >= 0 && (__ffs(grp) - 1) == (tid & 31)) atomicAdd(&cnt[key], __popc(grp));
            } else if (key >= 0) atomicAdd(&cnt[key], 1);
        }
    }
    __syncthreads();
    if (in_smem) { for (int i = tid; i < 4 * nparents; i += T) qc[i] = s_q[i]; __syncthreads(); }
}

// Splits plist[0..nparents) (all carrying `tag`, quadrant_counts already done) at once.  Children are created in the
// order of plist x quadrant (n1,n2,n3,n4), ids from *node_count; their ids go to child_order[], the ones with more
// than one point to expand_out[] (same order).  The candidates of the split nodes are relabelled with their child.
__device__ inline void split_nodes(const uint32_t* __restrict__ cand, int n, DistLevel& D, int tag, int nparents, int* node_count,
                                   int* s_i32, int* total_kids_out, int* total_exp_out, unsigned long long* expand_out, int* child_order)
{
    const int tid = threadIdx.x, T = blockDim.x;
    for (int j = tid; j < nparents; j += T) {
        int c[4];
        child_counts(D, j, c);
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
        child_counts(D, j, c);
        const int mx = nd.ulx + ((nd.urx - nd.ulx + 1) >> 1), my = nd.uly + ((nd.bry - nd.uly + 1) >> 1);
        int k = D.nkids[j], e = D.nexp[j];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (c[q] == 0) { nd.kid[q] = -1; continue; }
            DNode ch;
            ch.ulx = (short)((q & 1) ? mx : nd.ulx); ch.urx = (short)((q & 1) ? nd.urx : mx);
            ch.uly = (short)((q & 2) ? my : nd.uly); ch.bry = (short)((q & 2) ? nd.bry : my);
            ch.begin = 0; ch.count = c[q]; ch.leaf = c[q] == 1; ch.dead = 0; ch.tag = 0; ch.slot = 0;
            ch.kid[0] = ch.kid[1] = ch.kid[2] = ch.kid[3] = -1;
            const int id = base_id + k;
            D.nodes[id] = ch;
            nd.kid[q] = id;
            child_order[k] = id;
            if (c[q] > 1) { expand_out[e] = ((unsigned long long)(((unsigned)min(c[q], (1 << 20) - 1) << 12) | (unsigned)ch.ulx) << 32) | (unsigned)id; ++e; }
            ++k;
        }
        nd.dead = 1;
        D.nodes[pid] = nd;
    }
    __syncthreads();
    int* __restrict__ nof = D.node_of[0];
    const DNode* __restrict__ nodes = D.nodes;
    for (int p0 = tid; p0 < n; p0 += 4 * T) {
        int k[4]; uint32_t c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int p = p0 + u * T; k[u] = p < n ? nof[p] : -1; c[u] = p < n ? cand[p] : 0u; }
        uint2 bd[4]; int tg[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            bd[u] = make_uint2(0u, 0u); tg[u] = tag - 1;
            if (k[u] >= 0) { const uint32_t* nw32 = reinterpret_cast<const uint32_t*>(&nodes[k[u]]); bd[u] = make_uint2(nw32[0], nw32[1]); tg[u] = nodes[k[u]].tag; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (tg[u] == tag) nof[p0 + u * T] = nodes[k[u]].kid[quadrant_of_packed(bd[u], unpack_x(c[u]) - kRoiMargin, unpack_y(c[u]) - kRoiMargin)];
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
    int sort_elems;                // dynamic shared memory: the std::sort area first (sort_elems 64-bit elements + the scratch of stdsort::sort_cta) ...
    int sort_bytes;                // ... sort_bytes in all
    int arena_bytes;               // ... then (single-frame calls) room for the node-level state; 0 = node-level state in global memory
    int arena_nodes;               // test hook: > 0 caps the nodes the arena is carved for (forces the start-over path)
};

__global__ void __launch_bounds__(kDistThreads)
k_distribute(DistArgs A)
{
    PLVS_DYN_SMEM(unsigned long long, s_sort);           // max quota + 8 elements for the std::sort emulation, then its scratch, then the arena
    __shared__ int s_i32[32];
    __shared__ int s_nc, s_nk, s_ne, s_live, s_flag;
    __shared__ int s_q[4 * kQcntShared];
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
    // Latency mode (single-frame calls; the host passes an arena): the node-level state -- node pool, the list, the per-parent scratch, the
    // expandable-node lists -- lives in shared memory instead of L2, which is where the ~15 dependent phases of a sweep spent their time.  The
    // pointers of D are simply redirected; the code below is the same.  The arena holds 4 x quota + 128 nodes, plenty in practice (a sweep
    // stops at the quota); should a level ever need more, the kernel starts that level over with the state in global memory.
    bool in_smem = false;
    {
        const int scap = A.arena_nodes > 0 ? min(D.ncap, A.arena_nodes) : min(D.ncap, 4 * N + 128);
        if (A.arena_bytes > 0 && (long long)scap * (long long)(sizeof(DNode) + 2 * 8 + 6 * 4) <= (long long)A.arena_bytes && nIni <= scap) {
            char* base = reinterpret_cast<char*>(s_sort) + A.sort_bytes;
            D.expand[0] = reinterpret_cast<unsigned long long*>(base); base += 8 * (size_t)scap;
            D.expand[1] = reinterpret_cast<unsigned long long*>(base); base += 8 * (size_t)scap;
            D.nodes = reinterpret_cast<DNode*>(base); base += sizeof(DNode) * (size_t)scap;
            D.order[0] = reinterpret_cast<int*>(base); base += 4 * (size_t)scap;
            D.order[1] = reinterpret_cast<int*>(base); base += 4 * (size_t)scap;
            D.plist = reinterpret_cast<int*>(base); base += 4 * (size_t)scap;
            D.nkids = reinterpret_cast<int*>(base); base += 4 * (size_t)scap;
            D.nexp = reinterpret_cast<int*>(base); base += 4 * (size_t)scap;
            D.flag = reinterpret_cast<int*>(base);
            D.ncap = scap;
            in_smem = true;
        }
    }
    bool retry;
restart:
    retry = false;

    // ---- roots (src/ORBextractor.cc:626-664): every candidate goes to the root its x falls in (nIni is 1-3)
    for (int r = tid; r < nIni; r += T) {
        DNode nd;
        nd.ulx = (short)(int)(hX * (float)r); nd.uly = 0; nd.urx = (short)(int)(hX * (float)(r + 1)); nd.bry = (short)(maxY - minY);
        nd.begin = 0; nd.count = 0; nd.leaf = 0; nd.dead = 0; nd.tag = 0; nd.slot = 0; nd.kid[0] = nd.kid[1] = nd.kid[2] = nd.kid[3] = -1;
        D.nodes[r] = nd;
    }
    for (int i = tid; i < nIni; i += T) s_q[i] = 0;
    __syncthreads();
    for (int pb = 0; pb < n; pb += T) {
        const int p = pb + tid;
        int r = -1;
        if (p < n) {
            r = (int)((float)(unpack_x(cand[p]) - kRoiMargin) / hX);
            if (!(r >= 0 && r < nIni)) r = -1;
            D.node_of[0][p] = r;
        }
        const uint32_t grp = __match_any_sync(0xffffffffu, r);          // one to three roots: one atomic per warp and root
        if (r >= 0 && (__ffs(grp) - 1) == (tid & 31)) atomicAdd(&s_q[r], __popc(grp));
    }
    __syncthreads();
    for (int r = tid; r < nIni; r += T) { D.nodes[r].count = s_q[r]; D.nodes[r].leaf = s_q[r] == 1; }
    __syncthreads();
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
        if (s_nc + 4 * nparents > D.ncap) { if (in_smem) retry = true; else if (tid == 0) atomicExch(A.error, 1); break; }
        ++tag;
        for (int i = tid; i < live; i += T) {
            const int id = ord[i];
            if (D.nodes[id].leaf) ord2[D.flag[i]] = id;                       // kept leaves, old relative order
            else { const int j = nparents - 1 - (i - D.flag[i]); D.plist[j] = id; D.nodes[id].tag = tag; D.nodes[id].slot = j; }   // traversal = storage back -> front
        }
        __syncthreads();
        quadrant_counts(cand, n, D, tag, nparents, s_q);
        split_nodes(cand, n, D, tag, nparents, &s_nc, s_i32, &s_nk, &s_ne, D.expand[ecur], ord2 + n_leaves);
        ocur ^= 1;
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
            stdsort::sort_cta(s_sort, nexp, reinterpret_cast<int*>(s_sort + A.sort_elems));      // std::sort, range by range (stdsort_emul.cuh)
            ++tag;
            for (int j = tid; j < nexp; j += T) { const int id = (int)(uint32_t)s_sort[nexp - 1 - j]; D.plist[j] = id; D.nodes[id].tag = tag; D.nodes[id].slot = j; }   // largest first
            __syncthreads();
            quadrant_counts(cand, n, D, tag, nexp, s_q);
            // a split adds (#non-empty children - 1) nodes; the loop stops right after the split that reaches the quota
            for (int j = tid; j < nexp; j += T) {
                int c[4];
                child_counts(D, j, c);
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
            if (s_nc + 4 * nproc > D.ncap || live + 4 * nproc > D.ncap) { if (in_smem) retry = true; else if (tid == 0) atomicExch(A.error, 1); done = true; break; }
            // children are pushed to the list front == appended to the storage, in processing order
            split_nodes(cand, n, D, tag, nproc, &s_nc, s_i32, &s_nk, &s_ne, D.expand[ecur ^ 1], ord + live);
            ecur ^= 1;
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
    if (retry) {            // uniform: the arena was too small for this level -- once more with the node-level state in global memory
        D = A.scratch[frame * A.nlevels + level];
        in_smem = false;
        goto restart;
    }
    // ---- per surviving node, in list order (storage back -> front), the first maximum response (:842-862): inside a node the reference's keys
    // are in candidate-index order, so "first maximum" is the largest (response, lowest index) -- one atomicMax per candidate
    {
        const int* ord = D.order[ocur];
        unsigned long long* best = A.scratch[frame * A.nlevels + level].expand[0];     // free by now; indexed by node id.  Always the GLOBAL array: a 64-bit
                                                                                       // atomicMax is one fire-and-forget RED in L2, but a CAS loop in shared memory
        const int m = min(live, D.ncap);
        const int nc = s_nc;
        for (int i = tid; i < nc; i += T) best[i] = 0ull;
        __syncthreads();
        const int* __restrict__ nof = D.node_of[0];
        for (int p = tid; p < n; p += T) {
            const int k = nof[p];
            if (k >= 0) atomicMax(&best[k], ((unsigned long long)(uint32_t)unpack_s(cand[p]) << 32) | (unsigned long long)(0xffffffffu - (uint32_t)p));
        }
        __syncthreads();
        for (int i = tid; i < m; i += T) D.stage[i] = cand[0xffffffffu - (uint32_t)best[ord[live - 1 - i]]];
        if (tid == 0) *out_count = m;
    }
}

// packs the per-level picks of a frame into its keypoint list (level-major) and publishes the level offsets
__global__ void __launch_bounds__(256)
k_pack_selected(DistArgs A, int* __restrict__ sel_level_off /* (nlevels+1) per frame */, int* __restrict__ n_kp_host /* mapped host, per frame */)
{
    const int frame = blockIdx.x, tid = threadIdx.x;
    __shared__ int s_off[PLVS_MAX_LEVELS + 1], s_cnt[PLVS_MAX_LEVELS];
    __shared__ const uint32_t* s_stage[PLVS_MAX_LEVELS];
    // the per-level records are fetched by one thread each (one trip to L2 for all levels), then a short serial prefix
    if (tid < A.nlevels) { s_cnt[tid] = A.sel_count[frame * A.nlevels + tid]; s_stage[tid] = A.scratch[frame * A.nlevels + tid].stage; }
    __syncthreads();
    if (tid == 0) {
        int k = 0;
        for (int l = 0; l < A.nlevels; ++l) { s_off[l] = k; k += s_cnt[l]; }
        s_off[A.nlevels] = k;
        if (k > A.sel_cap) atomicExch(A.error, 2);
        n_kp_host[frame] = min(k, A.sel_cap);
    }
    __syncthreads();
    if (tid <= A.nlevels) sel_level_off[frame * (A.nlevels + 1) + tid] = min(s_off[tid], A.sel_cap);
    // one warp per level (round-robin when there are more levels than warps): the levels are copied side by side
    for (int l = tid >> 5; l < A.nlevels; l += 8) {
        const uint32_t* stage = s_stage[l];
        const int cnt = s_cnt[l], base = s_off[l];
        for (int i = tid & 31; i < cnt; i += 32) if (base + i < A.sel_cap) A.sel[(long long)frame * A.sel_cap + base + i] = stage[i];
    }
}

}  // namespace orb
}  // namespace plvs
