// Device code only: bag-of-words transform.  bow.cu includes it inside its anonymous namespace; tests/native/emu_kernels.cpp
// compiles the same text for the CPU (tests/native/cuda_emu.hpp).
//
// Frame::ComputeBoW (src/Frame.cc:1498-1505) -> DBoW2 TemplatedVocabulary::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1140-1271):
// every descriptor descends the vocabulary tree (k children per node, the first child with the smallest Hamming distance wins, :1246-1262), which
// yields its word, the word's weight and the node `levelsup` levels above the leaves; the features are then grouped by that node
// (DBoW2::FeatureVector: ascending node ids, ascending feature indices).  The tree lives in HBM with the descriptors of siblings stored
// contiguously (k x 32 bytes per inner node), so one level of one descent is one coalesced read by the first k lanes of a warp.
#pragma once

struct VocDev {
    const int* child_off;          // [n_nodes + 1]: the children of node i are the slots child_off[i] .. child_off[i+1] of the sibling arrays
    const int* child_id;           // node id per slot (file order = DBoW2's children vector)
    const uint8_t* child_desc;     // 32 bytes per slot
    const int* word_id;            // per node, -1 for inner nodes
    const double* weight;          // per node
    int L;
};

__device__ __forceinline__ int bow_hamming(const uint4 a0, const uint4 a1, const uint8_t* __restrict__ b)
{
    const uint4 b0 = *reinterpret_cast<const uint4*>(b), b1 = *reinterpret_cast<const uint4*>(b + 16);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// warp = feature
__global__ void __launch_bounds__(256)
k_bow_descend(VocDev V, const uint8_t* __restrict__ desc, int n, int levelsup, uint32_t* __restrict__ word, double* __restrict__ weight,
              uint32_t* __restrict__ node)
{
    const int f = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (f >= n) return;
    const uint8_t* d = desc + (size_t)f * 32;
    const uint4 a0 = *reinterpret_cast<const uint4*>(d), a1 = *reinterpret_cast<const uint4*>(d + 16);
    const int nid_level = V.L - levelsup;
    int final_id = 0, level = 0, nid = 0;             // nid stays 0 (root) when nid_level <= 0; a leaf above nid_level leaves it unset in the reference
    for (;;) {
        ++level;
        const int beg = V.child_off[final_id], cnt = V.child_off[final_id + 1] - beg;
        int best = INT_MAX, slot = INT_MAX;
        for (int c = lane; c < cnt; c += 32) {
            const int dist = bow_hamming(a0, a1, V.child_desc + (size_t)(beg + c) * 32);
            if (dist < best) { best = dist; slot = c; }            // ascending c per lane: the first minimum
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, o), os = __shfl_xor_sync(0xffffffffu, slot, o);
            if (ob < best || (ob == best && os < slot)) { best = ob; slot = os; }
        }
        final_id = V.child_id[beg + slot];
        if (level == nid_level) nid = final_id;
        if (V.child_off[final_id + 1] == V.child_off[final_id]) break;       // isLeaf(): no children
    }
    if (lane == 0) { word[f] = (uint32_t)V.word_id[final_id]; weight[f] = V.weight[final_id]; node[f] = (uint32_t)nid; }
}

// FeatureVector grouping, step 1: position of every kept feature (weight > 0, :1170) in (node, feature index) order, by counting
__global__ void __launch_bounds__(256)
k_bow_rank(const uint32_t* __restrict__ node, const double* __restrict__ weight, int n, int32_t* __restrict__ sorted_feat, uint32_t* __restrict__ sorted_node,
           int* __restrict__ n_kept)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n || !(weight[i] > 0)) return;
    const uint32_t mine = node[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
        if (!(weight[j] > 0)) continue;
        const uint32_t other = node[j];
        rank += (other < mine) || (other == mine && j < i);
    }
    sorted_feat[rank] = i; sorted_node[rank] = mine;
    atomicAdd(n_kept, 1);
}

// step 2 (one CTA): distinct nodes and their offsets; result[0] = number of nodes, fv_offsets[n_nodes] = number of kept features
__global__ void __launch_bounds__(1024)
k_bow_offsets(const uint32_t* __restrict__ sorted_node, const int* __restrict__ n_kept, uint32_t* __restrict__ fv_nodes, int32_t* __restrict__ fv_offsets,
              int* __restrict__ result)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = *n_kept;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int r = c0 + tid;
        const int f = (r < n && (r == 0 || sorted_node[r] != sorted_node[r - 1])) ? 1 : 0;
        int x = f;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += y; }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int pos = s_base + (wid ? s_warp[wid - 1] : 0) + x - f;
        if (f) { fv_nodes[pos] = sorted_node[r]; fv_offsets[pos] = r; }
        __syncthreads();
        if (tid == 1023) s_base += s_warp[31];
        __syncthreads();
    }
    if (tid == 0) { fv_offsets[s_base] = n; result[0] = s_base; }
}
