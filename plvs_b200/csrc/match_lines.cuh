// Device code of the line-descriptor k-NN: what LineMatcher::ComputeDescriptorMatches asks of
// cv::line_descriptor_c::BinaryDescriptorMatcher::knnMatch(query, train, matches, 2, mask, true) (src/LineMatcher.cc:2567-2615 ->
// Thirdparty/line_descriptor/src/binary_descriptor_matcher_custom.cpp:258-337).  The library answers with multi-index hashing over the 32 bytes
// of the 256-bit LBD descriptor; between equidistant neighbours the order is the order in which Mihasher::query (:633-788) discovers them, and the
// ratio test of the caller reads it.  That order is a per-pair key: (Hamming distance, smallest byte distance s*, first byte k* that attains it,
// position of the byte q[k*] ^ t[k*] in the library's enumeration of the 8-bit strings with s* ones, train index) -- so a brute-force scan that
// keeps the two smallest keys returns what the hash tables return.  Device code only; match.cu includes it.
#pragma once

// per-byte popcounts of a 32-bit word, packed one count per byte
__device__ __forceinline__ uint32_t byte_popcounts(uint32_t w)
{
    w = w - ((w >> 1) & 0x55555555u);
    w = (w & 0x33333333u) + ((w >> 2) & 0x33333333u);
    return (w + (w >> 4)) & 0x0f0f0f0fu;
}

// one warp per query descriptor, lanes over the train descriptors
__global__ void __launch_bounds__(256)
k_line_knn2(const uint8_t* __restrict__ query, int nq, const uint8_t* __restrict__ train, int nt, const int* __restrict__ enum_rank /*256*/,
            unsigned long long* __restrict__ best2 /*2 per query: the two smallest keys*/)
{
    __shared__ int s_rank[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_rank[i] = enum_rank[i];
    __syncthreads();
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (q >= nq) return;
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = reinterpret_cast<const uint32_t*>(query + (size_t)q * 32)[i];
    unsigned long long k1 = ~0ull, k2 = ~0ull;
    for (int t = lane; t < nt; t += 32) {
        const uint32_t* b = reinterpret_cast<const uint32_t*>(train + (size_t)t * 32);
        int ham = 0, smin = 9, kfirst = 0; uint32_t xbyte = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t x = a[i] ^ b[i];
            const uint32_t pc = byte_popcounts(x);
            ham += __popc(x);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = (int)((pc >> (8 * j)) & 0xffu);
                if (c < smin) { smin = c; kfirst = 4 * i + j; xbyte = (x >> (8 * j)) & 0xffu; }     // strict: the FIRST byte that attains the minimum
            }
        }
        const unsigned long long key = ((((((unsigned long long)ham * 16ull + (unsigned long long)smin) * 32ull + (unsigned long long)kfirst) * 128ull) +
                                         (unsigned long long)s_rank[xbyte]) << 20) + (unsigned long long)t;
        if (key < k1) { k2 = k1; k1 = key; } else if (key < k2) k2 = key;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const unsigned long long o1 = __shfl_xor_sync(0xffffffffu, k1, o), o2 = __shfl_xor_sync(0xffffffffu, k2, o);
        const unsigned long long lo = k1 < o1 ? k1 : o1, hi = k1 < o1 ? o1 : k1;
        const unsigned long long m2 = k2 < o2 ? k2 : o2;
        k1 = lo; k2 = hi < m2 ? hi : m2;
    }
    if (lane == 0) { best2[2 * (size_t)q] = k1; best2[2 * (size_t)q + 1] = k2; }
}

// compact result + ratio test (src/LineMatcher.cc:2590-2612): one CTA; rows of the queries whose mask entry is non-zero, in query order
__global__ void __launch_bounds__(1024)
k_line_rows(const unsigned long long* __restrict__ best2, int nq, const uint8_t* __restrict__ mask /*or NULL*/, float nn_ratio,
            int32_t* __restrict__ query_idx, int32_t* __restrict__ train_idx, float* __restrict__ dist, uint8_t* __restrict__ valid, int* __restrict__ result /*rows, valid*/)
{
    __shared__ int s_part[32], s_valid, s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) { s_valid = 0; s_base = 0; }
    __syncthreads();
    for (int base = 0; base < nq; base += 1024) {
        const int q = base + tid;
        const int keep = q < nq && (!mask || mask[q]) ? 1 : 0;
        int x = keep;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_part[wid] = x;
        __syncthreads();
        if (wid == 0) {
            int p = s_part[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, p, o); if (lane >= o) p += y; }
            s_part[lane] = p;
        }
        __syncthreads();
        const int row = s_base + (wid ? s_part[wid - 1] : 0) + x - keep;
        if (keep) {
            const unsigned long long k1 = best2[2 * (size_t)q], k2 = best2[2 * (size_t)q + 1];
            const float d1 = (float)(int)(k1 >> 36), d2 = (float)(int)(k2 >> 36);           // key = ((ham*16 + s)*32 + k)*128 + rank) << 20 | t: ham sits above bit 36
            query_idx[row] = q;
            train_idx[2 * row] = (int)(k1 & 0xfffffu); train_idx[2 * row + 1] = (int)(k2 & 0xfffffu);
            dist[2 * row] = d1; dist[2 * row + 1] = d2;
            const bool ok = d1 < nn_ratio * d2;
            valid[row] = ok ? 1 : 0;
            if (ok) atomicAdd(&s_valid, 1);
        }
        __syncthreads();
        if (tid == 0) s_base += s_part[31];
        __syncthreads();
    }
    if (tid == 0) { result[0] = s_base; result[1] = s_valid; }
}
