// Device code shared by the matcher kernels: feature grid (Frame::AssignFeaturesToGrid, src/Frame.cc:716-746), the cell window of
// Frame::GetFeaturesInArea (:1239-1261), 256-bit Hamming distance (src/ORBmatcher.cc:2198-2225), candidate packing, frame view.
// Device code only (no launches, no runtime calls): match.cu includes it inside its anonymous namespace, and tests/native/emu_kernels.cpp
// compiles the same text for the CPU (tests/native/cuda_emu.hpp) so the kernels can be exercised where there is no GPU.
#pragma once

constexpr int GRID_COLS = 64, GRID_ROWS = 48, GRID_CELLS = GRID_COLS * GRID_ROWS;
constexpr int TH_HIGH = 100, TH_LOW = 50, HISTO = 12;

struct GridParams { float min_x, min_y, max_x, max_y, inv_w, inv_h; };

// candidate packing: idx:16 | dist:9 | level:5
__device__ __forceinline__ uint32_t pack_cand(int idx, int dist, int level) { return (uint32_t)idx | ((uint32_t)dist << 16) | ((uint32_t)level << 25); }
__device__ __forceinline__ int cand_idx(uint32_t c) { return c & 0xffff; }
__device__ __forceinline__ int cand_dist(uint32_t c) { return (c >> 16) & 0x1ff; }
__device__ __forceinline__ int cand_level(uint32_t c) { return c >> 25; }

__device__ __forceinline__ int hamming256(const uint4 a0, const uint4 a1, const uint8_t* __restrict__ b)
{
    const uint4 b0 = *reinterpret_cast<const uint4*>(b), b1 = *reinterpret_cast<const uint4*>(b + 16);
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
           __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---------------------------------------------------------------------------------------------
// Frame::AssignFeaturesToGrid (src/Frame.cc:716-746): 64x48 cells, cell = round((p-min)*inv),
// indices appended in keypoint order.  One CTA: histogram, scan, unordered scatter, per-cell insertion sort
// (cells hold a handful of keypoints), so every cell list is ascending == the reference's push_back order.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void build_grid_cta(const plvs_keypoint* __restrict__ keys, int n, GridParams gp, int* __restrict__ cell_start /*GRID_CELLS+1*/,
                                               int* __restrict__ sorted /*n*/, int* __restrict__ kp_cell /*n*/)
{
    __shared__ int s_cnt[GRID_CELLS + 1];
    __shared__ int s_part[32];
    const int tid = threadIdx.x;
    for (int i = tid; i <= GRID_CELLS; i += 1024) s_cnt[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        const int px = (int)roundf((keys[i].x - gp.min_x) * gp.inv_w);
        const int py = (int)roundf((keys[i].y - gp.min_y) * gp.inv_h);
        int c = -1;
        if (px >= 0 && px < GRID_COLS && py >= 0 && py < GRID_ROWS) { c = px * GRID_ROWS + py; atomicAdd(&s_cnt[c], 1); }
        kp_cell[i] = c;
    }
    __syncthreads();
    // exclusive scan of 3072 counts: 3 per thread
    int v[3], sum = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) { v[k] = s_cnt[tid * 3 + k]; sum += v[k]; }
    int x = sum;
    const int lane = tid & 31, wid = tid >> 5;
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
    int base = (wid ? s_part[wid - 1] : 0) + x - sum;
#pragma unroll
    for (int k = 0; k < 3; ++k) { s_cnt[tid * 3 + k] = base; base += v[k]; }
    if (tid == 1023) s_cnt[GRID_CELLS] = base;
    __syncthreads();
    for (int i = tid; i <= GRID_CELLS; i += 1024) cell_start[i] = s_cnt[i];
    __syncthreads();
    // scatter in arbitrary order (s_cnt doubles as the per-cell cursor), then every cell list -- they hold
    // 0-3 entries in practice -- is put back into ascending index order (== push_back order) by one thread
    for (int i = tid; i < n; i += 1024) { const int c = kp_cell[i]; if (c >= 0) sorted[atomicAdd(&s_cnt[c], 1)] = i; }
    __syncthreads();
    for (int c = tid; c < GRID_CELLS; c += 1024) {
        const int b0 = cell_start[c], b1 = cell_start[c + 1];
        for (int i = b0 + 1; i < b1; ++i) {
            const int v = sorted[i];
            int j = i - 1;
            while (j >= b0 && sorted[j] > v) { sorted[j + 1] = sorted[j]; --j; }
            sorted[j + 1] = v;
        }
    }
}

__global__ void __launch_bounds__(1024)
k_build_grid(const plvs_keypoint* __restrict__ keys, int n, GridParams gp, int* __restrict__ cell_start, int* __restrict__ sorted, int* __restrict__ kp_cell)
{
    build_grid_cta(keys, n, gp, cell_start, sorted, kp_cell);
}

// The same at frame construction (the reference calls AssignFeaturesToGrid at the end of the Frame constructor, src/Frame.cc:598): one CTA per
// frame of an extractor batch, over the keypoints the extractor left in HBM.  count_of[b * count_stride] = number of keypoints of frame b.
__global__ void __launch_bounds__(1024)
k_build_grid_batch(const plvs_keypoint* __restrict__ keys, int key_stride, const int* __restrict__ count_of, int count_stride, GridParams gp,
                   int* __restrict__ cell_start /*B x (GRID_CELLS+1)*/, int* __restrict__ sorted /*B x key_stride*/, int* __restrict__ kp_cell /*B x key_stride*/)
{
    const int b = blockIdx.x;
    build_grid_cta(keys + (size_t)b * key_stride, min(count_of[(size_t)b * count_stride], key_stride), gp, cell_start + (size_t)b * (GRID_CELLS + 1),
                   sorted + (size_t)b * key_stride, kp_cell + (size_t)b * key_stride);
}

// Frame::GetFeaturesInArea cell window (src/Frame.cc:1239-1261); returns false if empty
__device__ __forceinline__ bool cell_window(const GridParams& gp, float x, float y, float r, int& c0, int& c1, int& r0, int& r1)
{
    c0 = max(0, (int)floorf((x - gp.min_x - r) * gp.inv_w));
    if (c0 >= GRID_COLS) return false;
    c1 = min(GRID_COLS - 1, (int)ceilf((x - gp.min_x + r) * gp.inv_w));
    if (c1 < 0) return false;
    r0 = max(0, (int)floorf((y - gp.min_y - r) * gp.inv_h));
    if (r0 >= GRID_ROWS) return false;
    r1 = min(GRID_ROWS - 1, (int)ceilf((y - gp.min_y + r) * gp.inv_h));
    if (r1 < 0) return false;
    return true;
}

struct ViewDev {
    const plvs_keypoint* keys; const uint8_t* desc; const float* uright; int n;
    GridParams gp; float bf; float scale[PLVS_MAX_LEVELS]; float sigma2[PLVS_MAX_LEVELS];
};
