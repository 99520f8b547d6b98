// Device kernels of the ORB extractor (sm_100a).  Each kernel cites the reference row of
// SURVEY.md §8a it replaces; the integer recipes for the OpenCV primitives are those of §8c'.
#pragma once
#include "common.cuh"
#include "libm_sincosf.cuh"

namespace plvs {
namespace orb {

constexpr int kEdge = 19;           // EDGE_THRESHOLD (src/ORBextractor.cc:106)
constexpr int kHalfPatch = 15;      // HALF_PATCH_SIZE
constexpr int kRoiMargin = kEdge - 3;
constexpr int kMaxCell = 80;        // cell image side is < 70 + 6 (W=35 => wCell < 70)
constexpr int kCellPitch = 96;      // shared-memory pitch of the cell image: kMaxCell + 15 rounded up to 16 (the TMA box starts at a 16-byte aligned x)

// ---- geometry tables (built once per image size on the host, resident in HBM) ---------------
struct LevelGeom {
    int w, h, pitch;        // level image, pitch in bytes (multiple of 128)
    long long off;          // byte offset of the level inside one frame's pyramid buffer
    int cell_begin, cell_count;
    int slot_begin, slot_count;   // candidate slots of this level inside one frame's slot array
    int tap_x_off, tap_y_off;     // offsets into the bilinear tap table (dst x / dst y of this level)
    float scale;                  // mvScaleFactor[level]
    int patch_size;               // (int)(PATCH_SIZE * scale)  (src/ORBextractor.cc:1037)
};

struct CellDesc {           // one FAST cell = one cv::FAST call of the reference (src/ORBextractor.cc:930-971)
    short level;
    short x0, y0, x1, y1;   // cell image rectangle in level coordinates, [x0,x1) x [y0,y1)
    int slot_off;           // first candidate slot (inside the frame's slot array)
    int cap;                // slots reserved: ceil((w-6)/2)*ceil((h-6)/2) bounds the strict 3x3 maxima
};

struct TileDesc { short level, tx, ty, pad; };

struct BilinearTap { unsigned short i0, i1; short c0, c1; };   // 8 bytes

// candidate / keypoint packing: x:12 | y:12 | score:8
__host__ __device__ inline uint32_t pack_xys(int x, int y, int s) { return (uint32_t)x | ((uint32_t)y << 12) | ((uint32_t)s << 24); }
__host__ __device__ inline int unpack_x(uint32_t v) { return v & 0xfff; }
__host__ __device__ inline int unpack_y(uint32_t v) { return (v >> 12) & 0xfff; }
__host__ __device__ inline int unpack_s(uint32_t v) { return v >> 24; }

// ---------------------------------------------------------------------------------------------
// a2  ComputePyramid: one level from the previous one (src/ORBextractor.cc:1494 -> cv::resize
//     INTER_LINEAR, 11-bit fixed point).  Thread = 4 horizontally adjacent output pixels.
//     HBM traffic: reads w_{l-1}*h_{l-1} (through L1/L2), writes w_l*h_l once, coalesced 4 B/thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_resize_level(uint8_t* __restrict__ pyr, long long frame_stride, LevelGeom src, LevelGeom dst,
               const BilinearTap* __restrict__ taps)
{
    const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x4 >= dst.w || y >= dst.h) return;
    uint8_t* frame = pyr + (long long)blockIdx.z * frame_stride;
    const uint8_t* s = frame + src.off;
    uint8_t* d = frame + dst.off + (long long)y * dst.pitch;
    const BilinearTap ty = taps[dst.tap_y_off + y];
    const uint8_t* r0 = s + (long long)ty.i0 * src.pitch;
    const uint8_t* r1 = s + (long long)ty.i1 * src.pitch;
    uint32_t out = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int x = x4 + k;
        if (x < dst.w) {
            const BilinearTap tx = taps[dst.tap_x_off + x];
            const int h0 = r0[tx.i0] * tx.c0 + r0[tx.i1] * tx.c1;
            const int h1 = r1[tx.i0] * tx.c0 + r1[tx.i1] * tx.c1;
            const int v = (((ty.c0 * (h0 >> 4)) >> 16) + ((ty.c1 * (h1 >> 4)) >> 16) + 2) >> 2;
            out |= (uint32_t)(v & 0xff) << (8 * k);
        }
    }
    if (x4 + 3 < dst.w) *reinterpret_cast<uint32_t*>(d + x4) = out;       // pitch%128==0 => aligned
    else for (int k = 0; x4 + k < dst.w; ++k) d[x4 + k] = (uint8_t)(out >> (8 * k));
}

// ---------------------------------------------------------------------------------------------
// Step before extraction: cv::cvtColor(.., COLOR_{BGR,RGB,BGRA,RGBA}2GRAY) on 8-bit images
// (src/Tracking.cc:1797-1810).  OpenCV 4's fixed point: (R*9798 + G*19235 + B*3735 + 2^14) >> 15 -- checked against the real cv2
// of this image for all 2^24 colours (tests/test_oracle_orb.py).  Writes level 0 of the pyramid directly.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_color_to_gray(const uint8_t* __restrict__ img, long long frame_stride_in, int stride, int nch, int is_rgb,
                uint8_t* __restrict__ pyr, long long frame_stride, LevelGeom g0)
{
    const int x = blockIdx.x * 256 + threadIdx.x, y = blockIdx.y;
    if (x >= g0.w) return;
    const uint8_t* p = img + (long long)blockIdx.z * frame_stride_in + (long long)y * stride + (long long)x * nch;
    const uint32_t c0 = p[0], c1 = p[1], c2 = p[2];
    const uint32_t b = is_rgb ? c2 : c0, r = is_rgb ? c0 : c2;
    pyr[(long long)blockIdx.z * frame_stride + g0.off + (long long)y * g0.pitch + x] = (uint8_t)((r * 9798u + c1 * 19235u + b * 3735u + (1u << 14)) >> 15);
}

// Step after extraction (§8f rank 2): Frame::ComputeStereoFromRGBD (src/Frame.cc:2251-2279) on the device-resident keypoints:
// d = imDepth.at<float>(v, u) with the float coordinates truncated, mvDepth = d and mvuRight = xUn - bf/d where d > 0, else -1.
__global__ void __launch_bounds__(256)
k_stereo_from_rgbd(const plvs_keypoint* __restrict__ keys, int n, const float* __restrict__ keys_un_x, const float* __restrict__ depth, int w, int h,
                   int stride_f, float bf, float* __restrict__ uright, float* __restrict__ kdepth)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const plvs_keypoint kp = keys[i];
    const int u = (int)kp.x, v = (int)kp.y;
    float ur = -1.f, dz = -1.f;
    if (u >= 0 && u < w && v >= 0 && v < h) {
        const float d = depth[(size_t)v * stride_f + u];
        if (d > 0) { dz = d; ur = (keys_un_x ? keys_un_x[i] : kp.x) - bf / d; }
    }
    uright[i] = ur; kdepth[i] = dz;
}

#include "orb_undistort.cuh"

// ---------------------------------------------------------------------------------------------
// a7  7x7 Gaussian, sigma 2, BORDER_REFLECT_101 of the level itself (src/ORBextractor.cc:1343-1344;
//     OpenCV u8 fixed point: taps [18,34,48,56,48,34,18]/256, 8.8 row pass, 16.16 column pass,
//     one rounding).  Tile 128x16 outputs, halo staged in shared memory.
// ---------------------------------------------------------------------------------------------
constexpr int kBlurTW = 128, kBlurTH = 16;

__device__ __forceinline__ int reflect101(int p, int n)
{
    if (p < 0) p = -p;
    if (p >= n) p = 2 * n - 2 - p;
    return p;          // valid for n >= 4 and |overshoot| <= 3 (levels are always > 7 px)
}

// Staging: the (16+6) x (128+6) input window of a tile.  Tiles whose window lies inside the level in x (no reflection left or
// right) fetch it with one 1-D bulk asynchronous copy (TMA) per row -- 160 bytes from the 16-byte aligned column x0-16; rows are
// reflected by choosing the source row -- completed through an mbarrier; tiles at the left / right edge of a level, where
// BORDER_REFLECT_101 mirrors columns, use byte loads.  Column c of the window (c = 0 <-> x0-3) sits at s_in[r][c + 13].
constexpr int kBlurSW = 160, kBlurSO = 13;

__global__ void __launch_bounds__(256)
k_blur(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur, long long frame_stride,
       const LevelGeom* __restrict__ levels, const TileDesc* __restrict__ tiles)
{
    __shared__ __align__(128) uint8_t s_in[kBlurTH + 6][kBlurSW];
    __shared__ uint16_t s_row[kBlurTH + 6][kBlurTW];
    __shared__ __align__(8) unsigned long long s_bar;
    const TileDesc t = tiles[blockIdx.x];
    const LevelGeom g = levels[t.level];
    const uint8_t* src = pyr + (long long)blockIdx.y * frame_stride + g.off;
    uint8_t* dst = blur + (long long)blockIdx.y * frame_stride + g.off;
    const int x0 = t.tx * kBlurTW, y0 = t.ty * kBlurTH;
    const int tid = threadIdx.x;
    const bool interior = x0 >= 16 && x0 + kBlurTW + 3 <= g.w;      // then [x0-16, x0+144) is inside the row pitch (a multiple of 128)
    if (interior) {
        const uint32_t bar = tma_smem_u32(&s_bar);
        if (tid == 0) tma_mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            tma_mbar_expect_tx(bar, (kBlurTH + 6) * kBlurSW);
            for (int r = 0; r < kBlurTH + 6; ++r) {
                const int yy = reflect101(min(y0 + r - 3, g.h + 2), g.h);
                tma_bulk_g2s(tma_smem_u32(&s_in[r][0]), src + (long long)yy * g.pitch + (x0 - 16), kBlurSW, bar);
            }
        }
        tma_mbar_wait(bar, 0);
    } else {
        for (int i = tid; i < (kBlurTH + 6) * (kBlurTW + 6); i += 256) {
            const int r = i / (kBlurTW + 6), c = i - r * (kBlurTW + 6);
            const int yy = reflect101(min(y0 + r - 3, g.h + 2), g.h);
            const int xx = reflect101(min(x0 + c - 3, g.w + 2), g.w);
            s_in[r][c + kBlurSO] = src[(long long)yy * g.pitch + xx];
        }
    }
    __syncthreads();
    for (int i = tid; i < (kBlurTH + 6) * kBlurTW; i += 256) {
        const int r = i / kBlurTW, c = i - r * kBlurTW;
        const uint8_t* p = &s_in[r][c + kBlurSO];
        s_row[r][c] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
    }
    __syncthreads();
    // column pass: thread = 4 adjacent pixels of one row, two rows per thread (128*16/4/256 = 2)
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
        const int q = tid + rep * 256;
        const int r = q / (kBlurTW / 4), c4 = (q - r * (kBlurTW / 4)) * 4;
        const int y = y0 + r, x = x0 + c4;
        if (y >= g.h || x >= g.w) continue;
        uint32_t out = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t acc = 18u * (s_row[r][c4 + k] + s_row[r + 6][c4 + k]) + 34u * (s_row[r + 1][c4 + k] + s_row[r + 5][c4 + k]) +
                                 48u * (s_row[r + 2][c4 + k] + s_row[r + 4][c4 + k]) + 56u * s_row[r + 3][c4 + k];
            out |= ((acc + 32768u) >> 16) << (8 * k);
        }
        uint8_t* d = dst + (long long)y * g.pitch + x;
        if (x + 3 < g.w) *reinterpret_cast<uint32_t*>(d) = out;
        else for (int k = 0; x + k < g.w; ++k) d[k] = (uint8_t)(out >> (8 * k));
    }
}

// ---------------------------------------------------------------------------------------------
// a3  per-cell FAST-9/16 + 3x3 NMS + iniTh->minTh fallback (src/ORBextractor.cc:919-998).
//     One CTA = one cell = one cv::FAST call of the reference.  The OpenCV corner score does not
//     depend on the threshold and "corner at t" <=> score >= t, so ONE score pass serves both
//     thresholds: maxima(t) = {strict 3x3 maxima of the cell-masked score map} ∩ {score >= t}.
//     Candidates leave in the reference's order (raster inside the cell) into the cell's slots.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int fast_corner_score(const uint8_t* p, int stride, int min_th, int use_tree)
{
    const int v = p[0];
    int d[16];
    d[0] = v - p[3 * stride];       d[1] = v - p[3 * stride + 1];   d[2] = v - p[2 * stride + 2];   d[3] = v - p[stride + 3];
    d[4] = v - p[3];                d[5] = v - p[-stride + 3];      d[6] = v - p[-2 * stride + 2];  d[7] = v - p[-3 * stride + 1];
    d[8] = v - p[-3 * stride];      d[9] = v - p[-3 * stride - 1];  d[10] = v - p[-2 * stride - 2]; d[11] = v - p[-stride - 3];
    d[12] = v - p[-3];              d[13] = v - p[stride - 3];      d[14] = v - p[2 * stride - 2];  d[15] = v - p[3 * stride - 1];
    // quick reject at min_th: 16-bit masks of ring pixels darker / brighter than the centre by > min_th
    uint32_t mb = 0, md = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { mb |= (uint32_t)(d[k] > min_th) << k; md |= (uint32_t)(d[k] < -min_th) << k; }
    mb |= mb << 16; md |= md << 16;
    uint32_t rb = mb & (mb >> 1); rb &= rb >> 2; rb &= rb >> 4; rb &= mb >> 8;     // 9 contiguous ones
    uint32_t rd = md & (md >> 1); rd &= rd >> 2; rd &= rd >> 4; rd &= md >> 8;
    if (((rb | rd) & 0xffffu) == 0) return 0;
    // exact score = largest t for which the pixel is still a corner (cv::FAST's cornerScore).  Default (use_tree == 2): OpenCV's own
    // min/max sequence, ~180 VIMNMX; use_tree == 0 (PLVS_FAST_TREE=0): bisection on t with the 16-bit arc test (compares and bit logic
    // only, ~3x the instructions) -- the form round 1 shipped because a doubling-window min/max tree gave wrong scores inside this kernel
    // on the B200; the sequence below was checked on the device against the oracle's score map pixel by pixel
    // (tools/fast_tree_probe.py: 0 mismatches on all 8 levels) and by the extractor tests.
    if (use_tree == 2) {
        // cv::cornerScore<16> as written (modules/features2d/src/fast_score.cpp), without its early `continue`s (they only skip work):
        // arcs start at even k; a = min over the 8 ring values k+1..k+8 serves the two arcs {k..k+8} and {k+1..k+9}
        int a0 = min_th;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            int a = min(d[(k + 1) & 15], d[(k + 2) & 15]);
            a = min(a, d[(k + 3) & 15]); a = min(a, d[(k + 4) & 15]); a = min(a, d[(k + 5) & 15]);
            a = min(a, d[(k + 6) & 15]); a = min(a, d[(k + 7) & 15]); a = min(a, d[(k + 8) & 15]);
            a0 = max(a0, min(a, d[k]));
            a0 = max(a0, min(a, d[(k + 9) & 15]));
        }
        int b0 = -a0;
#pragma unroll
        for (int k = 0; k < 16; k += 2) {
            int b = max(d[(k + 1) & 15], d[(k + 2) & 15]);
            b = max(b, d[(k + 3) & 15]); b = max(b, d[(k + 4) & 15]); b = max(b, d[(k + 5) & 15]);
            b = max(b, d[(k + 6) & 15]); b = max(b, d[(k + 7) & 15]); b = max(b, d[(k + 8) & 15]);
            b0 = min(b0, max(b, d[k]));
            b0 = min(b0, max(b, d[(k + 9) & 15]));
        }
        return -b0 - 1;
    }
    int lo = min_th, hi = 255;
#pragma unroll 1
    while (hi - lo > 1) {
        const int t = (lo + hi) >> 1;
        uint32_t b = 0, k9 = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) { b |= (uint32_t)(d[k] > t) << k; k9 |= (uint32_t)(d[k] < -t) << k; }
        b |= b << 16; k9 |= k9 << 16;
        uint32_t x = b & (b >> 1); x &= x >> 2; x &= x >> 4; x &= b >> 8;
        uint32_t y = k9 & (k9 >> 1); y &= y >> 2; y &= y >> 4; y &= k9 >> 8;
        if ((x | y) & 0xffffu) lo = t; else hi = t;
    }
    return lo;      // <= 254
}

// One tensor map per pyramid level: a 3-D view (x, y, frame) of the level inside the batch's pyramid buffer, box = kCellPitch x kMaxCell x 1
// bytes.  The TMA engine then stages a whole FAST cell with ONE instruction issued by one thread -- cp.async.bulk.tensor, SASS UTMALDG --
// into a dense kCellPitch-pitch tile.  The innermost start coordinate of a tiled u8 load has to be a multiple of 16 bytes on the B200 (measured
// with tools/tma_probe.cu: x = 16, 32, 48, 64 load correctly, x = 53 raises "illegal instruction"), and cells begin at arbitrary x: the box
// starts at x0 & ~15 and the kernel reads the tile `x0 & 15` bytes in.  Out-of-image parts of the box arrive as zeros and are never read.
// (The CPU execution model has no tensor maps: plain loads there.)
struct PyramidMaps {
#ifndef PLVS_CUDA_EMU
    CUtensorMap level[PLVS_MAX_LEVELS];
#else
    int unused;
#endif
};

#if !defined(PLVS_CUDA_EMU) && defined(__CUDACC__)
__device__ __forceinline__ void tma_tile3d_g2s(uint32_t dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(z), "r"(bar) : "memory");
}
#endif

__global__ void __launch_bounds__(256)
k_fast_cells(const __grid_constant__ PyramidMaps maps, int use_tma, const uint8_t* __restrict__ pyr, long long frame_stride,
             const LevelGeom* __restrict__ levels, const CellDesc* __restrict__ cells,
             uint32_t* __restrict__ slots, long long slots_per_frame,
             int* __restrict__ cell_count, int cells_per_frame, int ini_th, int min_th, int use_tree,
             uint8_t* __restrict__ dbg_score /* optional: score map in pyramid layout (inspection) */)
{
    __shared__ __align__(128) uint8_t s_img_raw[kCellPitch * kMaxCell];
    __shared__ uint8_t s_sc[kMaxCell * kMaxCell];
    __shared__ int s_cnt[2][8];
    __shared__ __align__(8) unsigned long long s_bar;
    const CellDesc c = cells[blockIdx.x];
    const LevelGeom g = levels[c.level];
    const uint8_t* src = pyr + (long long)blockIdx.y * frame_stride + g.off;
    const int w = c.x1 - c.x0, h = c.y1 - c.y0;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int iw = w - 6, ih = h - 6;
    uint32_t* out = slots + (long long)blockIdx.y * slots_per_frame + c.slot_off;
    if (iw <= 0 || ih <= 0) { if (tid == 0) cell_count[blockIdx.y * cells_per_frame + blockIdx.x] = 0; return; }
    uint8_t* s_img = s_img_raw;          // cell pixel (r, cc) at s_img[r * kCellPitch + cc]
#if !defined(PLVS_CUDA_EMU)
    if (use_tma) {
        const uint32_t bar = tma_smem_u32(&s_bar);
        if (tid == 0) tma_mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            tma_mbar_expect_tx(bar, kCellPitch * kMaxCell);          // the whole box is delivered, zeros where it leaves the image
            tma_tile3d_g2s(tma_smem_u32(s_img_raw), &maps.level[c.level], c.x0 & ~15, c.y0, (int)blockIdx.y, bar);
        }
        s_img = s_img_raw + (c.x0 & 15);
        for (int i = tid; i < kMaxCell * kMaxCell / 4; i += 256) reinterpret_cast<uint32_t*>(s_sc)[i] = 0u;
        tma_mbar_wait(bar, 0);
    } else
#endif
    {
        for (int i = tid; i < w * h; i += 256) {
            const int r = i / w, cc = i - r * w;
            s_img[r * kCellPitch + cc] = src[(long long)(c.y0 + r) * g.pitch + c.x0 + cc];
            s_sc[r * kMaxCell + cc] = 0;
        }
    }
    __syncthreads();
    const int n = iw * ih;
    for (int i = tid; i < n; i += 256) {
        const int r = i / iw + 3, cc = i - (r - 3) * iw + 3;
        s_sc[r * kMaxCell + cc] = (uint8_t)fast_corner_score(&s_img[r * kCellPitch + cc], kCellPitch, min_th, use_tree);
    }
    __syncthreads();
    if (dbg_score) {
        uint8_t* dd = dbg_score + (long long)blockIdx.y * frame_stride + g.off;
        for (int i = tid; i < n; i += 256) {
            const int r = i / iw + 3, cc = i - (r - 3) * iw + 3;
            dd[(long long)(c.y0 + r) * g.pitch + c.x0 + cc] = s_sc[r * kMaxCell + cc];
        }
    }
    // each warp owns a contiguous raster range so emission order needs no block-wide scan
    const int per_warp = ((n + 7) / 8 + 31) & ~31;
    const int beg = wid * per_warp, end = min(n, beg + per_warp);
    int cnt_lo = 0, cnt_hi = 0;
    for (int i = beg + lane; i < beg + per_warp; i += 32) {
        int flag = 0;
        if (i < end) {
            const int r = i / iw + 3, cc = i - (r - 3) * iw + 3;
            const uint8_t* p = &s_sc[r * kMaxCell + cc];
            const int s = p[0];
            if (s > 0 && s > p[-1] && s > p[1] && s > p[-kMaxCell - 1] && s > p[-kMaxCell] && s > p[-kMaxCell + 1] &&
                s > p[kMaxCell - 1] && s > p[kMaxCell] && s > p[kMaxCell + 1])
                flag = s >= ini_th ? 2 : 1;
            // reuse the (now dead) image tile as the flag plane
            s_img[r * kCellPitch + cc] = (uint8_t)flag;
        }
        cnt_lo += __popc(__ballot_sync(0xffffffffu, flag != 0));
        cnt_hi += __popc(__ballot_sync(0xffffffffu, flag == 2));
    }
    if (lane == 0) { s_cnt[0][wid] = cnt_lo; s_cnt[1][wid] = cnt_hi; }
    __syncthreads();
    int tot_hi = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot_hi += s_cnt[1][k];
    const int sel = tot_hi > 0 ? 1 : 0;       // iniThFAST found something -> keep those; else minThFAST set
    int base = 0, total = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (k < wid) base += s_cnt[sel][k]; total += s_cnt[sel][k]; }
    for (int i = beg + lane; i < beg + per_warp; i += 32) {
        int keep = 0, r = 0, cc = 0;
        if (i < end) {
            r = i / iw + 3; cc = i - (r - 3) * iw + 3;
            const int flag = s_img[r * kCellPitch + cc];
            keep = sel ? (flag == 2) : (flag != 0);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = base + __popc(m & ((1u << lane) - 1));
            if (pos < c.cap) out[pos] = pack_xys(c.x0 + cc, c.y0 + r, s_sc[r * kMaxCell + cc]);
        }
        base += __popc(m);
    }
    if (tid == 0) cell_count[blockIdx.y * cells_per_frame + blockIdx.x] = min(total, c.cap);
}

// ---------------------------------------------------------------------------------------------
// Compaction: per (level, frame) CTA, exclusive scan of the level's cell counts and an ordered
// copy of the candidates (cell-row-major, raster inside the cell == vToDistributeKeys order)
// into a dense array in HBM and, mirrored, into mapped pinned host memory for the distributor.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_compact(const uint32_t* __restrict__ slots, long long slots_per_frame,
          const int* __restrict__ cell_count, int cells_per_frame,
          const LevelGeom* __restrict__ levels, const CellDesc* __restrict__ cells, int nlevels,
          uint32_t* __restrict__ cand_dev, uint32_t* __restrict__ cand_host, int* __restrict__ cand_count_dev,
          int* __restrict__ cand_count_host)
{
    __shared__ int s_off[4096 + 1], s_slot[4096];
    __shared__ int s_warp[32];
    const int level = blockIdx.x, frame = blockIdx.y;
    const LevelGeom g = levels[level];
    const int* cnt = cell_count + frame * cells_per_frame + g.cell_begin;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, T = blockDim.x, nw = T >> 5;
    const int ncell = min(g.cell_count, 4096);
    int running = 0;
    for (int base = 0; base < ncell; base += T) {
        const int i = base + tid;
        const int v = i < ncell ? cnt[i] : 0;
        if (i < ncell) s_slot[i] = cells[g.cell_begin + i].slot_off;
        int x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        int wbase = 0, chunk_total = 0;
        for (int k = 0; k < nw; ++k) { const int sv = s_warp[k]; if (k < wid) wbase += sv; chunk_total += sv; }
        if (i < ncell) s_off[i] = running + wbase + x - v;
        running += chunk_total;
        __syncthreads();
    }
    if (tid == 0) {
        s_off[ncell] = running;
        cand_count_dev[frame * nlevels + level] = running;
        cand_count_host[frame * nlevels + level] = running;
    }
    __syncthreads();
    const uint32_t* in = slots + (long long)frame * slots_per_frame;
    uint32_t* od = cand_dev + (long long)frame * slots_per_frame + g.slot_begin;
    uint32_t* oh = cand_host + (long long)frame * slots_per_frame + g.slot_begin;
    // flat copy: thread <-> output position; its cell is the last one whose offset is <= the position (empty cells share an offset with the next
    // non-empty one, which is the last of the run).  Every load is independent: one trip to L2 per thread instead of one per cell and warp.
    for (int jb = 0; jb < running; jb += 4 * T) {       // four positions per thread and step: their loads are in flight together
        uint32_t v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jb + u * T + tid;
            v[u] = 0u;
            if (j < running) {
                int lo = 0, hi = ncell;                  // s_off[lo] <= j < s_off[hi]
                while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_off[mid] <= j) lo = mid; else hi = mid; }
                v[u] = in[s_slot[lo] + (j - s_off[lo])];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = jb + u * T + tid;
            if (j < running) { od[j] = v[u]; if (cand_host) oh[j] = v[u]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// a5/a6/a8/a9  one warp per selected keypoint: IC_Angle on the unblurred level (integer moments,
// scalar cv::fastAtan2 in non-fused fp32; src/ORBextractor.cc:110-137), steered rBRIEF-256 on the
// blurred level (glibc-exact cosf/sinf, non-fused x*b+y*a, cvRound; :141-180), keypoint fix-up
// (+border is already in the coordinates, octave, size; :1036-1046) and scaling to level 0 (:1363).
// ---------------------------------------------------------------------------------------------
__constant__ int c_umax[16];

__device__ __forceinline__ float fast_atan2_deg(float y, float x)
{
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = __fmul_rn(0.9997878412794807f, s), p3 = __fmul_rn(-0.3258083974640975f, s);
    const float p5 = __fmul_rn(0.1555786518463281f, s), p7 = __fmul_rn(-0.04432655554792128f, s);
    const float eps = (float)2.2204460492503131e-16;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, eps)); c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, eps)); c2 = __fmul_rn(c, c);
        a = __fadd_rn(90.f, -__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fadd_rn(180.f, -a);
    if (y < 0) a = __fadd_rn(360.f, -a);
    return a;
}

__global__ void __launch_bounds__(256)
k_orient_describe(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur, long long frame_stride,
                  const LevelGeom* __restrict__ levels, int nlevels,
                  const uint32_t* __restrict__ sel, const int* __restrict__ sel_level_off /* (nlevels+1) per frame */,
                  int sel_cap,
                  plvs_keypoint* __restrict__ kp_dev, uint8_t* __restrict__ desc_dev,
                  plvs_keypoint* __restrict__ kp_host, uint8_t* __restrict__ desc_host,
                  const uint32_t* __restrict__ pattern /* [8][32] words: x0 | y0 << 8 | x1 << 16 | y1 << 24 (int8) of test pair 8*lane + t at [t][lane] */)
{
    const int frame = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int* loff = sel_level_off + frame * (nlevels + 1);
    if (k >= loff[nlevels]) return;
    int level = 0;
    while (k >= loff[level + 1]) ++level;
    const LevelGeom g = levels[level];
    const uint32_t rec = sel[(long long)frame * sel_cap + k];
    const int x = unpack_x(rec), y = unpack_y(rec), resp = unpack_s(rec);
    const long long fo = (long long)frame * frame_stride + g.off;

    // ---- orientation: lane <-> column u = lane-15 of the radius-15 disc
    int m10 = 0, m01 = 0;
    if (lane < 31) {
        const int u = lane - 15;
        const int vmax = c_umax[u < 0 ? -u : u];       // the disc is symmetric: |v| <= umax[|u|]
        const uint8_t* c = pyr + fo + (long long)y * g.pitch + x + u;
        int col = 0;
        for (int v = -vmax; v <= vmax; ++v) { const int I = c[(long long)v * g.pitch]; col += I; m01 += v * I; }
        m10 = u * col;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) { m10 += __shfl_xor_sync(0xffffffffu, m10, o); m01 += __shfl_xor_sync(0xffffffffu, m01, o); }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    // ---- descriptor: lane <-> byte
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    float a, b;
    libm_sincosf(__fmul_rn(angle, factorPI), &a, &b);
    const uint8_t* c = blur + fo + (long long)y * g.pitch + x;
    // the 256 test pairs: lane <-> byte, 8 pairs per lane.  Packed four int8 per word and laid out [pair][lane], a warp fetches the eight words
    // of its lanes with eight coalesced loads (a lane-strided table in constant memory costs one serialised fetch per lane and coordinate)
    uint32_t pw[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) pw[t] = __ldg(pattern + t * 32 + lane);
    int val = 0;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
        const float x0 = (float)(int)(int8_t)(pw[t] & 0xffu), y0 = (float)(int)(int8_t)((pw[t] >> 8) & 0xffu);
        const float x1 = (float)(int)(int8_t)((pw[t] >> 16) & 0xffu), y1 = (float)(int)(int8_t)(pw[t] >> 24);
        const int r0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, b), __fmul_rn(y0, a)));
        const int q0 = __float2int_rn(__fadd_rn(__fmul_rn(x0, a), -__fmul_rn(y0, b)));
        const int r1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, b), __fmul_rn(y1, a)));
        const int q1 = __float2int_rn(__fadd_rn(__fmul_rn(x1, a), -__fmul_rn(y1, b)));
        const int t0 = c[(long long)r0 * g.pitch + q0], t1 = c[(long long)r1 * g.pitch + q1];
        val |= (t0 < t1) << t;
    }
    const long long o = (long long)frame * sel_cap + k;
    desc_dev[o * 32 + lane] = (uint8_t)val;
    desc_host[o * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        plvs_keypoint kp;
        kp.x = (float)x; kp.y = (float)y;
        if (level != 0) { kp.x = __fmul_rn(kp.x, g.scale); kp.y = __fmul_rn(kp.y, g.scale); }
        kp.size = (float)g.patch_size; kp.angle = angle; kp.response = (float)resp;
        kp.octave = level; kp.class_id = -1;
        kp_dev[o] = kp;
        kp_host[o] = kp;
    }
}

}  // namespace orb
}  // namespace plvs
