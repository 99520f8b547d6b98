// ORB extractor: handle, geometry tables, launch sequence and the extern "C" entry points that
// replace PLVS2::ORBextractor (reference: include/ORBextractor.h:59-170, src/ORBextractor.cc).
#include <atomic>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <mutex>
#include <thread>
#include <vector>
#include "orb_distribute.hpp"
#include "orb_kernels.cuh"
#include "orb_distribute.cuh"
namespace framegrid {
#include "match_common.cuh"
}

using namespace plvs;
using namespace plvs::orb;

static const int h_pattern[1024] = {
#include "orb_pattern.inc"
};

struct plvs_orb {
    plvs_orb_params prm{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev = nullptr;
    // constructor tables (src/ORBextractor.cc:446-523)
    float scale[PLVS_MAX_LEVELS], inv_scale[PLVS_MAX_LEVELS], sigma2[PLVS_MAX_LEVELS], inv_sigma2[PLVS_MAX_LEVELS];
    int quota[PLVS_MAX_LEVELS];
    int umax[16];
    // geometry for the current image size
    int w = 0, h = 0, batch_cap = 0;
    std::vector<LevelGeom> lv;
    std::vector<CellDesc> cells;
    std::vector<TileDesc> blur_tiles;
    long long frame_stride = 0;      // bytes of one frame's pyramid
    long long slots_per_frame = 0;
    int sel_cap = 0;                 // keypoint capacity per frame
    DevBuf<uint8_t> d_pyr, d_blur, d_dbg_score, d_color;
    PyramidMaps maps{};              // TMA tensor maps of the pyramid levels (k_fast_cells); valid when use_tma
    bool use_tma = false;
    DevBuf<float> d_uright, d_kdepth, d_depth_img, d_keys_un_x;
    DevBuf<plvs_keypoint> d_kp_un;       // mvKeysUn of the last batch (plvs_orb_undistort)
    PinBuf<plvs_keypoint> p_kp_un;
    PinBuf<float> p_uright;
    bool debug = false;
    DevBuf<LevelGeom> d_lv;
    DevBuf<CellDesc> d_cells;
    DevBuf<TileDesc> d_tiles;
    DevBuf<BilinearTap> d_taps;
    DevBuf<uint32_t> d_slots, d_cand, d_sel, d_pattern;
    DevBuf<int> d_sel_off, d_sel_count, d_quota, d_dist_i32;
    DevBuf<unsigned long long> d_dist_u64;
    DevBuf<DNode> d_dist_nodes;
    DevBuf<uint32_t> d_dist_stage;
    DevBuf<DistLevel> d_dist_levels;
    PinBuf<int> p_nkp, p_err;
    bool host_distribute = false; int fast_tree = 0;
    int dist_smem = 0, dist_arena = 0, dist_sort_elems = 0;
    DevBuf<int> d_cell_count, d_cand_count;
    DevBuf<plvs_keypoint> d_kp;
    DevBuf<uint8_t> d_desc;
    bool grid_on = false;            // plvs_orb_set_frame_grid: AssignFeaturesToGrid at frame construction
    framegrid::GridParams grid_gp{};
    DevBuf<int> d_grid_start, d_grid_sorted, d_grid_cell;
    PinBuf<uint32_t> p_cand;         // compacted candidates (device writes, host reads)
    PinBuf<int> p_cand_count;
    PinBuf<uint32_t> p_sel;          // selected keypoints (host writes, device reads)
    PinBuf<int> p_sel_off;
    PinBuf<plvs_keypoint> p_kp;      // results (device writes, host reads)
    PinBuf<uint8_t> p_desc;
    std::vector<int> n_kp;           // per frame of the last batch
    std::vector<char> lapped;
    int last_batch = 0;
    uint64_t serial = 0, epoch = 0;
    plvs_orb_stats stats{};
    KernelTimer timer;
    std::mutex mu;
};

namespace {

int rhe(float v) { return (int)lrintf(v); }
int rhe(double v) { return (int)lrint(v); }

void build_tables(plvs_orb* o)
{
    const plvs_orb_params& p = o->prm;
    o->scale[0] = 1.0f; o->sigma2[0] = 1.0f;
    for (int i = 1; i < p.nlevels; ++i) { o->scale[i] = o->scale[i - 1] * p.scale_factor; o->sigma2[i] = o->scale[i] * o->scale[i]; }
    for (int i = 0; i < p.nlevels; ++i) { o->inv_scale[i] = 1.0f / o->scale[i]; o->inv_sigma2[i] = 1.0f / o->sigma2[i]; }
    const float factor = 1.0f / p.scale_factor;
    float want = p.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)p.nlevels));
    int sum = 0;
    for (int l = 0; l < p.nlevels - 1; ++l) { o->quota[l] = rhe(want); sum += o->quota[l]; want *= factor; }
    o->quota[p.nlevels - 1] = std::max(p.nfeatures - sum, 0);
    const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) o->umax[v] = rhe(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (o->umax[v0] == o->umax[v0 + 1]) ++v0;
        o->umax[v] = v0;
        ++v0;
    }
}

// cv::resize INTER_LINEAR coefficient tables for one axis (SURVEY.md §8c' item 1)
void axis_taps(int S, int D, BilinearTap* t)
{
    const double sc = 1. / ((double)D / S);
    for (int d = 0; d < D; ++d) {
        float f = (float)((d + 0.5) * sc - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= S - 1) { s = S - 1; f = 0.f; }
        t[d].i0 = (unsigned short)s;
        t[d].i1 = (unsigned short)std::min(s + 1, S - 1);
        t[d].c0 = (short)std::min(std::max(rhe((1.f - f) * 2048.f), -32768), 32767);
        t[d].c1 = (short)std::min(std::max(rhe(f * 2048.f), -32768), 32767);
    }
}

int setup_geometry(plvs_orb* o, int w, int h, int batch)
{
    if (o->w == w && o->h == h && batch <= o->batch_cap) return PLVS_OK;
    const int nl = o->prm.nlevels;
    if (w > 4095 || h > 4095) { set_error("image larger than 4095 px is not supported by the 12-bit candidate packing"); return PLVS_EINVAL; }
    o->lv.assign(nl, LevelGeom{});
    o->cells.clear(); o->blur_tiles.clear();
    long long off = 0;
    int tap_total = 0, slot_total = 0;
    std::vector<BilinearTap> taps;
    for (int l = 0; l < nl; ++l) {
        LevelGeom& g = o->lv[l];
        g.w = rhe((float)w * o->inv_scale[l]);              // src/ORBextractor.cc:1485-1486
        g.h = rhe((float)h * o->inv_scale[l]);
        if (g.w < 8 || g.h < 8) { set_error("pyramid level %d is %dx%d: too small", l, g.w, g.h); return PLVS_EINVAL; }
        g.pitch = (int)align_up(g.w, 128);
        g.off = off;
        off += (long long)g.pitch * g.h;
        g.scale = o->scale[l];
        g.patch_size = (int)(31 * o->scale[l]);
        g.tap_x_off = tap_total; g.tap_y_off = tap_total + g.w;
        tap_total += g.w + g.h;
        taps.resize(tap_total);
        if (l > 0) {
            axis_taps(o->lv[l - 1].w, g.w, taps.data() + g.tap_x_off);
            axis_taps(o->lv[l - 1].h, g.h, taps.data() + g.tap_y_off);
        }
        // FAST cells (src/ORBextractor.cc:872-947)
        g.cell_begin = (int)o->cells.size();
        g.slot_begin = slot_total;
        const int minB = kRoiMargin, maxBX = g.w - kEdge + 3, maxBY = g.h - kEdge + 3;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        if (width > 0 && height > 0) {
            const int nCols = (int)(width / 35.f), nRows = (int)(height / 35.f);
            if (nCols > 0 && nRows > 0) {
                const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
                for (int i = 0; i < nRows; ++i) {
                    const float iniY = (float)(minB + i * hCell);
                    float maxY = iniY + hCell + 6;
                    if (iniY >= maxBY - 3) continue;
                    if (maxY > maxBY) maxY = (float)maxBY;
                    for (int j = 0; j < nCols; ++j) {
                        const float iniX = (float)(minB + j * wCell);
                        float maxX = iniX + wCell + 6;
                        if (iniX >= maxBX - 6) continue;
                        if (maxX > maxBX) maxX = (float)maxBX;
                        CellDesc c;
                        c.level = (short)l;
                        c.x0 = (short)iniX; c.y0 = (short)iniY; c.x1 = (short)maxX; c.y1 = (short)maxY;
                        const int iw = std::max(c.x1 - c.x0 - 6, 0), ih = std::max(c.y1 - c.y0 - 6, 0);
                        c.cap = ((iw + 1) / 2) * ((ih + 1) / 2);
                        c.slot_off = slot_total;
                        slot_total += c.cap;
                        if (c.x1 - c.x0 > kMaxCell || c.y1 - c.y0 > kMaxCell) { set_error("FAST cell larger than %d px", kMaxCell); return PLVS_EINVAL; }
                        o->cells.push_back(c);
                    }
                }
            }
        }
        g.cell_count = (int)o->cells.size() - g.cell_begin;
        g.slot_count = slot_total - g.slot_begin;
        if (g.cell_count > 4096) { set_error("more than 4096 FAST cells in level %d", l); return PLVS_EINVAL; }
        for (int ty = 0; ty < div_up(g.h, kBlurTH); ++ty)
            for (int tx = 0; tx < div_up(g.w, kBlurTW); ++tx) o->blur_tiles.push_back(TileDesc{(short)l, (short)tx, (short)ty, 0});
    }
    o->frame_stride = (long long)align_up((size_t)off, 256);
    o->slots_per_frame = slot_total;
    // the distributor may return a few more than the quota per level (it finishes the current split)
    o->sel_cap = (int)align_up((size_t)(o->prm.nfeatures * 2 + 64 * nl), 32);
    o->w = w; o->h = h; o->batch_cap = std::max(batch, o->batch_cap);
    const int B = o->batch_cap;
    int rc;
    if ((rc = o->d_pyr.alloc((size_t)o->frame_stride * B))) return rc;
    if ((rc = o->d_blur.alloc((size_t)o->frame_stride * B))) return rc;
    if (o->debug) {
        if ((rc = o->d_dbg_score.alloc((size_t)o->frame_stride * B))) return rc;
        PLVS_CUDA(cudaMemsetAsync(o->d_dbg_score.p, 0, (size_t)o->frame_stride * B, o->stream));
    }
    o->use_tma = false;
#ifndef PLVS_CUDA_EMU
    {
        // one tensor map per level over (x, y, frame): k_fast_cells stages its cells with cp.async.bulk.tensor (profiles/r02_tma_fast.md);
        // PLVS_ORB_TMA=0 goes back to plain loads
        const char* e = getenv("PLVS_ORB_TMA");
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                     CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
        if (!(e && e[0] == '0') && cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn && qres == cudaDriverEntryPointSuccess) {
            bool ok = true;
            for (int l = 0; l < nl && ok; ++l) {
                const LevelGeom& g = o->lv[l];
                const cuuint64_t dims[3] = {(cuuint64_t)g.w, (cuuint64_t)g.h, (cuuint64_t)B};
                const cuuint64_t strides[2] = {(cuuint64_t)g.pitch, (cuuint64_t)o->frame_stride};
                const cuuint32_t box[3] = {(cuuint32_t)kCellPitch, (cuuint32_t)kMaxCell, 1u}, estr[3] = {1u, 1u, 1u};
                ok = ((EncodeFn)fn)(&o->maps.level[l], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, o->d_pyr.p + g.off, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
            }
            o->use_tma = ok;
        }
    }
#endif
    if ((rc = o->d_lv.alloc(nl))) return rc;
    if ((rc = o->d_cells.alloc(o->cells.size()))) return rc;
    if ((rc = o->d_tiles.alloc(o->blur_tiles.size()))) return rc;
    if ((rc = o->d_taps.alloc(taps.size()))) return rc;
    if ((rc = o->d_slots.alloc((size_t)slot_total * B))) return rc;
    if ((rc = o->d_cand.alloc((size_t)slot_total * B))) return rc;
    if ((rc = o->d_cell_count.alloc(o->cells.size() * B))) return rc;
    if ((rc = o->d_cand_count.alloc((size_t)nl * B))) return rc;
    if ((rc = o->d_kp.alloc((size_t)o->sel_cap * B))) return rc;
    if ((rc = o->d_desc.alloc((size_t)o->sel_cap * B * 32))) return rc;
    if ((rc = o->p_cand.alloc((size_t)slot_total * B))) return rc;
    if ((rc = o->p_cand_count.alloc((size_t)nl * B))) return rc;
    if ((rc = o->p_sel.alloc((size_t)o->sel_cap * B))) return rc;
    if ((rc = o->d_sel.alloc((size_t)o->sel_cap * B))) return rc;
    if ((rc = o->d_sel_off.alloc((size_t)(nl + 1) * B))) return rc;
    {
        // scratch of the device distributor: per (frame, level) carved out of a few big allocations
        size_t i32_per_frame = 0, u64_per_frame = 0, node_per_frame = 0, stage_per_frame = 0;
        int max_quota = 0;
        std::vector<int> ncap(nl);
        for (int l = 0; l < nl; ++l) {
            const size_t nmax = (size_t)o->lv[l].slot_count + 8;
            ncap[l] = 16 * o->quota[l] + 256;
            max_quota = std::max(max_quota, o->quota[l]);
            i32_per_frame += 4 * nmax + 6 * (size_t)ncap[l];            // perm[2], node_of[2] | order[2], plist, nkids, nexp, flag
            u64_per_frame += 2 * (nmax + 1) + 2 * (size_t)ncap[l];     // scan_a, scan_b | expand[2]
            node_per_frame += ncap[l];
            stage_per_frame += ncap[l];
        }
        if ((rc = o->d_dist_i32.alloc(i32_per_frame * B)) || (rc = o->d_dist_u64.alloc(u64_per_frame * B)) || (rc = o->d_dist_nodes.alloc(node_per_frame * B)) ||
            (rc = o->d_dist_stage.alloc(stage_per_frame * B)) || (rc = o->d_dist_levels.alloc((size_t)nl * B)) || (rc = o->d_sel_count.alloc((size_t)nl * B)) ||
            (rc = o->d_quota.alloc(nl)) || (rc = o->p_nkp.alloc(B)) || (rc = o->p_err.alloc(1))) return rc;
        std::vector<DistLevel> tab((size_t)nl * B);
        for (int b = 0; b < B; ++b) {
            int* pi = o->d_dist_i32.p + (size_t)b * i32_per_frame;
            unsigned long long* pu = o->d_dist_u64.p + (size_t)b * u64_per_frame;
            DNode* pn = o->d_dist_nodes.p + (size_t)b * node_per_frame;
            uint32_t* ps = o->d_dist_stage.p + (size_t)b * stage_per_frame;
            for (int l = 0; l < nl; ++l) {
                const size_t nmax = (size_t)o->lv[l].slot_count + 8;
                DistLevel& d = tab[(size_t)b * nl + l];
                d.perm[0] = pi; pi += nmax; d.perm[1] = pi; pi += nmax; d.node_of[0] = pi; pi += nmax; d.node_of[1] = pi; pi += nmax;
                d.order[0] = pi; pi += ncap[l]; d.order[1] = pi; pi += ncap[l]; d.plist = pi; pi += ncap[l];
                d.nkids = pi; pi += ncap[l]; d.nexp = pi; pi += ncap[l]; d.flag = pi; pi += ncap[l];
                d.scan_a = pu; pu += nmax + 1; d.scan_b = pu; pu += nmax + 1; d.expand[0] = pu; pu += ncap[l]; d.expand[1] = pu; pu += ncap[l];
                d.nodes = pn; pn += ncap[l]; d.stage = ps; ps += ncap[l];
                d.ncap = ncap[l];
            }
        }
        PLVS_CUDA(cudaMemcpyAsync(o->d_dist_levels.p, tab.data(), tab.size() * sizeof(DistLevel), cudaMemcpyHostToDevice, o->stream));
        PLVS_CUDA(cudaMemcpyAsync(o->d_quota.p, o->quota, nl * sizeof(int), cudaMemcpyHostToDevice, o->stream));
        PLVS_CUDA(cudaStreamSynchronize(o->stream));          // tab is a local
        o->dist_sort_elems = max_quota + 8;
        o->dist_smem = (int)align_up((size_t)o->dist_sort_elems * sizeof(unsigned long long) + (size_t)stdsort::sort_cta_scratch_ints(o->dist_sort_elems) * sizeof(int), 16);
        // single-frame calls keep the distributor's node-level state in shared memory (orb_distribute.cuh): an arena for the largest level that fits
        {
            const size_t per_node = sizeof(DNode) + 2 * 8 + 6 * 4, limit = (size_t)190 * 1024 - (size_t)o->dist_smem;
            size_t arena = 0;
            for (int l = 0; l < nl; ++l) { const size_t need = (size_t)std::min(ncap[l], 4 * o->quota[l] + 128) * per_node; if (need <= limit) arena = std::max(arena, need); }
            const char* e = getenv("PLVS_ORB_DIST_SMEM");
            o->dist_arena = (e && e[0] == '0') ? 0 : (int)arena;
        }
        if (o->dist_smem + o->dist_arena > 48 * 1024) PLVS_CUDA(cudaFuncSetAttribute(k_distribute, cudaFuncAttributeMaxDynamicSharedMemorySize, o->dist_smem + o->dist_arena));
    }
    if ((rc = o->p_sel_off.alloc((size_t)(nl + 1) * B))) return rc;
    if ((rc = o->p_kp.alloc((size_t)o->sel_cap * B))) return rc;
    if ((rc = o->p_desc.alloc((size_t)o->sel_cap * B * 32))) return rc;
    PLVS_CUDA(cudaMemcpyAsync(o->d_lv.p, o->lv.data(), nl * sizeof(LevelGeom), cudaMemcpyHostToDevice, o->stream));
    PLVS_CUDA(cudaMemcpyAsync(o->d_cells.p, o->cells.data(), o->cells.size() * sizeof(CellDesc), cudaMemcpyHostToDevice, o->stream));
    PLVS_CUDA(cudaMemcpyAsync(o->d_tiles.p, o->blur_tiles.data(), o->blur_tiles.size() * sizeof(TileDesc), cudaMemcpyHostToDevice, o->stream));
    PLVS_CUDA(cudaMemcpyAsync(o->d_taps.p, taps.data(), taps.size() * sizeof(BilinearTap), cudaMemcpyHostToDevice, o->stream));
    {
        std::vector<uint32_t> packed(256);
        for (int lane = 0; lane < 32; ++lane)
            for (int t = 0; t < 8; ++t) {
                const int* q = h_pattern + lane * 32 + 4 * t;
                packed[(size_t)t * 32 + lane] = (uint32_t)(uint8_t)(int8_t)q[0] | ((uint32_t)(uint8_t)(int8_t)q[1] << 8) | ((uint32_t)(uint8_t)(int8_t)q[2] << 16) | ((uint32_t)(uint8_t)(int8_t)q[3] << 24);
            }
        if ((rc = o->d_pattern.alloc(256))) return rc;
        PLVS_CUDA(cudaMemcpy(o->d_pattern.p, packed.data(), 256 * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    PLVS_CUDA(cudaMemcpyToSymbolAsync(c_umax, o->umax, sizeof(o->umax), 0, cudaMemcpyHostToDevice, o->stream));
    PLVS_CUDA(cudaStreamSynchronize(o->stream));
    return PLVS_OK;
}

}  // namespace

extern "C" {

int plvs_orb_create(const plvs_orb_params* p, int device, plvs_orb** out)
{
    if (!p || !out) { set_error("null argument"); return PLVS_EINVAL; }
    if (p->nlevels < 1 || p->nlevels > PLVS_MAX_LEVELS || p->nfeatures < 1 || !(p->scale_factor > 1.0f)) {
        set_error("bad ORB parameters"); return PLVS_EINVAL;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: libplvs_b200 has no CPU fallback"); return PLVS_ENODEV; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return PLVS_EINVAL; }
    PLVS_CUDA(cudaSetDevice(device));
    plvs_orb* o = new plvs_orb();
    o->prm = *p; o->device = device; o->timer.component = 1;
    { static std::atomic<uint64_t> counter{1}; o->serial = counter.fetch_add(1); }
    { const char* e = getenv("PLVS_ORB_DEBUG"); o->debug = e && e[0] == '1'; }
    { const char* e = getenv("PLVS_FAST_TREE"); o->fast_tree = e ? (std::atoi(e) ? 2 : 0) : 2; }     // 2 = cv::cornerScore's min/max sequence (default), 0 = bisection
    { const char* e = getenv("PLVS_ORB_HOST_DISTRIBUTE"); o->host_distribute = e && e[0] == '1'; }   // A/B aid: run DistributeOctTree on host threads
    build_tables(o);
    cudaError_t e1 = create_handle_stream(&o->stream, 1);
    cudaError_t e2 = e1 == cudaSuccess ? cudaEventCreateWithFlags(&o->ev, cudaEventDisableTiming) : e1;
    if (e2 != cudaSuccess) { delete o; set_error("stream/event creation failed: %s", cudaGetErrorString(e2)); return PLVS_ENODEV; }
    *out = o;
    return PLVS_OK;
}

void plvs_orb_destroy(plvs_orb* o)
{
    if (!o) return;
    cudaSetDevice(o->device);
    if (o->stream) { cudaStreamSynchronize(o->stream); cudaStreamDestroy(o->stream); }
    if (o->ev) cudaEventDestroy(o->ev);
    delete o;
}

int plvs_orb_tables(const plvs_orb* o, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2, int32_t* fpl)
{
    if (!o) return PLVS_EINVAL;
    for (int i = 0; i < o->prm.nlevels; ++i) {
        if (scale) scale[i] = o->scale[i];
        if (inv_scale) inv_scale[i] = o->inv_scale[i];
        if (sigma2) sigma2[i] = o->sigma2[i];
        if (inv_sigma2) inv_sigma2[i] = o->inv_sigma2[i];
        if (fpl) fpl[i] = o->quota[i];
    }
    return PLVS_OK;
}

static int extract_impl(plvs_orb* o, int batch, const uint8_t* gray, int w, int h, int stride,
                        size_t frame_stride_in, int on_device, int lap0, int lap1,
                        plvs_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_out, int nch, int is_rgb);

int plvs_orb_extract_batch(plvs_orb* o, int batch, const uint8_t* gray, int w, int h, int stride,
                           size_t frame_stride_in, int on_device, int lap0, int lap1,
                           plvs_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_out)
{
    return extract_impl(o, batch, gray, w, h, stride, frame_stride_in, on_device, lap0, lap1, kps, desc, cap, n_out, mono_out, 1, 0);
}

int plvs_orb_extract_batch_color(plvs_orb* o, int batch, const uint8_t* img, int w, int h, int stride, size_t frame_stride_in, int nch, int is_rgb,
                                 int on_device, int lap0, int lap1, plvs_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_out)
{
    if (nch != 3 && nch != 4) { set_error("colour extraction takes 3- or 4-channel 8-bit images"); return PLVS_EINVAL; }
    return extract_impl(o, batch, img, w, h, stride, frame_stride_in, on_device, lap0, lap1, kps, desc, cap, n_out, mono_out, nch, is_rgb);
}

static int extract_impl(plvs_orb* o, int batch, const uint8_t* gray, int w, int h, int stride,
                        size_t frame_stride_in, int on_device, int lap0, int lap1,
                        plvs_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_out, int nch, int is_rgb)
{
    if (!o || !n_out || batch < 1) { set_error("null/invalid argument"); return PLVS_EINVAL; }
    if (!gray || w <= 0 || h <= 0) { for (int b = 0; b < batch; ++b) { n_out[b] = 0; if (mono_out) mono_out[b] = -1; } return PLVS_OK; }  // operator() returns -1 on empty image
    if (stride < w * nch) { set_error("stride < width"); return PLVS_EINVAL; }
    std::lock_guard<std::mutex> lock(o->mu);
    PLVS_CUDA(cudaSetDevice(o->device));
    int rc = setup_geometry(o, w, h, batch);
    if (rc) return rc;
    const int nl = o->prm.nlevels;
    cudaStream_t st = o->stream;
    int launches = 0;

    // level 0 <- input (ComputePyramid level 0; the 19-px border frame is never read on this path)
    const LevelGeom& g0 = o->lv[0];
    if (nch > 1) {
        // cv::cvtColor(.., COLOR_*2GRAY) (src/Tracking.cc:1797-1810) fused in front of the pyramid: the colour frames are staged once
        const uint8_t* d_img = gray;
        long long fs_in = (long long)frame_stride_in;
        if (!on_device) {
            const size_t per = (size_t)stride * h;
            if ((rc = o->d_color.alloc(per * batch))) return rc;
            for (int b = 0; b < batch; ++b)
                PLVS_CUDA(cudaMemcpyAsync(o->d_color.p + per * b, gray + b * frame_stride_in, per, cudaMemcpyHostToDevice, st));
            d_img = o->d_color.p; fs_in = (long long)per;
        }
        k_color_to_gray<<<dim3(div_up(w, 256), h, batch), 256, 0, st>>>(d_img, fs_in, stride, nch, is_rgb, o->d_pyr.p, o->frame_stride, g0);
        ++launches;
    } else if (on_device) {
        for (int b = 0; b < batch; ++b)
            PLVS_CUDA(cudaMemcpy2DAsync(o->d_pyr.p + (size_t)b * o->frame_stride, g0.pitch, gray + b * frame_stride_in, stride, w, h,
                                        cudaMemcpyDeviceToDevice, st));
    } else {
        for (int b = 0; b < batch; ++b)
            PLVS_CUDA(cudaMemcpy2DAsync(o->d_pyr.p + (size_t)b * o->frame_stride, g0.pitch, gray + b * frame_stride_in, stride, w, h,
                                        cudaMemcpyHostToDevice, st));
    }
    o->timer.begin(PLVS_ORB_K_RESIZE, st);
    for (int l = 1; l < nl; ++l) {
        const LevelGeom& g = o->lv[l];
        dim3 grid(div_up(g.w, 128), div_up(g.h, 8), batch), block(32, 8);
        k_resize_level<<<grid, block, 0, st>>>(o->d_pyr.p, o->frame_stride, o->lv[l - 1], g, o->d_taps.p);
        ++launches;
    }
    o->timer.end(st);
    o->timer.begin(PLVS_ORB_K_FAST, st);
    k_fast_cells<<<dim3((unsigned)o->cells.size(), batch), 256, 0, st>>>(o->maps, o->use_tma ? 1 : 0, o->d_pyr.p, o->frame_stride, o->d_lv.p, o->d_cells.p, o->d_slots.p,
                                                                          o->slots_per_frame, o->d_cell_count.p, (int)o->cells.size(),
                                                                          o->prm.ini_th_fast, o->prm.min_th_fast, o->fast_tree, o->debug ? o->d_dbg_score.p : nullptr);
    o->timer.end(st);
    o->timer.begin(PLVS_ORB_K_COMPACT, st);
    k_compact<<<dim3(nl, batch), 1024, 0, st>>>(o->d_slots.p, o->slots_per_frame, o->d_cell_count.p, (int)o->cells.size(), o->d_lv.p, o->d_cells.p, nl,
                                                o->d_cand.p, o->host_distribute ? o->p_cand.d : nullptr, o->d_cand_count.p, o->p_cand_count.d);      // the host copy of the candidates (22 k x 4 B over PCIe per VGA frame) only feeds the host distributor
    o->timer.end(st);
    launches += 2;
    int64_t ncand = 0, nkp = 0;
    o->n_kp.assign(batch, 0);
    auto tp0 = std::chrono::steady_clock::now(), tp1 = tp0, tp2 = tp0;
    if (!o->host_distribute) {
        // ---- device path: distribute -> pack -> blur -> orient/describe, one synchronisation at the very end
        DistArgs da{};
        da.cand = o->d_cand.p; da.cand_count = o->d_cand_count.p; da.slots_per_frame = o->slots_per_frame; da.levels = o->d_lv.p; da.nlevels = nl;
        da.quota = o->d_quota.p; da.scratch = o->d_dist_levels.p; da.sel = o->d_sel.p; da.sel_count = o->d_sel_count.p; da.sel_cap = o->sel_cap;
        da.error = o->p_err.d;
        o->p_err.h[0] = 0;
        o->timer.begin(PLVS_ORB_K_DISTRIBUTE, st);
        da.sort_bytes = o->dist_smem; da.sort_elems = o->dist_sort_elems;
        { const char* e = getenv("PLVS_ORB_DIST_ARENA_NODES"); da.arena_nodes = e ? std::max(0, atoi(e)) : 0; }      // test hook
        da.arena_bytes = batch == 1 ? o->dist_arena : 0;          // batches are throughput-bound and share the SMs' shared memory with the other stages
        k_distribute<<<dim3(nl, batch), kDistThreads, o->dist_smem + da.arena_bytes, st>>>(da);
        k_pack_selected<<<batch, 256, 0, st>>>(da, o->d_sel_off.p, o->p_nkp.d);
        o->timer.end(st);
        o->timer.begin(PLVS_ORB_K_BLUR, st);
        k_blur<<<dim3((unsigned)o->blur_tiles.size(), batch), 256, 0, st>>>(o->d_pyr.p, o->d_blur.p, o->frame_stride, o->d_lv.p, o->d_tiles.p);
        o->timer.end(st);
        o->timer.begin(PLVS_ORB_K_DESCRIBE, st);
        k_orient_describe<<<dim3(div_up(o->sel_cap, 8), batch), 256, 0, st>>>(o->d_pyr.p, o->d_blur.p, o->frame_stride, o->d_lv.p, nl, o->d_sel.p, o->d_sel_off.p,
                                                                              o->sel_cap, o->d_kp.p, o->d_desc.p, o->p_kp.d, o->p_desc.d, o->d_pattern.p);
        o->timer.end(st);
        launches += 4;
        if (o->grid_on) {
            if ((rc = o->d_grid_start.alloc((size_t)(framegrid::GRID_CELLS + 1) * o->batch_cap)) || (rc = o->d_grid_sorted.alloc((size_t)o->sel_cap * o->batch_cap)) ||
                (rc = o->d_grid_cell.alloc((size_t)o->sel_cap * o->batch_cap))) return rc;
            o->timer.begin(PLVS_ORB_K_GRID, st);
            framegrid::k_build_grid_batch<<<batch, 1024, 0, st>>>(o->d_kp.p, o->sel_cap, o->d_sel_off.p + nl, nl + 1, o->grid_gp, o->d_grid_start.p, o->d_grid_sorted.p, o->d_grid_cell.p);
            o->timer.end(st);
            ++launches;
        }
        PLVS_CUDA(cudaGetLastError());
        tp0 = tp1 = tp2 = std::chrono::steady_clock::now();
        PLVS_CUDA(cudaStreamSynchronize(st));
        if (o->p_err.h[0]) { set_error("keypoint distributor ran out of %s", o->p_err.h[0] == 1 ? "node pool" : "keypoint slots"); return PLVS_ENOMEM; }
        for (int b = 0; b < batch; ++b) {
            o->n_kp[b] = o->p_nkp.h[b]; nkp += o->n_kp[b];
            for (int l = 0; l < nl; ++l) ncand += o->p_cand_count.h[b * nl + l];
        }
    } else {
    PLVS_CUDA(cudaEventRecord(o->ev, st));
    o->timer.begin(PLVS_ORB_K_BLUR, st);
    // the blur does not depend on the keypoints: it overlaps the host-side distribution
    k_blur<<<dim3((unsigned)o->blur_tiles.size(), batch), 256, 0, st>>>(o->d_pyr.p, o->d_blur.p, o->frame_stride, o->d_lv.p, o->d_tiles.p);
    o->timer.end(st);
    ++launches;
    PLVS_CUDA(cudaGetLastError());
    tp0 = std::chrono::steady_clock::now();
    PLVS_CUDA(cudaEventSynchronize(o->ev));
    tp1 = std::chrono::steady_clock::now();

    // ---- DistributeOctTree per (frame, level) on host threads (PLVS_ORB_HOST_DISTRIBUTE=1: A/B aid for the device kernel)
    const int ntask = batch * nl;
    std::vector<std::vector<int>> picked(ntask);
    {
        const int nthreads = std::max(1, std::min<int>({ntask, 16, (int)std::thread::hardware_concurrency()}));
        std::vector<std::thread> pool;
        std::atomic<int> next{0};
        auto work = [&]() {
            Distributor dist;
            std::vector<int> xs, ys, rs;
            for (;;) {
                const int t = next.fetch_add(1);
                if (t >= ntask) break;
                const int b = t / nl, l = t % nl;
                const LevelGeom& g = o->lv[l];
                const int n = o->p_cand_count.h[b * nl + l];
                const uint32_t* c = o->p_cand.h + (size_t)b * o->slots_per_frame + g.slot_begin;
                xs.resize(n); ys.resize(n); rs.resize(n);
                for (int i = 0; i < n; ++i) { xs[i] = unpack_x(c[i]) - kRoiMargin; ys[i] = unpack_y(c[i]) - kRoiMargin; rs[i] = unpack_s(c[i]); }
                dist.run(n, xs.data(), ys.data(), rs.data(), kRoiMargin, g.w - kEdge + 3, kRoiMargin, g.h - kEdge + 3, o->quota[l], picked[t]);
            }
        };
        if (nthreads == 1) work();
        else { for (int i = 0; i < nthreads; ++i) pool.emplace_back(work); for (auto& th : pool) th.join(); }
    }
    tp2 = std::chrono::steady_clock::now();
    int max_k = 0;
    for (int b = 0; b < batch; ++b) {
        int* loff = o->p_sel_off.h + (size_t)b * (nl + 1);
        uint32_t* sel = o->p_sel.h + (size_t)b * o->sel_cap;
        int k = 0;
        for (int l = 0; l < nl; ++l) {
            loff[l] = k;
            const uint32_t* c = o->p_cand.h + (size_t)b * o->slots_per_frame + o->lv[l].slot_begin;
            ncand += o->p_cand_count.h[b * nl + l];
            for (int id : picked[b * nl + l]) { if (k < o->sel_cap) sel[k] = c[id]; ++k; }
        }
        if (k > o->sel_cap) { set_error("internal keypoint capacity %d exceeded (%d)", o->sel_cap, k); return PLVS_ENOMEM; }
        loff[nl] = k;
        o->n_kp[b] = k;
        nkp += k;
        max_k = std::max(max_k, k);
    }
    if (max_k > 0) {
        // selected keypoints go down in two DMA copies from pinned memory (kernels do not read host memory)
        PLVS_CUDA(cudaMemcpyAsync(o->d_sel.p, o->p_sel.h, (size_t)o->sel_cap * batch * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        PLVS_CUDA(cudaMemcpyAsync(o->d_sel_off.p, o->p_sel_off.h, (size_t)(nl + 1) * batch * sizeof(int), cudaMemcpyHostToDevice, st));
        o->timer.begin(PLVS_ORB_K_DESCRIBE, st);
        k_orient_describe<<<dim3(div_up(max_k, 8), batch), 256, 0, st>>>(o->d_pyr.p, o->d_blur.p, o->frame_stride, o->d_lv.p, nl, o->d_sel.p, o->d_sel_off.p,
                                                                          o->sel_cap, o->d_kp.p, o->d_desc.p, o->p_kp.d, o->p_desc.d, o->d_pattern.p);
        o->timer.end(st);
        ++launches;
    }
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaStreamSynchronize(st));
    }
    const auto tp3 = std::chrono::steady_clock::now();
    o->timer.collect();

    // ---- assemble (src/ORBextractor.cc:1267-1389): mono indices from the front, lapping-area ones from the back
    o->lapped.assign(batch, 0);
    int ret = PLVS_OK;
    for (int b = 0; b < batch; ++b) {
        const int n = o->n_kp[b];
        n_out[b] = n;
        if (n > cap) { ret = PLVS_ECAP; if (mono_out) mono_out[b] = 0; continue; }
        const plvs_keypoint* sk = o->p_kp.h + (size_t)b * o->sel_cap;
        const uint8_t* sd = o->p_desc.h + (size_t)b * o->sel_cap * 32;
        plvs_keypoint* dk = kps ? kps + (size_t)b * cap : nullptr;
        uint8_t* dd = desc ? desc + (size_t)b * cap * 32 : nullptr;
        int mono = 0, stereo = n - 1;
        bool any_lap = false;
        for (int i = 0; i < n; ++i) {
            const bool lap = sk[i].x >= (float)lap0 && sk[i].x <= (float)lap1;
            const int slot = lap ? stereo-- : mono++;
            any_lap |= lap;
            if (dk) dk[slot] = sk[i];
            if (dd) std::memcpy(dd + (size_t)slot * 32, sd + (size_t)i * 32, 32);
        }
        o->lapped[b] = any_lap;
        if (mono_out) mono_out[b] = mono;
    }
    if (ret == PLVS_ECAP) set_error("keypoint capacity too small");
    o->last_batch = batch;
    ++o->epoch;
    o->stats.pyramid_pixels = 0;
    for (int l = 0; l < nl; ++l) o->stats.pyramid_pixels += (int64_t)o->lv[l].w * o->lv[l].h;
    o->stats.candidates = ncand; o->stats.keypoints = nkp; o->stats.kernel_launches = launches;
    count_d2h((size_t)nkp * (sizeof(plvs_keypoint) + 32) + (size_t)batch * (nl + 1) * 4);     // keypoints + descriptors + counters written into mapped host memory by the kernels
    {
        const auto tp4 = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<float, std::milli>(b - a).count(); };
        o->stats.host_wait_candidates_ms = ms(tp0, tp1); o->stats.host_distribute_ms = ms(tp1, tp2);
        o->stats.host_wait_describe_ms = ms(tp2, tp3); o->stats.host_assemble_ms = ms(tp3, tp4);
    }
    return ret;
}

int plvs_orb_extract(plvs_orb* o, const uint8_t* gray, int w, int h, int stride, int on_device, int lap0, int lap1,
                     plvs_keypoint* kps, uint8_t* desc, int cap, int* n_out, int* mono_out)
{
    return plvs_orb_extract_batch(o, 1, gray, w, h, stride, 0, on_device, lap0, lap1, kps, desc, cap, n_out, mono_out);
}

int plvs_orb_pyramid_level(const plvs_orb* o, int frame, int level, int blurred, const uint8_t** dptr, int* w, int* h, int* pitch)
{
    if (!o || frame < 0 || frame >= o->last_batch || level < 0 || level >= o->prm.nlevels) { set_error("bad frame/level"); return PLVS_EINVAL; }
    const LevelGeom& g = o->lv[level];
    if (blurred == 2 && !o->debug) { set_error("score map needs PLVS_ORB_DEBUG=1 at create time"); return PLVS_ESTATE; }
    if (dptr) *dptr = (blurred == 2 ? o->d_dbg_score.p : blurred ? o->d_blur.p : o->d_pyr.p) + (size_t)frame * o->frame_stride + g.off;
    if (w) *w = g.w;
    if (h) *h = g.h;
    if (pitch) *pitch = g.pitch;
    return PLVS_OK;
}

int plvs_orb_pyramid_view(const plvs_orb* o, int frame, int blurred, plvs_pyramid_view* out)
{
    if (!o || !out || frame < 0 || frame >= o->last_batch) { set_error("bad frame"); return PLVS_EINVAL; }
    std::memset(out, 0, sizeof(*out));
    out->nlevels = o->prm.nlevels;
    for (int l = 0; l < o->prm.nlevels; ++l) {
        const LevelGeom& g = o->lv[l];
        out->data[l] = (blurred ? o->d_blur.p : o->d_pyr.p) + (size_t)frame * o->frame_stride + g.off;
        out->w[l] = g.w; out->h[l] = g.h; out->pitch[l] = g.pitch;
    }
    return PLVS_OK;
}

int plvs_orb_download_level(plvs_orb* o, int frame, int level, int blurred, uint8_t* host, int host_stride)
{
    const uint8_t* d; int w, h, pitch;
    int rc = plvs_orb_pyramid_level(o, frame, level, blurred, &d, &w, &h, &pitch);
    if (rc) return rc;
    if (!host || host_stride < w) { set_error("bad host buffer"); return PLVS_EINVAL; }
    std::lock_guard<std::mutex> lock(o->mu);
    PLVS_CUDA(cudaSetDevice(o->device));
    PLVS_CUDA(cudaMemcpy2DAsync(host, host_stride, d, pitch, w, h, cudaMemcpyDeviceToHost, o->stream));
    PLVS_CUDA(cudaStreamSynchronize(o->stream));
    return PLVS_OK;
}

int plvs_orb_stereo_from_rgbd(plvs_orb* o, int frame, const float* depth, int w, int h, int stride_bytes, int on_device, float bf,
                              const float* keys_un_x, float* uright, float* depth_out, const float** d_uright)
{
    if (!o || !depth || frame < 0 || frame >= o->last_batch || w <= 0 || h <= 0 || stride_bytes < w * 4 || (stride_bytes & 3)) { set_error("bad argument"); return PLVS_EINVAL; }
    if (o->lapped[frame]) { set_error("device keypoints unavailable: they were reordered by the lapping area"); return PLVS_ESTATE; }
    std::lock_guard<std::mutex> lock(o->mu);
    PLVS_CUDA(cudaSetDevice(o->device));
    cudaStream_t st = o->stream;
    const int n = o->n_kp[frame];
    int rc;
    const size_t cap = (size_t)std::max(o->sel_cap, 1);
    if ((rc = o->d_uright.alloc(cap * o->last_batch)) || (rc = o->d_kdepth.alloc(cap * o->last_batch)) || (rc = o->p_uright.alloc(2 * cap))) return rc;
    const float* d_depth = depth;
    if (!on_device) {
        if ((rc = o->d_depth_img.alloc((size_t)stride_bytes / 4 * h))) return rc;
        PLVS_CUDA(cudaMemcpyAsync(o->d_depth_img.p, depth, (size_t)stride_bytes * h, cudaMemcpyHostToDevice, st));
        d_depth = o->d_depth_img.p;
    }
    const float* d_un = nullptr;
    if (keys_un_x && n) {
        if ((rc = o->d_keys_un_x.alloc(cap))) return rc;
        PLVS_CUDA(cudaMemcpyAsync(o->d_keys_un_x.p, keys_un_x, (size_t)n * 4, cudaMemcpyHostToDevice, st));
        d_un = o->d_keys_un_x.p;
    }
    float* du = o->d_uright.p + cap * frame; float* dd = o->d_kdepth.p + cap * frame;
    if (n) k_stereo_from_rgbd<<<div_up(n, 256), 256, 0, st>>>(o->d_kp.p + (size_t)frame * o->sel_cap, n, d_un, d_depth, w, h, stride_bytes / 4, bf, du, dd);
    if (uright || depth_out) {
        PLVS_CUDA(cudaMemcpyAsync(o->p_uright.h, du, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
        PLVS_CUDA(cudaMemcpyAsync(o->p_uright.h + cap, dd, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    }
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaStreamSynchronize(st));
    if (uright) std::memcpy(uright, o->p_uright.h, (size_t)n * 4);
    if (depth_out) std::memcpy(depth_out, o->p_uright.h + cap, (size_t)n * 4);
    if (d_uright) *d_uright = du;
    return PLVS_OK;
}

int plvs_orb_undistort(plvs_orb* o, int frame, const float K[4], const float* dist, int ndist, plvs_keypoint* keys_un, const plvs_keypoint** d_keys_un)
{
    if (!o || !K || frame < 0 || frame >= o->last_batch || ndist < 0 || ndist > 14 || (ndist && !dist)) { set_error("bad argument"); return PLVS_EINVAL; }
    if (o->lapped[frame]) { set_error("device keypoints unavailable: they were reordered by the lapping area"); return PLVS_ESTATE; }
    std::lock_guard<std::mutex> lock(o->mu);
    PLVS_CUDA(cudaSetDevice(o->device));
    cudaStream_t st = o->stream;
    const int n = o->n_kp[frame];
    const size_t cap = (size_t)std::max(o->sel_cap, 1);
    int rc;
    if ((rc = o->d_kp_un.alloc(cap * o->last_batch)) || (rc = o->p_kp_un.alloc(cap))) return rc;
    const plvs_keypoint* src = o->d_kp.p + (size_t)frame * o->sel_cap;
    plvs_keypoint* dst = o->d_kp_un.p + cap * frame;
    if (ndist == 0 || dist[0] == 0.0f) {           // mDistCoef(0) == 0: mvKeysUn = mvKeys (src/Frame.cc:1510-1514)
        if (n) PLVS_CUDA(cudaMemcpyAsync(dst, src, (size_t)n * sizeof(plvs_keypoint), cudaMemcpyDeviceToDevice, st));
    } else if (n) {
        UndistortParams P{};
        P.fx = K[0]; P.fy = K[1]; P.cx = K[2]; P.cy = K[3];
        for (int i = 0; i < ndist; ++i) P.k[i] = dist[i];
        k_undistort_keypoints<<<div_up(n, 256), 256, 0, st>>>(src, n, P, dst);
    }
    if (keys_un && n) PLVS_CUDA(cudaMemcpyAsync(o->p_kp_un.h, dst, (size_t)n * sizeof(plvs_keypoint), cudaMemcpyDeviceToHost, st));
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaStreamSynchronize(st));
    if (keys_un && n) std::memcpy(keys_un, o->p_kp_un.h, (size_t)n * sizeof(plvs_keypoint));
    if (d_keys_un) *d_keys_un = dst;
    return PLVS_OK;
}

int plvs_orb_device_result(const plvs_orb* o, int frame, plvs_orb_device_view* out)
{
    if (!o || !out || frame < 0 || frame >= o->last_batch) { set_error("bad frame"); return PLVS_EINVAL; }
    if (o->lapped[frame]) { set_error("device view unavailable: keypoints were reordered by the lapping area"); return PLVS_ESTATE; }
    out->n = o->n_kp[frame];
    out->keys = o->d_kp.p + (size_t)frame * o->sel_cap;
    out->desc = o->d_desc.p + (size_t)frame * o->sel_cap * 32;
    out->cache_key = (o->serial << 44) | ((o->epoch & 0xfffffffffull) << 8) | (uint64_t)(frame & 0xff);
    const bool grid = o->grid_on && !o->host_distribute && o->d_grid_start.p;
    out->grid_cell_start = grid ? o->d_grid_start.p + (size_t)frame * (framegrid::GRID_CELLS + 1) : nullptr;
    out->grid_sorted = grid ? o->d_grid_sorted.p + (size_t)frame * o->sel_cap : nullptr;
    return PLVS_OK;
}

int plvs_orb_set_frame_grid(plvs_orb* o, const float bounds[6])
{
    if (!o) { set_error("null handle"); return PLVS_EINVAL; }
    if (!bounds) { o->grid_on = false; return PLVS_OK; }
    if (!(bounds[2] > bounds[0]) || !(bounds[3] > bounds[1]) || !(bounds[4] > 0.f) || !(bounds[5] > 0.f)) { set_error("bad grid bounds"); return PLVS_EINVAL; }
    o->grid_gp = framegrid::GridParams{bounds[0], bounds[1], bounds[2], bounds[3], bounds[4], bounds[5]};
    o->grid_on = true;
    return PLVS_OK;
}

int plvs_orb_candidates(const plvs_orb* o, int frame, int level, int32_t* x, int32_t* y, int32_t* score, int cap, int* n_out)
{
    if (!o || !n_out || frame < 0 || frame >= o->last_batch || level < 0 || level >= o->prm.nlevels) { set_error("bad frame/level"); return PLVS_EINVAL; }
    const int nl = o->prm.nlevels;
    const int n = o->p_cand_count.h[frame * nl + level];
    *n_out = n;
    if (n > cap) { set_error("candidate capacity too small"); return PLVS_ECAP; }
    const uint32_t* c = o->p_cand.h + (size_t)frame * o->slots_per_frame + o->lv[level].slot_begin;
    if (!o->host_distribute && n > 0) {       // inspection call: the candidates stayed on the device, fetch this level now
        cudaSetDevice(o->device);
        if (cudaMemcpy(const_cast<uint32_t*>(c), o->d_cand.p + (size_t)frame * o->slots_per_frame + o->lv[level].slot_begin, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("candidate download failed"); return PLVS_ENODEV;
        }
    }
    for (int i = 0; i < n; ++i) {
        if (x) x[i] = unpack_x(c[i]);
        if (y) y[i] = unpack_y(c[i]);
        if (score) score[i] = unpack_s(c[i]);
    }
    return PLVS_OK;
}

int plvs_orb_kernel_times(plvs_orb* o, float* ms, int32_t* launches, int reset)
{
    if (!o) return PLVS_EINVAL;
    std::lock_guard<std::mutex> lock(o->mu);
    for (int i = 0; i < KernelTimer::kSlots; ++i) { if (ms) ms[i] = o->timer.ms[i]; if (launches) launches[i] = o->timer.count[i]; }
    if (reset) o->timer.reset();
    return PLVS_OK;
}

int plvs_orb_last_stats(const plvs_orb* o, plvs_orb_stats* out)
{
    if (!o || !out) return PLVS_EINVAL;
    *out = o->stats;
    return PLVS_OK;
}

}  // extern "C"
