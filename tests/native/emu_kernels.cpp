// TEST INFRASTRUCTURE ONLY -- runs the product's device-only kernel headers on the CPU execution model of cuda_emu.hpp, with the same launch
// geometry and the same host-side sequencing as the entry points in plvs_b200/csrc/{match,tsdf,orb}.cu, so that tests/test_emulated_kernels.py
// can compare them with the oracle in a container without a GPU.  What this does and does not show is stated in cuda_emu.hpp.
#include "cuda_emu.hpp"

#include "../../include/plvs_b200.h"

namespace {
#include "../../plvs_b200/csrc/match_common.cuh"
#include "../../plvs_b200/csrc/match_frustum.cuh"
#include "../../plvs_b200/csrc/match_init.cuh"
#include "../../plvs_b200/csrc/orb_undistort.cuh"
#include "../../plvs_b200/csrc/tsdf_hash.cuh"
#include "../../plvs_b200/csrc/mesh_kernels.cuh"
#include "../../plvs_b200/csrc/bow_kernels.cuh"

inline int div_up(int a, int b) { return (a + b - 1) / b; }

// a deliberately broken kernel for the model's own test: thread 0 publishes a value, everybody reads it, and the __syncthreads between is missing
__global__ void k_missing_barrier_demo(int* flag, int* out)
{
    if (threadIdx.x == 0) *flag = 42;
    out[threadIdx.x] = *flag;
}

ViewDev view_dev(const plvs_frame_view* v)
{
    ViewDev o;
    o.keys = v->keys; o.desc = v->desc; o.uright = v->uright; o.n = v->n;
    o.gp = GridParams{v->min_x, v->min_y, v->max_x, v->max_y, v->grid_inv_w, v->grid_inv_h};
    o.bf = v->bf;
    for (int i = 0; i < PLVS_MAX_LEVELS; ++i) { o.scale[i] = v->scale_factors[i]; o.sigma2[i] = v->level_sigma2[i]; }
    return o;
}

}  // namespace

extern "C" {

// k_build_grid alone: cell_start[GRID_CELLS + 1], sorted[n]
int emu_build_grid(const plvs_frame_view* f, int32_t* cell_start, int32_t* sorted)
{
    try {
        const ViewDev V = view_dev(f);
        std::vector<int> kp_cell(std::max(f->n, 1));
        emu::launch(dim3(1), dim3(1024), 0, [&] { k_build_grid(V.keys, V.n, V.gp, cell_start, sorted, kp_cell.data()); });
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emu: %s\n", e.what()); return -1; }
}

// plvs_match_initialization's sequence (match.cu): k_build_grid(F2) -> k_init_candidates -> k_init_resolve, redone with a larger cap when a window overflows
int emu_match_initialization(const plvs_frame_view* f1, const plvs_frame_view* f2, float* prev_matched, int window_size, float nn_ratio, int check_orientation,
                             int32_t* matches12, int* nmatches, int initial_cap)
{
    try {
        const ViewDev V1 = view_dev(f1), V2 = view_dev(f2);
        const int n1 = f1->n, n2 = f2->n;
        *nmatches = 0;
        for (int i = 0; i < n1; ++i) matches12[i] = -1;
        if (n1 == 0 || n2 == 0) return 0;
        std::vector<int> cell_start(GRID_CELLS + 1), sorted(n2), kp_cell(n2), cand_n(n1), m12(n1), m21(3 * std::max(n1, n2)), bin_of(n1), result(4);
        std::vector<float2> prev(n1), prev_out(n1);
        std::memcpy(prev.data(), prev_matched, (size_t)n1 * 8);
        emu::launch(dim3(1), dim3(1024), 0, [&] { k_build_grid(V2.keys, n2, V2.gp, cell_start.data(), sorted.data(), kp_cell.data()); });
        int cap = initial_cap;
        std::vector<uint32_t> cand;
        for (;;) {
            cand.assign((size_t)n1 * cap, 0u);
            emu::launch(dim3(div_up(n1, 8)), dim3(256), 0, [&] {
                k_init_candidates(V1, V2, cell_start.data(), sorted.data(), prev.data(), (float)window_size, cand.data(), cand_n.data(), cap); });
            emu::launch(dim3(1), dim3(32), (size_t)n2 * sizeof(uint16_t), [&] {
                k_init_resolve(cand.data(), cand_n.data(), cap, V1.keys, V2.keys, n1, n2, nn_ratio, check_orientation, m12.data(), m21.data(), bin_of.data(),
                               prev.data(), matches12, prev_out.data(), result.data()); });
            int mx = 0;
            for (int i = 0; i < n1; ++i) mx = std::max(mx, cand_n[i]);
            if (mx <= cap) break;
            while (cap < mx) cap *= 2;
        }
        std::memcpy(prev_matched, prev_out.data(), (size_t)n1 * 8);
        *nmatches = result[0];
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emu: %s\n", e.what()); return -1; }
}

// plvs_match_in_frustum's kernels: k_in_frustum, then k_compact_queries (the resident path's first step)
int emu_in_frustum(const plvs_frustum* fr, const float* thresholds, const plvs_map_point* pts, int n, plvs_mp_query* queries, uint8_t* in_view, int* n_in_view,
                   plvs_mp_query* compacted, int32_t* src_index)
{
    try {
        FrustumDev D;
        D.f = *fr;
        for (int i = 0; i < PLVS_MAX_LEVELS; ++i) D.T[i] = thresholds[i];
        int count = 0;
        if (n) emu::launch(dim3(div_up(n, 256)), dim3(256), 0, [&] { k_in_frustum(D, pts, n, queries, in_view, &count); });
        *n_in_view = count;
        if (n && compacted) emu::launch(dim3(1), dim3(1024), 0, [&] { k_compact_queries(queries, in_view, n, compacted, src_index); });
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emu: %s\n", e.what()); return -1; }
}

// plvs_orb_undistort's kernel; k14 = OpenCV's 14 distortion coefficients (zeros beyond what the camera has)
int emu_undistort(const plvs_keypoint* keys, int n, double fx, double fy, double cx, double cy, const double* k14, plvs_keypoint* out)
{
    try {
        UndistortParams P{};
        P.fx = fx; P.fy = fy; P.cx = cx; P.cy = cy;
        for (int i = 0; i < 14; ++i) P.k[i] = k14[i];
        if (n) emu::launch(dim3(div_up(n, 256)), dim3(256), 0, [&] { k_undistort_keypoints(keys, n, P, out); });
        return 0;
    } catch (const std::exception& e) { std::fprintf(stderr, "emu: %s\n", e.what()); return -1; }
}

// plvs_tsdf_update_meshes' sequence over a block pool given as arrays (block b: key block_key[3b..], voxels at b*4096): hash table built here
// with the product's hash function and probing, blocks sorted by key, k_mesh_count -> prefix -> k_mesh_emit -> k_mesh_shade.
// Returns the number of non-empty meshes; keys/counts sized for nb entries, vertex arrays for cap_verts vertices (call with 0 to size).
int emu_update_meshes(int nb, const int32_t* block_key, const float* sdf_pool, const float* w_pool, const uint32_t* rgba_pool, float res, int use_color,
                      int32_t* keys, int32_t* counts, float* verts, float* normals, float* colors, long long cap_verts, long long* n_verts)
{
    try {
        uint32_t hash_size = 64;
        while (hash_size < (uint32_t)(4 * std::max(nb, 1))) hash_size <<= 1;
        const uint32_t mask = hash_size - 1;
        std::vector<HashEntry> tab(hash_size, HashEntry{0, 0, 0, HASH_EMPTY});
        for (int b = 0; b < nb; ++b) {
            uint32_t s = hash_key(block_key[3 * b], block_key[3 * b + 1], block_key[3 * b + 2], mask);
            while (tab[s].idx != HASH_EMPTY) s = (s + 1) & mask;
            tab[s] = HashEntry{block_key[3 * b], block_key[3 * b + 1], block_key[3 * b + 2], b};
        }
        std::vector<int> list(nb);
        for (int b = 0; b < nb; ++b) list[b] = b;
        std::sort(list.begin(), list.end(), [&](int a, int b) {
            return std::lexicographical_compare(block_key + 3 * a, block_key + 3 * a + 3, block_key + 3 * b, block_key + 3 * b + 3); });
        *n_verts = 0;
        if (nb == 0) return 0;
        std::vector<int> tri(nb);
        emu::launch(dim3(nb), dim3(256), 0, [&] { k_mesh_count(list.data(), block_key, tab.data(), mask, sdf_pool, w_pool, tri.data()); });
        std::vector<long long> base(nb);
        long long nv = 0; int nm = 0;
        for (int i = 0; i < nb; ++i) {
            base[i] = nv; nv += 3ll * tri[i];
            if (tri[i]) {
                if (keys) for (int k = 0; k < 3; ++k) keys[3 * nm + k] = block_key[3 * list[i] + k];
                if (counts) counts[nm] = 3 * tri[i];
                ++nm;
            }
        }
        *n_verts = nv;
        if (!verts || nv == 0) return nm;
        if (cap_verts < nv) return -2;
        const MeshParams M{res, 1.f / res, 0.5f * res, 1.0f / ((float)16 * res), use_color};
        emu::launch(dim3(nb), dim3(256), 0, [&] { k_mesh_emit(list.data(), block_key, tab.data(), mask, sdf_pool, w_pool, base.data(), M, verts, normals, nullptr, nullptr); });
        emu::launch(dim3((unsigned)((nv + 255) / 256)), dim3(256), 0, [&] {
            k_mesh_shade(nv, tab.data(), mask, block_key, sdf_pool, w_pool, rgba_pool, M, verts, normals, colors); });
        return nm;
    } catch (const std::exception& e) { std::fprintf(stderr, "emu: %s\n", e.what()); return -1; }
}

// plvs_voc_transform's kernels (bow.cu): k_bow_descend -> k_bow_rank -> k_bow_offsets over a vocabulary given in the sibling-contiguous layout
int emu_bow_transform(const int32_t* child_off, const int32_t* child_id, const uint8_t* child_desc, const int32_t* word_id, const double* node_weight, int L,
                      const uint8_t* desc, int n, int levelsup, uint32_t* word, double* weight, uint32_t* node, uint32_t* fv_nodes, int32_t* fv_offsets,
                      int32_t* fv_features, int* n_fv_nodes)
{
    try {
        *n_fv_nodes = 0; fv_offsets[0] = 0;
        if (n == 0) return 0;
        const VocDev V{child_off, child_id, child_desc, word_id, node_weight, L};
        std::vector<uint32_t> sorted_node(n);
        int cnt[4] = {0, 0, 0, 0};
        emu::launch(dim3(div_up(n, 8)), dim3(256), 0, [&] { k_bow_descend(V, desc, n, levelsup, word, weight, node); });
        emu::launch(dim3(div_up(n, 256)), dim3(256), 0, [&] { k_bow_rank(node, weight, n, fv_features, sorted_node.data(), &cnt[0]); });
        emu::launch(dim3(1), dim3(1024), 0, [&] { k_bow_offsets(sorted_node.data(), &cnt[0], fv_nodes, fv_offsets, &cnt[1]); });
        *n_fv_nodes = cnt[1];
        return cnt[0];
    } catch (const std::exception& e) { std::fprintf(stderr, "emu: %s\n", e.what()); return -1; }
}

// the model's own test: how many of 256 threads saw the value thread 0 publishes without a barrier
int emu_missing_barrier_demo(void)
{
    try {
        int flag = 0; std::vector<int> out(256, -1);
        emu::launch(dim3(1), dim3(256), 0, [&] { k_missing_barrier_demo(&flag, out.data()); });
        int seen = 0;
        for (int v : out) seen += v == 42;
        return seen;
    } catch (const std::exception& e) { std::fprintf(stderr, "emu: %s\n", e.what()); return -1; }
}

}  // extern "C"
