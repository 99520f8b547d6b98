// 256-bit Hamming matching: the three ORBmatcher entry points of the hot path behind the C ABI.
// reference: src/ORBmatcher.cc:71-157 (SearchByProjection, frame <- local map points),
//            :1774-1993 (SearchByProjection, current <- last frame), :999-1242 (SearchForTriangulation),
//            :2198-2225 (DescriptorDistance), src/Frame.cc:716-746,1231-1316 (feature grid).
//
// Design (B200): descriptors are 32-byte rows read as 2 x uint4; distances are 8 x __popc; one
// warp serves one query and keeps the reference's candidate order with ballot compaction.  The
// reference's searches are *sequential*: a map point claims a keypoint and later points skip it.
// Instead of serialising, phase A computes every query's ordered candidate list (all the Hamming
// work, fully parallel) and phase B resolves the claims by fixed-point iteration: each round every
// query re-selects assuming the claims of LOWER-numbered queries from the previous round.  After
// round k the first k queries are final (their inputs no longer change), so the fixed point is
// exactly the sequential result; real frames converge in 2-4 rounds.
#include <mutex>
#include <vector>
#include "common.cuh"

using namespace plvs;

namespace {

#include "match_common.cuh"
#include "match_lines.cuh"

// ---------------------------------------------------------------------------------------------
// Phase A (both projection searches): one warp per query walks the window column by column
// (ix outer, iy inner, insertion order inside a cell == the order GetFeaturesInArea returns),
// applies the static gates and stores (idx, dist, octave) in that order.
// mode 0: map points (levels [l-1,l], radius by viewing cosine, xR gate with r*scale)
// mode 1: last frame (level window by motion direction, radius th*scale, xR gate with bf*invz)
// ---------------------------------------------------------------------------------------------
// What a query picks from its best / second-best unblocked candidates (keys = distance << 16 | list position, 0xffffffff = none; e1 / e2 = the
// list entries they point at): src/ORBmatcher.cc:130-150 for the map search (TH_HIGH, then the level-aware ratio test), :1940-1958 for the
// last-frame search (ORBdist / TH_HIGH only).  Returns the keypoint index or -1.
template <int MODE>
__device__ __forceinline__ int decide_target(uint32_t k1, uint32_t k2, uint32_t e1, uint32_t e2, float nn_ratio, int th_high)
{
    const int bestDist = k1 == 0xffffffffu ? 256 : (int)(k1 >> 16);
    if (bestDist > th_high) return -1;
    if (MODE == 0) {
        int bestDist2 = 256, bestLevel2 = -1;
        if (k2 != 0xffffffffu) { bestDist2 = (int)(k2 >> 16); bestLevel2 = cand_level(e2); }
        const int bestLevel = cand_level(e1);
        if (!(bestLevel == bestLevel2 && (float)bestDist > nn_ratio * (float)bestDist2) &&
            (bestLevel != bestLevel2 || (float)bestDist <= nn_ratio * (float)bestDist2)) return cand_idx(e1);
        return -1;
    }
    return cand_idx(e1);
}

// per query, what round 0 of the claim resolution (nobody has claimed anything yet; pre-claimed keypoints are blocked) leaves behind
struct Round0 { int32_t target; uint16_t w1, w2, nwatch, m; int32_t pad; };   // chosen keypoint (-1: none); the keypoints of the best / second-best unblocked
                                                                             // candidate (the first watch set, nwatch of them valid); list length min(count, cap)
static_assert(sizeof(Round0) == 16, "Round0 layout");

template <int MODE>
__global__ void __launch_bounds__(256)
k_candidates(ViewDev F, const int* __restrict__ cell_start, const int* __restrict__ sorted,
             const void* __restrict__ queries, int nq, float th, int far_points, float th_far, int forward, int backward,
             uint32_t* __restrict__ cand, int* __restrict__ cand_n, int cap, int* __restrict__ max_count,
             const uint8_t* __restrict__ claimed_in, float nn_ratio, int th_high, Round0* __restrict__ round0)
{
    // The window of GetFeaturesInArea is a run of grid columns; the cells r0..r1 of one column are contiguous in the cell-column-major
    // CSR, so the reference's visiting order is the concatenation of one [pbeg, pend) range per column.  The lanes fetch the ranges of up
    // to 32 columns at once, a warp scan turns them into one flat index space, and the warp then walks 32 flat positions per step: the
    // dependent-load chain is cell_start -> sorted -> {keypoint, descriptor} per 32 candidates instead of per column.
    __shared__ int s_beg[8][32], s_excl[8][33];
    PLVS_GRID_DEP_LAUNCH();                 // the resolve CTA may become resident (and set its tables up) while the windows are walked
    const int w = threadIdx.x >> 5;
    const int q = blockIdx.x * 8 + w;
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    float px, py, radius, xr_ref, xr_tol;
    int minL, maxL;
    const uint8_t* qd;
    bool active = true;
    if (MODE == 0) {
        const plvs_mp_query& m = reinterpret_cast<const plvs_mp_query*>(queries)[q];
        if (far_points && m.track_depth > th_far) active = false;
        float r = ((double)m.view_cos > 0.998) ? 2.5f : 4.0f;
        if (th != 1.0f) r *= th;
        radius = r * F.scale[m.level];
        px = m.proj_x; py = m.proj_y; minL = m.level - 1; maxL = m.level;
        xr_ref = m.proj_xr; xr_tol = radius; qd = m.desc;
    } else {
        const plvs_last_query& m = reinterpret_cast<const plvs_last_query*>(queries)[q];
        if (m.invz < 0) active = false;
        if (m.u < F.gp.min_x || m.u > F.gp.max_x || m.v < F.gp.min_y || m.v > F.gp.max_y) active = false;
        radius = th * F.scale[m.last_octave];
        px = m.u; py = m.v;
        if (forward && backward) { minL = m.last_octave - 1; maxL = m.last_octave; }      // Sim3 searches: [l-1, l] around the predicted level (:588-590)
        else if (forward) { minL = m.last_octave; maxL = -1; }
        else if (backward) { minL = 0; maxL = m.last_octave; }
        else { minL = m.last_octave - 1; maxL = m.last_octave + 1; }
        xr_ref = m.u - F.bf * m.invz; xr_tol = radius; qd = m.desc;
    }
    int c0, c1, r0, r1, count = 0;
    uint32_t best1 = 0xffffffffu, best2 = 0xffffffffu;      // round 0: the two smallest (distance, position) keys among the candidates not pre-claimed
    uint32_t* out = cand + (size_t)q * cap;
    if (active) active = cell_window(F.gp, px, py, radius, c0, c1, r0, r1);
    if (active) {
        const bool check = (minL > 0) || (maxL >= 0);
        // query records are 60/56-byte structs: the embedded descriptor is only 4-byte aligned
        const uint32_t* qw = reinterpret_cast<const uint32_t*>(qd);
        const uint4 a0 = make_uint4(qw[0], qw[1], qw[2], qw[3]), a1 = make_uint4(qw[4], qw[5], qw[6], qw[7]);
        for (int cb = c0; cb <= c1; cb += 32) {            // at most two chunks: the grid has 64 columns
            const int ix = cb + lane;
            int pbeg = 0, cnt = 0;
            if (ix <= c1) { pbeg = cell_start[ix * GRID_ROWS + r0]; cnt = cell_start[ix * GRID_ROWS + r1 + 1] - pbeg; }
            int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
            const int total = __shfl_sync(0xffffffffu, incl, 31);
            __syncwarp();
            s_beg[w][lane] = pbeg; s_excl[w][lane] = incl - cnt;
            if (lane == 31) s_excl[w][32] = total;
            __syncwarp();
            for (int base = 0; base < total; base += 32) {
                const int f = base + lane;
                bool ok = f < total;
                int idx = 0, oct = 0, dist = 0;
                if (ok) {
                    int lo = 0;                                   // the column whose range holds flat position f: last j with excl[j] <= f
#pragma unroll
                    for (int step = 16; step; step >>= 1) if (s_excl[w][lo + step] <= f) lo += step;
                    idx = sorted[s_beg[w][lo] + (f - s_excl[w][lo])];
                    const plvs_keypoint kp = F.keys[idx];
                    oct = kp.octave;
                    if (check && (oct < minL || oct > maxL)) ok = false;     // the maxLevel test applies even for -1 (src/Frame.cc:1283-1286)
                    if (ok && !(fabsf(kp.x - px) < radius && fabsf(kp.y - py) < radius)) ok = false;
                    if (ok && F.uright) { const float ur = F.uright[idx]; if (ur > 0 && fabsf(xr_ref - ur) > xr_tol) ok = false; }
                    if (ok) dist = hamming256(a0, a1, F.desc + (size_t)idx * 32);
                }
                const uint32_t m = __ballot_sync(0xffffffffu, ok);
                uint32_t key = 0xffffffffu;
                if (ok) {
                    const int pos = count + __popc(m & ((1u << lane) - 1));
                    if (pos < cap) out[pos] = pack_cand(idx, dist, oct);
                    if (pos < 65536 && !(claimed_in && claimed_in[idx])) key = ((uint32_t)dist << 16) | (uint32_t)pos;
                }
                count += __popc(m);
                // the two smallest keys of this step (positions are unique, so keys are), merged into the running pair
                const uint32_t s1 = __reduce_min_sync(0xffffffffu, key);
                const uint32_t s2 = __reduce_min_sync(0xffffffffu, key == s1 ? 0xffffffffu : key);
                const uint32_t lo = min(best1, s1), hi = max(best1, s1);
                best2 = min(hi, min(best2, s2));
                best1 = lo;
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        cand_n[q] = count;
        if (max_count && count > cap) atomicMax(max_count, count);
        Round0 r; r.pad = 0; r.w1 = r.w2 = 0; r.nwatch = 0;
        r.m = (uint16_t)min(count, min(cap, 65535));
        r.target = -1;
        if (count <= cap) {           // otherwise the search is repeated with more room and this record is not used
            __syncwarp(1u);
            const uint32_t e1 = best1 != 0xffffffffu ? out[best1 & 0xffffu] : 0u, e2 = best2 != 0xffffffffu ? out[best2 & 0xffffu] : 0u;
            r.target = decide_target<MODE>(best1, best2, e1, e2, nn_ratio, th_high);
            if (best1 != 0xffffffffu) { r.w1 = (uint16_t)cand_idx(e1); r.nwatch = 1; }
            if (MODE == 0 && best2 != 0xffffffffu) { r.w2 = (uint16_t)cand_idx(e2); r.nwatch = 2; }
        }
        round0[q] = r;
    }
}

// ---------------------------------------------------------------------------------------------
// ORBmatcher::Fuse, search part (src/ORBmatcher.cc:1340-1406): one warp per map point.  Window from
// KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:1179-1229: no level filter, strict |dx|,|dy| < r), then the level gate
// [l-1, l], the chi-square gate on the reprojection error and the best Hamming distance; `dist < bestDist` is strict, so the
// first candidate in the reference's order (cell-column-major, insertion order inside a cell) wins ties: lanes keep
// (dist, position) keys and the warp takes the minimum.
// ---------------------------------------------------------------------------------------------
struct FuseSigma { float inv[PLVS_MAX_LEVELS]; };

template <bool CHI2>
__global__ void __launch_bounds__(256)
k_fuse(ViewDev K, FuseSigma sg, const int* __restrict__ cell_start, const int* __restrict__ sorted,
       const plvs_fuse_query* __restrict__ queries, int nq, float th, int32_t* __restrict__ best_idx, int32_t* __restrict__ best_dist,
       int* __restrict__ nfused)
{
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= nq) return;
    const plvs_fuse_query& m = queries[q];
    const int lvl = m.level;
    const float u = m.u, v = m.v, ur = m.ur;
    const float radius = th * K.scale[lvl];
    uint32_t best = 0xffffffffu;          // dist:9 << 20 | position:20
    int best_i = -1;
    int c0, c1, r0, r1;
    if (cell_window(K.gp, u, v, radius, c0, c1, r0, r1)) {
        const uint32_t* qw = reinterpret_cast<const uint32_t*>(m.desc);
        const uint4 a0 = make_uint4(qw[0], qw[1], qw[2], qw[3]), a1 = make_uint4(qw[4], qw[5], qw[6], qw[7]);
        int pos0 = 0;
        for (int ix = c0; ix <= c1; ++ix) {
            const int pbeg = cell_start[ix * GRID_ROWS + r0], pend = cell_start[ix * GRID_ROWS + r1 + 1];
            for (int p = pbeg + lane; p < pend; p += 32) {
                const int idx = sorted[p];
                const plvs_keypoint kp = K.keys[idx];
                if (!(fabsf(kp.x - u) < radius && fabsf(kp.y - v) < radius)) continue;
                const int kl = kp.octave;
                if (kl < lvl - 1 || kl > lvl) continue;
                if (CHI2) {                                  // Fuse(pKF, vpMapPoints, ...) only; the Sim3 overload has no such gate (:1514-1531)
                    const float ex = u - kp.x, ey = v - kp.y;
                    const float kr = K.uright ? K.uright[idx] : -1.f;
                    if (kr >= 0) {
                        const float er = ur - kr;
                        const float e2 = ex * ex + ey * ey + er * er;
                        if ((double)(e2 * sg.inv[kl]) > 7.8) continue;
                    } else {
                        const float e2 = ex * ex + ey * ey;
                        if ((double)(e2 * sg.inv[kl]) > 5.99) continue;
                    }
                }
                const uint32_t key = ((uint32_t)hamming256(a0, a1, K.desc + (size_t)idx * 32) << 20) | (uint32_t)(pos0 + (p - pbeg));
                if (key < best) { best = key; best_i = idx; }
            }
            pos0 += pend - pbeg;
        }
    }
    const uint32_t wmin = __reduce_min_sync(0xffffffffu, best);
    const uint32_t who = __ballot_sync(0xffffffffu, best == wmin && best != 0xffffffffu);
    if (who) {
        const int src = __ffs(who) - 1;
        const int idx = __shfl_sync(0xffffffffu, best_i, src);
        if (lane == 0) { best_idx[q] = idx; best_dist[q] = (int)(wmin >> 20); if ((int)(wmin >> 20) <= TH_LOW) atomicAdd(nfused, 1); }
    } else if (lane == 0) { best_idx[q] = -1; best_dist[q] = 256; }
}

// ---------------------------------------------------------------------------------------------
// Phase B: claim resolution.  The reference walks the queries in order and a keypoint taken by an earlier query
// (one whose map point has Observations() > 0) is skipped by the later ones (src/ORBmatcher.cc:113-115,1848-1850), so
// target(q) is a function of target(0..q-1).  It is evaluated as a fixed-point (Jacobi) iteration on ONE thread-block
// cluster of 8 co-scheduled CTAs: claim[idx] = lowest query (with Observations() > 0) currently targeting idx; in every
// round each query re-walks its candidate list skipping the candidates claimed by a LOWER query; rounds repeat until no
// target moves.  By induction on q the fixed point is the sequential result; the number of rounds is the longest chain of
// displaced queries (5-9 on the bench stream).  A round is: claim table r from L2 into shared memory, the list walks
// (lists cached in shared memory, lpq lanes per query, 4 candidates in flight per lane), lowest-query claims into table
// r+1 with atomicMin, table r+2 cleared, ONE cluster barrier.  Measured alternatives, both slower on the bench stream:
// a strict wavefront over the 8 query ranges (8 serial release/acquire hops) and per-range shared-memory fixed points
// between global rounds (every round pays a confirming sweep).
// ---------------------------------------------------------------------------------------------
constexpr int kResolveCtas = 8;

#ifdef PLVS_CUDA_EMU        // tests/native/cuda_emu.hpp: the CTAs of the cluster are resident together on the CPU model
__device__ __forceinline__ uint32_t cluster_cta_rank() { return emu::cluster_rank(); }
__device__ __forceinline__ void cluster_sync_all() { emu::cluster_barrier(); }
#else
__device__ __forceinline__ uint32_t cluster_cta_rank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n" "barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
#endif

// SMEM == true: the candidate lists of the range are cached in shared memory (row stride cap|1: conflict-free column
// walks); SMEM == false reads them from L2 (any size).
template <int MODE, bool SMEM>
__global__ void __cluster_dims__(kResolveCtas, 1, 1) __launch_bounds__(1024)
k_resolve(const uint32_t* __restrict__ cand, const int* __restrict__ cand_n, int cap, const void* __restrict__ queries, int nq,
          const plvs_keypoint* __restrict__ keys, int n, const uint8_t* __restrict__ claimed_in, float nn_ratio, int check_ori, int th_high,
          int* claims /*3*n, L2*/, int* target /*nq*/, int* flags /*[4] rotating change flags, zeroed by the host*/,
          int32_t* assign /*n, device*/, int32_t* assign_out /*n, mapped host*/, int* result /*[0]=nmatches,[1]=rounds*/, int per_cta,
          int lpq /*lanes per query: power of two <= 32, per_cta * lpq <= 1024*/, long long* trace /*dev tool: 32 clock64 stamps per CTA, or null*/)
{
    int tr_n = 0;
    auto stamp = [&]() { if (trace && threadIdx.x == 0 && tr_n < 32) trace[cluster_cta_rank() * 32 + tr_n++] = clock64(); };
    stamp();
    PLVS_DYN_SMEM(uint32_t, s_dyn);
    __shared__ int s_count, s_hist[HISTO], s_keep[HISTO];
    const int tid = threadIdx.x;
    const int crank = (int)cluster_cta_rank();
    const int INF = 0x7fffffff;
    const int stride = cap | 1;
    int* s_blocked = reinterpret_cast<int*>(s_dyn);                             // n: 1 = taken by a query of a lower range (or pre-claimed)
    int* s_claim = s_blocked + n;                                               // n: lowest own query targeting the keypoint
    uint32_t* s_list = s_dyn + 2 * n;                                           // per_cta * stride (SMEM only)
    const int q0 = crank * per_cta;
    const int qi = tid / lpq, lane = tid & (lpq - 1);          // lpq lanes share a query: strided list walk + shuffle merge
    const int q = q0 + qi;
    const bool mine = qi < per_cta && q < nq;
    int m = 0, my_target = -1;
    uint32_t my_flags = 0;
    if (mine) {
        m = min(cand_n[q], cap);
        my_flags = MODE == 0 ? reinterpret_cast<const plvs_mp_query*>(queries)[q].flags : reinterpret_cast<const plvs_last_query*>(queries)[q].flags;
    }
    for (int i = tid; i < n; i += 1024) s_blocked[i] = (claimed_in && claimed_in[i]) ? 1 : 0;   // pre-claimed keypoints never compete
    if (SMEM) {
        // stage the candidate lists: uint4 loads, 4 in flight per thread (rows are cap*4 bytes, cap is a multiple of 4)
        const int rows = min(per_cta, max(nq - q0, 0));
        const int c4 = cap >> 2, total4 = rows * c4;
        const uint4* src4 = reinterpret_cast<const uint4*>(cand + (size_t)q0 * cap);
        for (int i0 = tid; i0 < total4; i0 += 4 * 1024) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const int i = i0 + u * 1024; v[u] = i < total4 ? src4[i] : make_uint4(0, 0, 0, 0); }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * 1024;
                if (i >= total4) continue;
                const int j = i / c4, k = (i - j * c4) * 4;
                s_list[j * stride + k] = v[u].x; s_list[j * stride + k + 1] = v[u].y; s_list[j * stride + k + 2] = v[u].z; s_list[j * stride + k + 3] = v[u].w;
            }
        }
    }
    __syncthreads();
    stamp();

    // Jacobi rounds over all queries of the cluster: read claim table r (L2 -> shared), every query re-walks its list,
    // lowest-query claims go into table r+1 (atomicMin in L2), table r+2 is cleared; ONE cluster barrier per round.
    const int gtid = crank * 1024 + tid, gthreads = kResolveCtas * 1024;
    for (int i = gtid; i < 3 * n; i += gthreads) claims[i] = INF;
    cluster_sync_all();
    stamp();
    int rounds = 0;
    for (;;) {
        const int* cur = claims + (size_t)(rounds % 3) * n;
        int* nxt = claims + (size_t)((rounds + 1) % 3) * n;
        int* clr = claims + (size_t)((rounds + 2) % 3) * n;
        for (int i = tid; i < n; i += 1024) s_claim[i] = rounds ? __ldcg(&cur[i]) : INF;
        __syncthreads();
        // The reference's running best / second best (strict `<` in list order, the old best demoted to second) are the two
        // smallest (distance, position) keys of the unblocked candidates: each lane keeps the two smallest of its strided
        // share (4 candidates in flight per step), shuffles merge the lanes.
        uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
        if (mine) {
            for (int k = lane; k < m; k += 4 * lpq) {
                uint32_t key[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int kk = k + j * lpq;
                    const bool valid = kk < m;
                    const uint32_t e = valid ? (SMEM ? s_list[qi * stride + kk] : cand[(size_t)q * cap + kk]) : 0u;
                    const int idx = valid ? cand_idx(e) : 0;
                    const bool blocked = s_blocked[idx] || s_claim[idx] < q;
                    key[j] = (!valid || blocked) ? 0xffffffffu : (((uint32_t)cand_dist(e) << 16) | (uint32_t)kk);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) { if (key[j] < k1) { k2 = k1; k1 = key[j]; } else if (key[j] < k2) k2 = key[j]; }
            }
        }
        for (int o = 1; o < lpq; o <<= 1) {
            const uint32_t a1 = __shfl_xor_sync(0xffffffffu, k1, o), a2 = __shfl_xor_sync(0xffffffffu, k2, o);
            const uint32_t lo = min(k1, a1), hi = max(k1, a1);
            k2 = min(hi, min(k2, a2));
            k1 = lo;
        }
        bool changed_here = false;
        if (mine && lane == 0) {
            int t = -1;
            const int bestDist = k1 == 0xffffffffu ? 256 : (int)(k1 >> 16);
            if (bestDist <= th_high) {
                const uint32_t e1 = SMEM ? s_list[qi * stride + (k1 & 0xffffu)] : cand[(size_t)q * cap + (k1 & 0xffffu)];
                if (MODE == 0) {
                    int bestDist2 = 256, bestLevel2 = -1;
                    if (k2 != 0xffffffffu) {
                        const uint32_t e2 = SMEM ? s_list[qi * stride + (k2 & 0xffffu)] : cand[(size_t)q * cap + (k2 & 0xffffu)];
                        bestDist2 = (int)(k2 >> 16); bestLevel2 = cand_level(e2);
                    }
                    const int bestLevel = cand_level(e1);
                    if (!(bestLevel == bestLevel2 && (float)bestDist > nn_ratio * (float)bestDist2) &&
                        (bestLevel != bestLevel2 || (float)bestDist <= nn_ratio * (float)bestDist2)) t = cand_idx(e1);
                } else t = cand_idx(e1);
            }
            if (t != my_target || rounds == 0) { my_target = t; target[q] = t; changed_here = true; }
            if (t >= 0 && (my_flags & PLVS_Q_OBS_POSITIVE)) atomicMin(&nxt[t], q);
        }
        for (int i = gtid; i < n; i += gthreads) clr[i] = INF;
        if (gtid == 0) flags[(rounds + 2) & 3] = 0;
        if (__syncthreads_or(changed_here ? 1 : 0) && tid == 0) atomicOr(&flags[rounds & 3], 1);
        cluster_sync_all();
        stamp();
        const int changed = __ldcg(&flags[rounds & 3]);
        ++rounds;
        if (!changed) break;
    }
    if (crank != 0) return;        // the wrap-up is one cheap pass: CTA 0 finishes alone (no cluster barrier after this point)
    // final holders: the last (highest) query that wrote each keypoint
    for (int i = tid; i < n; i += 1024) assign[i] = -1;
    if (tid == 0) s_count = 0;
    if (tid < HISTO) { s_hist[tid] = 0; s_keep[tid] = 1; }
    __syncthreads();
    int local = 0;
    for (int q = tid; q < nq; q += 1024) {
        const int t = __ldcg(&target[q]);
        if (t < 0) continue;
        ++local;
        atomicMax(&assign[t], q);
        if (MODE == 1 && check_ori) {
            const float factor = HISTO / 360.0f;
            float rot = reinterpret_cast<const plvs_last_query*>(queries)[q].angle - keys[t].angle;
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO) bin = 0;
            atomicAdd(&s_hist[bin], 1);
        }
    }
    atomicAdd(&s_count, local);
    __syncthreads();
    if (MODE == 1 && check_ori) {
        if (tid == 0) {     // ComputeThreeMaxima (src/ORBmatcher.cc:2123-2164)
            int m1 = 0, m2 = 0, m3 = 0, i1 = -1, i2 = -1, i3 = -1;
            for (int i = 0; i < HISTO; ++i) {
                const int s = s_hist[i];
                if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
                else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
                else if (s > m3) { m3 = s; i3 = i; }
            }
            if ((float)m2 < 0.1f * (float)m1) { i2 = -1; i3 = -1; }
            else if ((float)m3 < 0.1f * (float)m1) { i3 = -1; }
            for (int i = 0; i < HISTO; ++i) s_keep[i] = (i == i1 || i == i2 || i == i3);
        }
        __syncthreads();
        int dropped = 0;
        for (int q = tid; q < nq; q += 1024) {
            const int t = __ldcg(&target[q]);
            if (t < 0) continue;
            const float factor = HISTO / 360.0f;
            float rot = reinterpret_cast<const plvs_last_query*>(queries)[q].angle - keys[t].angle;
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO) bin = 0;
            if (!s_keep[bin]) { assign[t] = -1; ++dropped; }       // the reference nulls the slot whoever holds it now
        }
        atomicSub(&s_count, dropped);
    }
    __syncthreads();
    for (int i = tid; i < n; i += 1024) assign_out[i] = assign[i];
    if (tid == 0) { result[0] = s_count; result[1] = rounds; }
    stamp();
}

// ---------------------------------------------------------------------------------------------
// Phase B on ONE CTA (the default): the same Jacobi fixed point as k_resolve, evaluated incrementally.  Round 0 (nobody has claimed anything
// yet) comes out of k_candidates for free.  From then on a round is: rebuild the claim table from the current targets (one atomicMin per
// query), then re-evaluate ONLY the queries whose decision can have changed.  A query's decision is a function of which of its candidates are
// blocked by a lower query, and only of the candidates up to (in list-key order) its second-best unblocked one (map search: the ratio test
// reads the second best) or its best unblocked one (last-frame search): anything behind them can flip without consequence.  So every
// evaluation leaves a WATCH SET -- those few candidates with the blocked / free state it saw (pre-claimed keypoints never change and are left
// out) -- and a round just compares the watch set with the new table (a handful of shared-memory reads per query); only on a difference is
// the candidate list walked again (from L2: a few dozen queries per round).  Queries whose watch set does not fit kWatch entries are walked
// every round.  The shared memory holds two claim tables and the per-query records -- no candidate lists -- so the kernel needs no large
// opt-in allocation and starts on any SM with room for one CTA.  By induction over the rounds every query holds exactly what a full
// re-evaluation would give it, so the fixed point is the one k_resolve reaches: the sequential result.
// result: [0] matches, [1] rounds, [2] list walks after round 0, [3] largest raw candidate count if it exceeded `cap`.
// ---------------------------------------------------------------------------------------------
constexpr int kWatch = 6;
struct WatchRec { uint16_t kp[kWatch]; uint8_t n, blocked; uint16_t m; };        // n == 255: re-evaluate every round; m: length of the candidate list
constexpr int kWalkBatch = 4;

// One evaluation of query q against the claim table `cur` by one warp: best / second-best free candidate -> target, and the new watch set
// (every candidate up to the last one the decision read, except pre-claimed ones, with the state just seen).  Lists of up to 32 entries -- the
// common case -- are handled without loops: the lanes hold the entries (e_first), two REDUX give the two smallest keys, shuffles their
// records, one ballot the watch set, one REDUX.OR its blocked bits.  Returns (lane 0) whether the target changed.
template <int MODE>
__device__ __forceinline__ bool reevaluate(int q, int m, uint32_t e_first, const uint32_t* __restrict__ row, const int* __restrict__ cur,
                                           WatchRec* __restrict__ s_watch, int* __restrict__ s_target, float nn_ratio, int th_high, int lane32)
{
    uint16_t* wkp = reinterpret_cast<uint16_t*>(&s_watch[q]);
    int t, cnt; uint32_t bits;
    if (m <= 32) {
        const bool valid = lane32 < m;
        const int idx = cand_idx(e_first);
        const int c = valid ? cur[idx] : -1;
        const uint32_t kk = ((uint32_t)cand_dist(e_first) << 16) | (uint32_t)lane32;
        const uint32_t key = (valid && !(c < q)) ? kk : 0xffffffffu;
        const uint32_t k1 = __reduce_min_sync(0xffffffffu, key);
        const uint32_t k2 = __reduce_min_sync(0xffffffffu, key == k1 ? 0xffffffffu : key);
        const uint32_t e1 = __shfl_sync(0xffffffffu, e_first, (int)(k1 & 31u)), e2 = __shfl_sync(0xffffffffu, e_first, (int)(k2 & 31u));
        t = decide_target<MODE>(k1, k2, k1 != 0xffffffffu ? e1 : 0u, k2 != 0xffffffffu ? e2 : 0u, nn_ratio, th_high);
        const uint32_t limit = MODE == 0 ? k2 : k1;              // 0xffffffff (not enough free candidates): the whole list matters
        const bool rel = valid && kk <= limit && c != -1;
        const uint32_t relmask = __ballot_sync(0xffffffffu, rel);
        const int pos = __popc(relmask & ((1u << lane32) - 1));
        cnt = __popc(relmask);
        if (rel && pos < kWatch) wkp[pos] = (uint16_t)idx;
        bits = __reduce_or_sync(0xffffffffu, (rel && pos < kWatch && c < q) ? (1u << pos) : 0u);
    } else {
        uint32_t k1 = 0xffffffffu, k2 = 0xffffffffu;
        for (int k0 = 0; k0 < m; k0 += 32) {
            const int k = k0 + lane32;
            uint32_t key = 0xffffffffu;
            if (k < m) { const uint32_t e = k0 == 0 ? e_first : row[k]; if (!(cur[cand_idx(e)] < q)) key = ((uint32_t)cand_dist(e) << 16) | (uint32_t)k; }
            const uint32_t s1 = __reduce_min_sync(0xffffffffu, key);
            const uint32_t s2 = __reduce_min_sync(0xffffffffu, key == s1 ? 0xffffffffu : key);
            const uint32_t lo = min(k1, s1), hi = max(k1, s1);
            k2 = min(hi, min(k2, s2));
            k1 = lo;
        }
        const uint32_t e1 = k1 != 0xffffffffu ? row[k1 & 0xffffu] : 0u, e2 = k2 != 0xffffffffu ? row[k2 & 0xffffu] : 0u;
        t = decide_target<MODE>(k1, k2, e1, e2, nn_ratio, th_high);
        const uint32_t limit = MODE == 0 ? k2 : k1;
        cnt = 0; bits = 0;
        for (int k0 = 0; k0 < m && cnt <= kWatch; k0 += 32) {
            const int k = k0 + lane32;
            bool rel = false, blk = false; uint32_t e = 0;
            if (k < m) {
                e = k0 == 0 ? e_first : row[k];
                const int c = cur[cand_idx(e)];
                rel = ((((uint32_t)cand_dist(e) << 16) | (uint32_t)k) <= limit) && c != -1;
                blk = c < q;
            }
            const uint32_t relmask = __ballot_sync(0xffffffffu, rel);
            const int pos = cnt + __popc(relmask & ((1u << lane32) - 1));
            if (rel && pos < kWatch) wkp[pos] = (uint16_t)cand_idx(e);
            bits |= __reduce_or_sync(0xffffffffu, (rel && pos < kWatch && blk) ? (1u << pos) : 0u);
            cnt += __popc(relmask);
        }
    }
    bool changed = false;
    __syncwarp();              // every lane has read this query's record (its length, when the batch was fetched) before lane 0 rewrites the word
    if (lane32 == 0) {
        reinterpret_cast<uint32_t*>(&s_watch[q])[3] = (cnt > kWatch ? 255u : (uint32_t)cnt) | (bits << 8) | ((uint32_t)m << 16);
        if (t != s_target[q]) { s_target[q] = t; changed = true; }
    }
    return changed;
}

static_assert(sizeof(WatchRec) == 16, "WatchRec layout");

template <int MODE>
__global__ void __launch_bounds__(1024)
k_resolve_cta(const uint32_t* __restrict__ cand, const int* __restrict__ cand_n, int cap, const void* __restrict__ queries, int nq,
              const plvs_keypoint* __restrict__ keys, int n, const uint8_t* __restrict__ claimed_in, float nn_ratio, int check_ori, int th_high,
              const Round0* __restrict__ round0, int32_t* __restrict__ assign_out /*n, mapped host*/, int* __restrict__ result, int* __restrict__ max_count)
{
    PLVS_DYN_SMEM_ALIGNED(uint32_t, s_dyn, 16);
    __shared__ int s_count, s_hist[HISTO], s_keep[HISTO], s_walks, s_nwalk;
    const int tid = threadIdx.x, lane32 = tid & 31, wid = tid >> 5, nthr = blockDim.x;
    const int INF = 0x7fffffff;
    WatchRec* s_watch = reinterpret_cast<WatchRec*>(s_dyn);                    // nq (16-byte records first: alignment)
    int* tab0 = reinterpret_cast<int*>(s_dyn + 4 * (size_t)nq);               // n
    int* tab1 = tab0 + n;                                                      // n
    int* s_target = tab1 + n;                                                  // nq
    int* s_wlist = s_target + nq;                                              // nq: queries queued for a re-evaluation this round
    uint8_t* s_obs = reinterpret_cast<uint8_t*>(s_wlist + nq);                 // nq: Observations() > 0
    if (tid == 0) s_walks = 0;
    for (int i = tid; i < n; i += nthr) { const int v = (claimed_in && claimed_in[i]) ? -1 : INF; tab0[i] = v; tab1[i] = v; }
    for (int q = tid; q < nq; q += nthr) {
        const uint32_t fl = MODE == 0 ? reinterpret_cast<const plvs_mp_query*>(queries)[q].flags : reinterpret_cast<const plvs_last_query*>(queries)[q].flags;
        s_obs[q] = (fl & PLVS_Q_OBS_POSITIVE) ? 1 : 0;
    }
    PLVS_GRID_DEP_WAIT();                   // everything above read the caller's inputs only; from here on: what k_candidates wrote
    for (int q = tid; q < nq; q += nthr) {
        // round 0: the watch set is the best (and, for the map search, the second-best) candidate that was free; with fewer free candidates
        // than that, ANY candidate coming free would matter -- but at round 0 nothing is blocked except pre-claims, which never come free
        const Round0 r = round0[q];
        s_target[q] = r.target;
        WatchRec w; w.blocked = 0; w.n = (uint8_t)r.nwatch; w.m = r.m;
#pragma unroll
        for (int j = 0; j < kWatch; ++j) w.kp[j] = 0;
        w.kp[0] = r.w1; w.kp[1] = r.w2;
        s_watch[q] = w;
    }
    __syncthreads();

    // A round: (B) claim table of the current targets -- lowest query with Observations() > 0 per keypoint, -1 marks a pre-claimed keypoint --
    // into the clean table; (C) every query compares its watch set with it and queues up if something it depends on changed, while the table of
    // the previous round is wiped for the next one; (D) the queued queries are evaluated again.  Three CTA barriers per round.
    int* cur = tab0; int* nxt = tab1;       // both clean (pre-claims only) here
    int rounds = 1;                         // round 0 ran inside k_candidates
    long long t_mark = clock64(), t_b = 0, t_c = 0, t_d = 0;      // thread 0: cycles per phase (inspection: plvs_match_last_phase_cycles)
    const long long t_start = t_mark;
    for (;;) {
        for (int q = tid; q < nq; q += nthr) { const int t = s_target[q]; if (t >= 0 && s_obs[q]) atomicMin(&nxt[t], q); }
        if (tid == 0) s_nwalk = 0;
        __syncthreads();
        { const long long t = clock64(); t_b += t - t_mark; t_mark = t; }
        { int* t = cur; cur = nxt; nxt = t; }
        bool changed = false;
        for (int q0 = 0; q0 < nq; q0 += nthr) {
            const int q = q0 + tid;
            const uint4 raw = q < nq ? reinterpret_cast<const uint4*>(s_watch)[q] : make_uint4(0u, 0u, 0u, 0u);                 // kp[0..5] | n, blocked, m
            const uint32_t wn = raw.w & 0xffu, wb = (raw.w >> 8) & 0xffu;
            bool walk = wn == 255u;
            if (!walk && q < nq) {
                const uint32_t kp[kWatch] = {raw.x & 0xffffu, raw.x >> 16, raw.y & 0xffffu, raw.y >> 16, raw.z & 0xffffu, raw.z >> 16};
#pragma unroll
                for (int j = 0; j < kWatch; ++j)
                    if ((uint32_t)j < wn) walk |= ((cur[kp[j]] < q) != (((wb >> j) & 1u) != 0u));
            }
            // queue up (order irrelevant: an evaluation reads `cur` and writes its own records only); one shared-memory atomic per warp
            const uint32_t wm = __ballot_sync(0xffffffffu, walk);
            int wbase = 0;
            if (lane32 == 0 && wm) wbase = atomicAdd(&s_nwalk, __popc(wm));
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            if (walk) s_wlist[wbase + __popc(wm & ((1u << lane32) - 1))] = q;
        }
        for (int i = tid; i < n; i += nthr) nxt[i] = nxt[i] < 0 ? -1 : INF;
        __syncthreads();
        { const long long t = clock64(); t_c += t - t_mark; t_mark = t; }
        // (D) one warp per queued query: the lanes take the list entries (rows are read from L2, coalesced; the first 32 entries of up to four
        // queries are requested before the first one is evaluated, so a warp pays the L2 round trip once per four queries), warp reductions
        // give the best and second-best free candidate, ballots the new watch set
        const int nwalk = s_nwalk;
        for (int i0 = wid; i0 < nwalk; i0 += (nthr >> 5) * kWalkBatch) {
            int qv[kWalkBatch], mv[kWalkBatch]; uint32_t ev[kWalkBatch];
#pragma unroll
            for (int u = 0; u < kWalkBatch; ++u) {
                const int i = i0 + u * (nthr >> 5);
                qv[u] = i < nwalk ? s_wlist[i] : -1;
                mv[u] = qv[u] >= 0 ? (int)s_watch[qv[u]].m : 0;
                ev[u] = lane32 < mv[u] ? cand[(size_t)qv[u] * cap + lane32] : 0u;
            }
#pragma unroll 1
            for (int u = 0; u < kWalkBatch; ++u) {        // one copy of the evaluation in the code; the batch only exists to have the loads in flight together
                const int q = u == 0 ? qv[0] : u == 1 ? qv[1] : u == 2 ? qv[2] : qv[3];
                if (q < 0) break;                      // warp-uniform
                const int m = u == 0 ? mv[0] : u == 1 ? mv[1] : u == 2 ? mv[2] : mv[3];
                const uint32_t e_first = u == 0 ? ev[0] : u == 1 ? ev[1] : u == 2 ? ev[2] : ev[3];
                if (reevaluate<MODE>(q, m, e_first, cand + (size_t)q * cap, cur, s_watch, s_target, nn_ratio, th_high, lane32)) changed = true;
            }
        }
        if (tid == 0) s_walks += nwalk;
        ++rounds;
        const int go_on = __syncthreads_or(changed ? 1 : 0);
        { const long long t = clock64(); t_d += t - t_mark; t_mark = t; }
        if (!go_on) break;
    }
    // final holders: the last (highest) query that wrote each keypoint
    int* assign = nxt;
    __syncthreads();
    for (int i = tid; i < n; i += nthr) assign[i] = -1;
    if (tid == 0) s_count = 0;
    if (tid < HISTO) { s_hist[tid] = 0; s_keep[tid] = 1; }
    __syncthreads();
    int local = 0;
    for (int q = tid; q < nq; q += nthr) {
        const int t = s_target[q];
        if (t < 0) continue;
        ++local;
        atomicMax(&assign[t], q);
        if (MODE == 1 && check_ori) {
            const float factor = HISTO / 360.0f;
            float rot = reinterpret_cast<const plvs_last_query*>(queries)[q].angle - keys[t].angle;
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO) bin = 0;
            atomicAdd(&s_hist[bin], 1);
        }
    }
    atomicAdd(&s_count, local);
    __syncthreads();
    if (MODE == 1 && check_ori) {
        if (tid == 0) {     // ComputeThreeMaxima (src/ORBmatcher.cc:2123-2164)
            int m1 = 0, m2 = 0, m3 = 0, i1 = -1, i2 = -1, i3 = -1;
            for (int i = 0; i < HISTO; ++i) {
                const int sh = s_hist[i];
                if (sh > m1) { m3 = m2; m2 = m1; m1 = sh; i3 = i2; i2 = i1; i1 = i; }
                else if (sh > m2) { m3 = m2; m2 = sh; i3 = i2; i2 = i; }
                else if (sh > m3) { m3 = sh; i3 = i; }
            }
            if ((float)m2 < 0.1f * (float)m1) { i2 = -1; i3 = -1; }
            else if ((float)m3 < 0.1f * (float)m1) { i3 = -1; }
            for (int i = 0; i < HISTO; ++i) s_keep[i] = (i == i1 || i == i2 || i == i3);
        }
        __syncthreads();
        int dropped = 0;
        for (int q = tid; q < nq; q += nthr) {
            const int t = s_target[q];
            if (t < 0) continue;
            const float factor = HISTO / 360.0f;
            float rot = reinterpret_cast<const plvs_last_query*>(queries)[q].angle - keys[t].angle;
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO) bin = 0;
            if (!s_keep[bin]) { assign[t] = -1; ++dropped; }       // the reference nulls the slot whoever holds it now
        }
        atomicSub(&s_count, dropped);
    }
    __syncthreads();
    for (int i = tid; i < n; i += nthr) assign_out[i] = assign[i];
    if (tid == 0) {
        result[0] = s_count; result[1] = rounds; result[2] = s_walks; result[3] = *max_count; *max_count = 0;
        result[4] = (int)t_b; result[5] = (int)t_c; result[6] = (int)t_d; result[7] = (int)(clock64() - t_start);
    }
}

// ---------------------------------------------------------------------------------------------
// SearchForTriangulation (src/ORBmatcher.cc:1059-1208): one warp per feature of KF1 (entries of its feature
// vector, flattened); lanes sweep the KF2 features of the same vocabulary node.  Winner = smallest
// distance <= TH_LOW among the gated candidates, LAST one on ties (the `dist>bestDist` skip).
// ---------------------------------------------------------------------------------------------
struct FvDev { int n_nodes; const uint32_t* ids; const int* off; const int* feat; };

__global__ void __launch_bounds__(256)
k_triangulate(ViewDev K1, ViewDev K2, FvDev f1, FvDev f2, const uint8_t* __restrict__ has1, const uint8_t* __restrict__ has2,
              const float* __restrict__ F12, float epx, float epy, int only_stereo, int coarse, int total1, int32_t* __restrict__ match12)
{
    // one warp per entry of KF1's feature vector (flattened); its vocabulary node comes from two binary searches
    const int e = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= total1 || e >= f1.off[f1.n_nodes]) return;
    int lo = 0, hi = f1.n_nodes;                        // last node with off[a] <= e
    while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (f1.off[mid] <= e) lo = mid; else hi = mid; }
    const uint32_t id = f1.ids[lo];
    int b = -1;
    lo = 0; hi = f2.n_nodes - 1;
    while (lo <= hi) { const int mid = (lo + hi) >> 1; const uint32_t v = f2.ids[mid]; if (v == id) { b = mid; break; } if (v < id) lo = mid + 1; else hi = mid - 1; }
    if (b < 0) return;
    const int b0 = f2.off[b], b1 = f2.off[b + 1];
    {
        const int idx1 = f1.feat[e];
        if (has1[idx1]) return;
        const bool stereo1 = K1.uright && K1.uright[idx1] >= 0;
        if (only_stereo && !stereo1) return;
        const plvs_keypoint kp1 = K1.keys[idx1];
        const uint8_t* d1 = K1.desc + (size_t)idx1 * 32;
        const uint4 a0 = *reinterpret_cast<const uint4*>(d1), a1 = *reinterpret_cast<const uint4*>(d1 + 16);
        // epipolar line l = x1' F12 (src/CameraModels/Pinhole.cpp:133-136)
        const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
        const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
        const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
        const float den = la * la + lb * lb;
        int best = 0x7fffffff;     // key = dist * 65536 + (65535 - pos): min key == min dist, then max pos
        int bestIdx = -1;
        for (int p = b0 + lane; p < b1; p += 32) {
            const int idx2 = f2.feat[p];
            if (has2[idx2]) continue;
            const bool stereo2 = K2.uright && K2.uright[idx2] >= 0;
            if (only_stereo && !stereo2) continue;
            const int dist = hamming256(a0, a1, K2.desc + (size_t)idx2 * 32);
            if (dist > TH_LOW) continue;
            const plvs_keypoint kp2 = K2.keys[idx2];
            if (!stereo1 && !stereo2) {
                const float dx = epx - kp2.x, dy = epy - kp2.y;
                if (dx * dx + dy * dy < 100 * K2.scale[kp2.octave]) continue;
            }
            if (!coarse) {
                if (den == 0) continue;
                const float num = la * kp2.x + lb * kp2.y + lc;
                const float dsqr = num * num / den;
                if (!((double)dsqr < 3.84 * (double)K2.sigma2[kp2.octave])) continue;
            }
            const int key = dist * 65536 + (65535 - (p - b0));
            if (key < best) { best = key; bestIdx = idx2; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, o), oi = __shfl_xor_sync(0xffffffffu, bestIdx, o);
            if (ob < best) { best = ob; bestIdx = oi; }
        }
        if (lane == 0) match12[idx1] = bestIdx;
    }
}

__global__ void __launch_bounds__(1024)
k_tri_finish(const plvs_keypoint* __restrict__ k1, const plvs_keypoint* __restrict__ k2, int n1, int check_ori,
             int32_t* __restrict__ match12, int32_t* __restrict__ match_out, int* __restrict__ result, int reverse = 0)
{
    // reverse: rot = k2[j].angle - k1[i].angle (SearchByBoW indexes the matches by the FRAME feature, rot = kpKF - kpF)
    __shared__ int s_hist[HISTO], s_keep[HISTO], s_count;
    const int tid = threadIdx.x;
    if (tid < HISTO) { s_hist[tid] = 0; s_keep[tid] = 1; }
    if (tid == 0) s_count = 0;
    __syncthreads();
    const float factor = HISTO / 360.0f;
    int local = 0;
    for (int i = tid; i < n1; i += 1024) {
        const int j = match12[i];
        if (j < 0) continue;
        ++local;
        if (check_ori) {
            float rot = reverse ? k2[j].angle - k1[i].angle : k1[i].angle - k2[j].angle;
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO) bin = 0;
            atomicAdd(&s_hist[bin], 1);
        }
    }
    atomicAdd(&s_count, local);
    __syncthreads();
    if (check_ori) {
        if (tid == 0) {
            int m1 = 0, m2 = 0, m3 = 0, i1 = -1, i2 = -1, i3 = -1;
            for (int i = 0; i < HISTO; ++i) {
                const int s = s_hist[i];
                if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
                else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
                else if (s > m3) { m3 = s; i3 = i; }
            }
            if ((float)m2 < 0.1f * (float)m1) { i2 = -1; i3 = -1; }
            else if ((float)m3 < 0.1f * (float)m1) { i3 = -1; }
            for (int i = 0; i < HISTO; ++i) s_keep[i] = (i == i1 || i == i2 || i == i3);
        }
        __syncthreads();
        int dropped = 0;
        for (int i = tid; i < n1; i += 1024) {
            const int j = match12[i];
            if (j < 0) continue;
            float rot = reverse ? k2[j].angle - k1[i].angle : k1[i].angle - k2[j].angle;
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * factor);
            if (bin == HISTO) bin = 0;
            if (!s_keep[bin]) { match12[i] = -1; ++dropped; }
        }
        atomicSub(&s_count, dropped);
    }
    __syncthreads();
    for (int i = tid; i < n1; i += 1024) match_out[i] = match12[i];
    if (tid == 0) result[0] = s_count;
}

// ---------------------------------------------------------------------------------------------
// ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (src/ORBmatcher.cc:300-506, Nleft == -1).  The claim
// `if(vpMapPointMatches[realIdxF]) continue` only couples keyframe features of the SAME vocabulary node (a frame feature
// sits in exactly one node), so nodes are independent: one warp per keyframe node, the keyframe features of the node in
// sequence, lanes over the frame features of the node.  best = first index with the smallest distance, second = second
// smallest of the multiset (what the if / else-if pair computes, order-independent).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_bow(ViewDev K, ViewDev F, FvDev fK, FvDev fF, const uint8_t* __restrict__ has_mp, float ratio, int32_t* match_f,
      const uint8_t* __restrict__ has2 /*KF-KF overload: set-2 features must carry a good map point*/, int strict /*KF-KF: bestDist1 < TH_LOW*/,
      int32_t* out12 /*KF-KF: out12[idx1] = idx2*/)
{
    const int a = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (a >= fK.n_nodes) return;
    const uint32_t id = fK.ids[a];
    int b = -1, lo = 0, hi = fF.n_nodes - 1;
    while (lo <= hi) { const int mid = (lo + hi) >> 1; const uint32_t v = fF.ids[mid]; if (v == id) { b = mid; break; } if (v < id) lo = mid + 1; else hi = mid - 1; }
    if (b < 0) return;
    const int k0 = fK.off[a], k1 = fK.off[a + 1], b0 = fF.off[b], b1 = fF.off[b + 1];
    volatile int32_t* mf = match_f;
    for (int iK = k0; iK < k1; ++iK) {
        const int idxK = fK.feat[iK];
        if (!has_mp[idxK]) continue;
        const uint8_t* dK = K.desc + (size_t)idxK * 32;
        const uint4 a0 = *reinterpret_cast<const uint4*>(dK), a1 = *reinterpret_cast<const uint4*>(dK + 16);
        uint32_t bestkey = 0xffffffffu;      // dist << 16 | position in the node list
        int bi = -1, d1 = 256, d2 = 256;
        for (int p = b0 + lane; p < b1; p += 32) {
            const int idxF = fF.feat[p];
            if (mf[idxF] >= 0) continue;
            if (has2 && !has2[idxF]) continue;
            const int dist = hamming256(a0, a1, F.desc + (size_t)idxF * 32);
            const uint32_t key = ((uint32_t)dist << 16) | (uint32_t)min(p - b0, 65535);
            if (key < bestkey) { bestkey = key; bi = idxF; }
            if (dist < d1) { d2 = d1; d1 = dist; } else if (dist < d2) d2 = dist;
        }
        const int m1 = __reduce_min_sync(0xffffffffu, d1);
        const int winner = __ffs(__ballot_sync(0xffffffffu, d1 == m1)) - 1;
        const int m2 = __reduce_min_sync(0xffffffffu, lane == winner ? d2 : d1);
        const uint32_t wk = __reduce_min_sync(0xffffffffu, bestkey);
        const int src = __ffs(__ballot_sync(0xffffffffu, bestkey == wk)) - 1;
        const int idx = __shfl_sync(0xffffffffu, bi, src);
        if (lane == 0 && (strict ? m1 < TH_LOW : m1 <= TH_LOW) && (float)m1 < ratio * (float)m2) { mf[idx] = idxK; if (out12) out12[idxK] = idx; }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:428-455), batched over map points: all pairwise Hamming distances
// of a point's observed descriptors, per descriptor the median `sorted[0.5*(N-1)]` of its row, the first descriptor with the
// smallest median wins (`median < BestMedian`).  One warp per map point, distance matrix in shared memory (u16).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
k_distinctive(const uint8_t* __restrict__ desc, const int32_t* __restrict__ off, int32_t* __restrict__ best)
{
    PLVS_DYN_SMEM(uint16_t, s_d);
    const int pt = blockIdx.x, lane = threadIdx.x;
    const int o = off[pt], n = off[pt + 1] - o;
    if (n <= 0) { if (lane == 0) best[pt] = -1; return; }
    for (int e = lane; e < n * n; e += 32) {
        const int i = e / n, j = e - i * n;
        if (j < i) continue;
        int d = 0;
        if (j > i) {
            const uint4* a = reinterpret_cast<const uint4*>(desc + (size_t)(o + i) * 32);
            d = hamming256(a[0], a[1], desc + (size_t)(o + j) * 32);
        }
        s_d[i * n + j] = (uint16_t)d; s_d[j * n + i] = (uint16_t)d;
    }
    __syncwarp();
    const int k = (int)(0.5 * (n - 1));
    uint32_t key = 0xffffffffu;                       // median << 16 | index
    for (int i = lane; i < n; i += 32) {
        const uint16_t* row = s_d + i * n;
        int med = 0;
        for (int j = 0; j < n; ++j) {
            const int v = row[j];
            int less = 0, eq_before = 0;
            for (int m = 0; m < n; ++m) { const int u = row[m]; less += u < v; eq_before += (u == v) & (m < j); }
            if (less + eq_before == k) { med = v; break; }
        }
        key = min(key, ((uint32_t)med << 16) | (uint32_t)i);
    }
    key = __reduce_min_sync(0xffffffffu, key);
    if (lane == 0) best[pt] = (int)(key & 0xffffu);
}

#include "match_frustum.cuh"

#include "match_init.cuh"

__global__ void k_take_max(int* max_count, int* result) { result[3] = *max_count; *max_count = 0; }
__global__ void k_fill_i32(int32_t* p, int n, int32_t v) { const int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

}  // namespace

// ---------------------------------------------------------------------------------------------
// a17  Frame::ComputeStereoMatches (src/Frame.cc:1780-1983).  One warp per left keypoint:
//  (1) lanes sweep the right keypoints: row band (vRowIndices membership), octave window, disparity range,
//      Hamming; winner = smallest distance, first index on ties (`dist<bestDist` over ascending iR);
//  (2) if < (TH_HIGH+TH_LOW)/2: 11 SAD windows of 11x11 px on the unblurred level of the LEFT keypoint,
//      lanes over the 121 pixels, integer sums (exact), first minimum wins; parabola fit in fp32 as written;
//  (3) a single-CTA pass selects the median SAD (radix select on the 15-bit values) and applies the
//      1.5*1.4*median cut.  Columns left of the level are read through BORDER_REFLECT_101 like the
//      reference's bordered pyramid buffers.
// ---------------------------------------------------------------------------------------------
struct PyrDev { const uint8_t* data[PLVS_MAX_LEVELS]; int w[PLVS_MAX_LEVELS], h[PLVS_MAX_LEVELS], pitch[PLVS_MAX_LEVELS]; };
struct StereoScales { float scale[PLVS_MAX_LEVELS], inv[PLVS_MAX_LEVELS]; };

__device__ __forceinline__ int refl101(int p, int n) { if (p < 0) p = -p; if (p >= n) p = 2 * n - 2 - p; return p; }

__global__ void __launch_bounds__(256)
k_stereo_rows(const plvs_keypoint* __restrict__ kl, const uint8_t* __restrict__ dl, int n,
              const plvs_keypoint* __restrict__ kr, const uint8_t* __restrict__ dr, int nr,
              PyrDev PL, PyrDev PR, StereoScales S, int n_rows, float mb, float mbf,
              float* __restrict__ uright, float* __restrict__ depth, int* __restrict__ sad /*n: best SAD or -1*/)
{
    const int iL = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (iL >= n) return;
    const plvs_keypoint kpL = kl[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const int row = (int)vL;
    const float minD = 0.f, maxD = mbf / mb;
    const float minU = uL - maxD, maxU = uL - minD;
    float out_u = -1.0f, out_d = -1.0f; int out_sad = -1;
    if (!(maxU < 0) && row >= 0 && row < n_rows) {
        const uint4 a0 = *reinterpret_cast<const uint4*>(dl + (size_t)iL * 32), a1 = *reinterpret_cast<const uint4*>(dl + (size_t)iL * 32 + 16);
        int best = 0x7fffffff;          // dist * 65536 + iR: lexicographic (dist, iR)
        for (int iR = lane; iR < nr; iR += 32) {
            const plvs_keypoint r = kr[iR];
            const float rr = 2.0f * S.scale[r.octave];
            if (row < (int)floorf(r.y - rr) || row > (int)ceilf(r.y + rr)) continue;      // vRowIndices[row] membership
            if (r.octave < levelL - 1 || r.octave > levelL + 1) continue;
            if (!(r.x >= minU && r.x <= maxU)) continue;
            const int dist = hamming256(a0, a1, dr + (size_t)iR * 32);
            if (dist < TH_HIGH) best = min(best, dist * 65536 + iR);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (best != 0x7fffffff && (best >> 16) < (TH_HIGH + TH_LOW) / 2) {
            const int bestIdxR = best & 0xffff;
            const float uR0 = kr[bestIdxR].x;
            const float sf = S.inv[levelL];
            const int su = (int)roundf(kpL.x * sf), sv = (int)roundf(kpL.y * sf), sr = (int)roundf(uR0 * sf);
            const int w = 5, L = 5;
            const uint8_t* IL = PL.data[levelL]; const uint8_t* IR = PR.data[levelL];
            const int wl = PL.w[levelL], hl = PL.h[levelL], pl = PL.pitch[levelL], wr = PR.w[levelL], hr = PR.h[levelL], pr = PR.pitch[levelL];
            const float iniu = (float)sr + L - w, endu = (float)sr + L + w + 1;
            if (!(iniu < 0 || endu >= (float)wr)) {
                int vl[4], py[4], px[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int p = lane + 32 * k;
                    py[k] = p / 11 - w; px[k] = p % 11 - w;
                    vl[k] = p < 121 ? IL[(size_t)refl101(sv + py[k], hl) * pl + refl101(su + px[k], wl)] : 0;
                }
                float vd[11];
                int bestD = 0x7fffffff, bestinc = 0;
#pragma unroll
                for (int inc = -L; inc <= L; ++inc) {
                    int acc = 0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int p = lane + 32 * k;
                        if (p < 121) acc += abs(vl[k] - (int)IR[(size_t)refl101(sv + py[k], hr) * pr + refl101(sr + inc + px[k], wr)]);
                    }
#pragma unroll
                    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    const float dist = (float)acc;
                    if (dist < (float)bestD) { bestD = (int)dist; bestinc = inc; }
                    vd[L + inc] = dist;
                }
                if (!(bestinc == -L || bestinc == L)) {
                    float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
                    for (int k = 0; k < 11; ++k) { if (k == L + bestinc - 1) d1 = vd[k]; if (k == L + bestinc) d2 = vd[k]; if (k == L + bestinc + 1) d3 = vd[k]; }
                    const float deltaR = (d1 - d3) / (2.0f * (d1 + d3 - 2.0f * d2));
                    if (!(deltaR < -1 || deltaR > 1)) {
                        float bestuR = S.scale[levelL] * ((float)sr + (float)bestinc + deltaR);
                        float disparity = uL - bestuR;
                        if (disparity >= minD && disparity < maxD) {
                            if (disparity <= 0) { disparity = 0.01f; bestuR = uL - 0.01f; }
                            out_d = mbf / disparity; out_u = bestuR; out_sad = bestD;
                        }
                    }
                }
            }
        }
    }
    if (lane == 0) { uright[iL] = out_u; depth[iL] = out_d; sad[iL] = out_sad; }
}

__global__ void __launch_bounds__(1024)
k_stereo_median(int n, const int* __restrict__ sad, float* __restrict__ uright, float* __restrict__ depth, float* __restrict__ uright_out,
                float* __restrict__ depth_out, int* __restrict__ result)
{
    __shared__ int s_cnt, s_total, s_prefix, s_rank;
    const int tid = threadIdx.x;
    if (tid == 0) { s_total = 0; }
    __syncthreads();
    int local = 0;
    for (int i = tid; i < n; i += 1024) local += sad[i] >= 0;
    atomicAdd(&s_total, local);
    __syncthreads();
    const int total = s_total;
    int kept = 0;
    if (total > 0) {
        // k-th smallest (k = total/2, 0-based) by radix select from the most significant of 15 bits
        if (tid == 0) { s_prefix = 0; s_rank = total / 2; }
        __syncthreads();
        for (int bit = 14; bit >= 0; --bit) {
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            const int prefix = s_prefix, mask_hi = ~((1 << (bit + 1)) - 1) & 0x7fff;
            int c = 0;
            for (int i = tid; i < n; i += 1024) { const int v = sad[i]; if (v >= 0 && (v & mask_hi) == prefix && !(v & (1 << bit))) ++c; }
            atomicAdd(&s_cnt, c);
            __syncthreads();
            if (tid == 0) { if (s_rank >= s_cnt) { s_rank -= s_cnt; s_prefix |= 1 << bit; } }
            __syncthreads();
        }
        const float median = (float)s_prefix;
        const float thDist = 1.5f * 1.4f * median;
        for (int i = tid; i < n; i += 1024) {
            const int v = sad[i];
            if (v >= 0 && !((float)v < thDist)) { uright[i] = -1.f; depth[i] = -1.f; }
            else if (v >= 0) ++kept;
        }
    }
    if (tid == 0) s_cnt = 0;
    __syncthreads();
    atomicAdd(&s_cnt, kept);
    __syncthreads();
    for (int i = tid; i < n; i += 1024) { uright_out[i] = uright[i]; depth_out[i] = depth[i]; }
    if (tid == 0) result[0] = s_cnt;
}

struct plvs_match {
    int device = 0;
    cudaStream_t stream = nullptr;
    // device staging for host-resident views (two frames) and queries
    DevBuf<plvs_keypoint> d_keys[2];
    DevBuf<uint8_t> d_desc[2], d_has[2], d_claimed;
    DevBuf<float> d_uright[2], d_f12;
    DevBuf<uint8_t> d_query, d_mpts, d_res_q, d_res_raw;
    DevBuf<int32_t> d_res_src;
    PinBuf<int32_t> p_res_src;
    PinBuf<uint8_t> p_frustum;
    int res_n = -1, res_total = 0;    // in-view / all queries left on the device by plvs_match_in_frustum
    bool res_compacted = false;
    DevBuf<long long> d_trace;        // PLVS_RESOLVE_TRACE=1: phase stamps of k_resolve (development)
    DevBuf<int> d_cell_start, d_sorted, d_kp_cell, d_cand_n, d_claim_a, d_claim_b, d_target, d_state;
    DevBuf<uint32_t> d_cand;
    DevBuf<uint32_t> d_fv_ids[2];
    DevBuf<int> d_fv_off[2], d_fv_feat[2];
    DevBuf<int32_t> d_assign;
    DevBuf<float> d_su, d_sd;
    PinBuf<float> p_su, p_sd;
    PinBuf<int32_t> p_assign;
    PinBuf<int> p_result, p_cand_n;
    int cap = 128;
    DevBuf<uint8_t> d_stage; PinBuf<uint8_t> p_stage;     // one packed H2D per projection search
    bool state_zeroed = false;
    int last_phase[4] = {0, 0, 0, 0};                    // SM cycles of the last one-CTA resolve: claim table, watch-set compare, re-evaluations, rounds + wrap-up
    bool use_pdl = true;                                 // k_resolve_cta as a programmatic dependent launch behind k_candidates (PLVS_MATCH_PDL=0: plain)
    DevBuf<unsigned long long> d_line_best;              // plvs_line_knn2: the two smallest keys per query
    DevBuf<int> d_line_rank; bool line_rank_ready = false;
    DevBuf<uint8_t> d_line_u8; DevBuf<float> d_line_f32;
    int resolve_threads = 1024;                          // CTA size of k_resolve_cta (PLVS_MATCH_RESOLVE_THREADS: 256 / 512 / 1024)
    DevBuf<Round0> d_round0;                              // per query: what round 0 of the claim resolution leaves (written by k_candidates)
    int last_walks = 0;                                   // list re-evaluations of the last search after round 0 (statistics)
    int last_rounds = 0, last_launches = 0;
    uint64_t grid_key = 0; int grid_n = -1;
    KernelTimer timer;
    std::mutex mu;
};

namespace {

int stage_view(plvs_match* h, int slot, const plvs_frame_view* v, ViewDev* out)
{
    if (!v || v->n < 0 || v->n > 65535 || (v->n && (!v->keys || !v->desc))) { set_error("bad frame view"); return PLVS_EINVAL; }
    if (v->nlevels < 1 || v->nlevels > PLVS_MAX_LEVELS) { set_error("bad nlevels"); return PLVS_EINVAL; }
    out->n = v->n;
    out->gp = GridParams{v->min_x, v->min_y, v->max_x, v->max_y, v->grid_inv_w, v->grid_inv_h};
    out->bf = v->bf;
    for (int i = 0; i < PLVS_MAX_LEVELS; ++i) { out->scale[i] = v->scale_factors[i]; out->sigma2[i] = v->level_sigma2[i]; }
    if (v->on_device) {
        out->keys = v->keys; out->desc = v->desc; out->uright = v->uright;
        if (v->uright && (v->on_device & PLVS_VIEW_URIGHT_ON_HOST)) {
            // keypoints / descriptors stayed on the device after extraction, mvuRight was computed by the caller on the host
            // (Frame::ComputeStereoFromRGBD, outside the hot path): 4 bytes per keypoint are staged per call
            int rc;
            if ((rc = h->d_uright[slot].alloc((size_t)std::max(v->n, 1)))) return rc;
            PLVS_CUDA(cudaMemcpyAsync(h->d_uright[slot].p, v->uright, (size_t)v->n * 4, cudaMemcpyHostToDevice, h->stream));
            out->uright = h->d_uright[slot].p;
        }
        return PLVS_OK;
    }
    int rc;
    const size_t n = (size_t)std::max(v->n, 1);
    if ((rc = h->d_keys[slot].alloc(n))) return rc;
    if ((rc = h->d_desc[slot].alloc(n * 32))) return rc;
    PLVS_CUDA(cudaMemcpyAsync(h->d_keys[slot].p, v->keys, (size_t)v->n * sizeof(plvs_keypoint), cudaMemcpyHostToDevice, h->stream));
    PLVS_CUDA(cudaMemcpyAsync(h->d_desc[slot].p, v->desc, (size_t)v->n * 32, cudaMemcpyHostToDevice, h->stream));
    out->keys = h->d_keys[slot].p; out->desc = h->d_desc[slot].p; out->uright = nullptr;
    if (v->uright) {
        if ((rc = h->d_uright[slot].alloc(n))) return rc;
        PLVS_CUDA(cudaMemcpyAsync(h->d_uright[slot].p, v->uright, (size_t)v->n * 4, cudaMemcpyHostToDevice, h->stream));
        out->uright = h->d_uright[slot].p;
    }
    return PLVS_OK;
}

// the dynamic shared memory the one-CTA resolve may use (the opt-in limit of sm_100a is 227 KiB; 200 KiB leaves the static part and some air)
constexpr size_t kResolveSmemMax = 200 * 1024;

template <int MODE>
int run_projection(plvs_match* h, const plvs_frame_view* F, const void* q, size_t qsize, int nq, float th, float nn_ratio,
                   int far_points, float th_far, int forward, int backward, int check_ori,
                   const uint8_t* claimed_in, int32_t* assign, int* nmatches, int th_high = TH_HIGH, bool q_on_device = false)
{
    if (!h || !F || !assign || !nmatches || nq < 0 || (nq && !q)) { set_error("null/invalid argument"); return PLVS_EINVAL; }
    std::unique_lock<std::mutex> lock(h->mu, std::defer_lock);
    if (!q_on_device) lock.lock();            // the resident entry point already holds the handle's mutex
    PLVS_CUDA(cudaSetDevice(h->device));
    const int n = F->n;
    cudaStream_t st = h->stream;
    int rc;
    // Everything the caller holds on the host for this search -- the query records, the pre-claim flags and (when the keypoints stayed
    // on the device after extraction) mvuRight -- travels in ONE copy: packed into the handle's pinned slab (the caller's containers are
    // pageable std::vectors in PLVS), then a single asynchronous H2D.
    plvs_frame_view Fv = *F;
    if (n < 0 || n > 65535 || (n && (!F->keys || !F->desc))) { set_error("bad frame view"); return PLVS_EINVAL; }
    const bool frame_host = !F->on_device && n > 0;
    const bool ur_host = n > 0 && F->uright && (!F->on_device || (F->on_device & PLVS_VIEW_URIGHT_ON_HOST));
    struct Part { const void* src; size_t bytes, off; } parts[5] = {
        {q_on_device ? nullptr : q, q_on_device ? 0 : qsize * (size_t)nq, 0}, {frame_host ? F->keys : nullptr, frame_host ? (size_t)n * sizeof(plvs_keypoint) : 0, 0},
        {frame_host ? F->desc : nullptr, frame_host ? (size_t)n * 32 : 0, 0}, {ur_host ? F->uright : nullptr, ur_host ? (size_t)n * 4 : 0, 0},
        {claimed_in, claimed_in ? (size_t)n : 0, 0}};
    size_t stage_bytes = 0;
    for (Part& pt : parts) { pt.off = stage_bytes; stage_bytes = align_up(stage_bytes + pt.bytes, 16); }
    const void* dq = q;                        // device-resident queries (plvs_match_in_frustum) are used where they are
    const uint8_t* d_claimed = nullptr;
    if (stage_bytes && n > 0 && nq > 0) {
        if ((rc = h->p_stage.alloc(stage_bytes)) || (rc = h->d_stage.alloc(stage_bytes))) return rc;
        for (const Part& pt : parts) if (pt.bytes) std::memcpy(h->p_stage.h + pt.off, pt.src, pt.bytes);
        PLVS_CUDA(cudaMemcpyAsync(h->d_stage.p, h->p_stage.h, stage_bytes, cudaMemcpyHostToDevice, st));
        if (parts[0].bytes) dq = h->d_stage.p + parts[0].off;
        if (frame_host) { Fv.keys = reinterpret_cast<const plvs_keypoint*>(h->d_stage.p + parts[1].off); Fv.desc = h->d_stage.p + parts[2].off; Fv.on_device = 1; Fv.cache_key = 0; }
        if (ur_host) { Fv.uright = reinterpret_cast<const float*>(h->d_stage.p + parts[3].off); Fv.on_device = 1; }
        else if (frame_host) Fv.uright = nullptr;
        if (parts[4].bytes) d_claimed = h->d_stage.p + parts[4].off;
    }
    ViewDev V;
    rc = stage_view(h, 0, &Fv, &V);
    if (rc) return rc;
    *nmatches = 0;
    for (int i = 0; i < n; ++i) assign[i] = -1;
    if (n == 0 || nq == 0) return PLVS_OK;
    if ((rc = h->d_cell_start.alloc(GRID_CELLS + 1)) || (rc = h->d_sorted.alloc(n)) || (rc = h->d_kp_cell.alloc(n)) ||
        (rc = h->d_cand_n.alloc(nq)) || (rc = h->d_state.alloc(16)) || (rc = h->p_assign.alloc(n)) || (rc = h->p_result.alloc(8)))
        return rc;
    if (!h->state_zeroed) { PLVS_CUDA(cudaMemsetAsync(h->d_state.p, 0, 16 * sizeof(int), st)); h->state_zeroed = true; }
    int launches = 0;
    // the frame's grid: handed in with the view (built at frame construction, plvs_orb_set_frame_grid), else built here once per cache_key
    const int* g_start = h->d_cell_start.p; const int* g_sorted = h->d_sorted.p;
    if (F->on_device && F->grid_cell_start && F->grid_sorted) { g_start = F->grid_cell_start; g_sorted = F->grid_sorted; }
    else if (!(F->cache_key != 0 && F->cache_key == h->grid_key && n == h->grid_n)) {
        h->timer.begin(PLVS_MATCH_K_GRID, st);
        k_build_grid<<<1, 1024, 0, st>>>(V.keys, n, V.gp, h->d_cell_start.p, h->d_sorted.p, h->d_kp_cell.p);
        h->timer.end(st);
        ++launches;
        h->grid_key = F->cache_key; h->grid_n = n;
    }
    // one CTA holds the claim tables, the targets and (if they fit) the compacted lists; past that the 8-CTA cluster kernel takes over
    const size_t cta_smem = (size_t)16 * nq + (size_t)8 * n + (size_t)8 * nq + align_up((size_t)nq, 16);      // watch records, two claim tables, targets, walk queue, flags
    // development / test knob, read per call: PLVS_MATCH_RESOLVE=cluster forces the 8-CTA cluster kernel
    const char* e_res = std::getenv("PLVS_MATCH_RESOLVE");
    const bool force_cluster = e_res && std::strcmp(e_res, "cluster") == 0;
    const bool one_cta = !force_cluster && cta_smem <= kResolveSmemMax;
    if ((rc = h->d_round0.alloc((size_t)nq))) return rc;
    for (;;) {
        if ((rc = h->d_cand.alloc((size_t)nq * h->cap))) return rc;
        h->timer.begin(PLVS_MATCH_K_CANDIDATES, st);
        k_candidates<MODE><<<div_up(nq, 8), 256, 0, st>>>(V, g_start, g_sorted, dq, nq, th, far_points, th_far, forward, backward,
                                                           h->d_cand.p, h->d_cand_n.p, h->cap, h->d_state.p + 8, d_claimed, nn_ratio, th_high, h->d_round0.p);
        h->timer.end(st);
        ++launches;
        h->timer.begin(PLVS_MATCH_K_RESOLVE, st);
        if (one_cta) {
#ifndef PLVS_CUDA_EMU
            if (h->use_pdl && !h->timer.on(PLVS_MATCH_K_RESOLVE)) {
                // programmatic dependent launch: the CTA is scheduled while k_candidates still runs and waits at PLVS_GRID_DEP_WAIT()
                cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(1); cfg.blockDim = dim3((unsigned)h->resolve_threads); cfg.dynamicSmemBytes = cta_smem; cfg.stream = st;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                PLVS_CUDA(cudaLaunchKernelEx(&cfg, k_resolve_cta<MODE>, (const uint32_t*)h->d_cand.p, (const int*)h->d_cand_n.p, h->cap, (const void*)dq, nq, (const plvs_keypoint*)V.keys, n,
                                             (const uint8_t*)d_claimed, nn_ratio, check_ori, th_high, (const Round0*)h->d_round0.p, (int32_t*)h->p_assign.d, (int*)h->p_result.d, (int*)(h->d_state.p + 8)));
            } else
#endif
            k_resolve_cta<MODE><<<1, h->resolve_threads, cta_smem, st>>>(h->d_cand.p, h->d_cand_n.p, h->cap, dq, nq, V.keys, n, d_claimed, nn_ratio, check_ori, th_high,
                                                          h->d_round0.p, h->p_assign.d, h->p_result.d, h->d_state.p + 8);
        } else {
            if ((rc = h->d_claim_a.alloc((size_t)3 * n)) || (rc = h->d_target.alloc(nq)) || (rc = h->d_assign.alloc(n))) return rc;
            PLVS_CUDA(cudaMemsetAsync(h->d_state.p, 0, 8 * sizeof(int), st));
            const int per_cta = div_up(nq, kResolveCtas);
            int lpq = 1;
            while (lpq < 32 && per_cta * lpq * 2 <= 1024) lpq *= 2;
            const size_t smem = ((size_t)2 * n + (size_t)per_cta * (h->cap | 1)) * sizeof(uint32_t);
            const size_t smem_small = (size_t)2 * n * sizeof(uint32_t);
            if (per_cta > 1024) { set_error("more than %d queries per search are not supported", kResolveCtas * 1024); return PLVS_EINVAL; }
            if (smem_small > kResolveSmemMax) { set_error("frames of more than %d keypoints are not supported by the claim resolution", (int)(kResolveSmemMax / 8)); return PLVS_EINVAL; }
            if (smem <= kResolveSmemMax)
                k_resolve<MODE, true><<<kResolveCtas, 1024, smem, st>>>(h->d_cand.p, h->d_cand_n.p, h->cap, dq, nq, V.keys, n, d_claimed, nn_ratio, check_ori, th_high,
                                                                        h->d_claim_a.p, h->d_target.p, h->d_state.p, h->d_assign.p, h->p_assign.d, h->p_result.d, per_cta, lpq, nullptr);
            else
                k_resolve<MODE, false><<<kResolveCtas, 1024, smem_small, st>>>(h->d_cand.p, h->d_cand_n.p, h->cap, dq, nq, V.keys, n, d_claimed, nn_ratio, check_ori, th_high,
                                                                               h->d_claim_a.p, h->d_target.p, h->d_state.p, h->d_assign.p, h->p_assign.d, h->p_result.d, per_cta, lpq, nullptr);
            k_take_max<<<1, 1, 0, st>>>(h->d_state.p + 8, h->p_result.d);
        }
        h->timer.end(st);
        ++launches;
        PLVS_CUDA(cudaGetLastError());
        PLVS_CUDA(cudaStreamSynchronize(st));
        h->timer.collect();
        h->last_walks = one_cta ? h->p_result.h[2] : -1;
        for (int i = 0; i < 4; ++i) h->last_phase[i] = one_cta ? h->p_result.h[4 + i] : 0;
        const int mx = h->p_result.h[3];
        if (mx <= h->cap) break;
        while (h->cap < mx) h->cap *= 2;       // a window held more candidates than reserved: redo with room (exactness first)
    }
    std::memcpy(assign, h->p_assign.h, (size_t)n * 4);
    count_d2h((size_t)n * 4 + 16);               // the kernel wrote assign[n] and the result record into mapped host memory
    *nmatches = h->p_result.h[0];
    h->last_rounds = h->p_result.h[1];
    h->last_launches = launches;
    return PLVS_OK;
}

int stage_fv(plvs_match* h, int slot, const plvs_featvec* f, int on_device, FvDev* out)
{
    if (!f || f->n_nodes < 0) { set_error("bad feature vector"); return PLVS_EINVAL; }
    out->n_nodes = f->n_nodes;
    if (on_device) { out->ids = f->node_ids; out->off = f->offsets; out->feat = f->features; return PLVS_OK; }
    const int total = f->n_nodes ? f->offsets[f->n_nodes] : 0;
    int rc;
    if ((rc = h->d_fv_ids[slot].alloc(std::max(f->n_nodes, 1))) || (rc = h->d_fv_off[slot].alloc(f->n_nodes + 1)) ||
        (rc = h->d_fv_feat[slot].alloc(std::max(total, 1)))) return rc;
    if (f->n_nodes) {
        PLVS_CUDA(cudaMemcpyAsync(h->d_fv_ids[slot].p, f->node_ids, (size_t)f->n_nodes * 4, cudaMemcpyHostToDevice, h->stream));
        PLVS_CUDA(cudaMemcpyAsync(h->d_fv_off[slot].p, f->offsets, (size_t)(f->n_nodes + 1) * 4, cudaMemcpyHostToDevice, h->stream));
        PLVS_CUDA(cudaMemcpyAsync(h->d_fv_feat[slot].p, f->features, (size_t)total * 4, cudaMemcpyHostToDevice, h->stream));
    }
    out->ids = h->d_fv_ids[slot].p; out->off = h->d_fv_off[slot].p; out->feat = h->d_fv_feat[slot].p;
    return PLVS_OK;
}

}  // namespace

extern "C" {

int plvs_hamming256(const uint8_t* a, const uint8_t* b)
{
    uint64_t x[4], y[4];
    std::memcpy(x, a, 32); std::memcpy(y, b, 32);
    int d = 0;
    for (int i = 0; i < 4; ++i) d += __builtin_popcountll(x[i] ^ y[i]);
    return d;
}

int plvs_match_create(int device, plvs_match** out)
{
    if (!out) return PLVS_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_error("no CUDA device: libplvs_b200 has no CPU fallback"); return PLVS_ENODEV; }
    if (device < 0 || device >= ndev) { set_error("device %d out of range", device); return PLVS_EINVAL; }
    PLVS_CUDA(cudaSetDevice(device));
    // opt-in shared memory limits are per device: raised here, once per handle, for every kernel that can ask for more than 48 KiB
    PLVS_CUDA(cudaFuncSetAttribute(k_resolve_cta<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResolveSmemMax));
    PLVS_CUDA(cudaFuncSetAttribute(k_resolve_cta<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResolveSmemMax));
    PLVS_CUDA(cudaFuncSetAttribute(k_resolve<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResolveSmemMax));
    PLVS_CUDA(cudaFuncSetAttribute(k_resolve<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResolveSmemMax));
    PLVS_CUDA(cudaFuncSetAttribute(k_resolve<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResolveSmemMax));
    PLVS_CUDA(cudaFuncSetAttribute(k_resolve<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kResolveSmemMax));
    plvs_match* h = new plvs_match();
    h->device = device; h->timer.component = 2;
    if (const char* e = std::getenv("PLVS_MATCH_PDL")) h->use_pdl = e[0] != '0';
    if (const char* e = std::getenv("PLVS_MATCH_RESOLVE_THREADS")) { const int v = std::atoi(e); if (v == 256 || v == 512 || v == 1024) h->resolve_threads = v; }
    { cudaError_t e = create_handle_stream(&h->stream, 2);
      if (e != cudaSuccess) { delete h; set_error("stream creation failed: %s", cudaGetErrorString(e)); return PLVS_ENODEV; } }
    *out = h;
    return PLVS_OK;
}

void plvs_match_destroy(plvs_match* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    delete h;
}

int plvs_match_kernel_times(plvs_match* h, float* ms, int32_t* launches, int reset)
{
    if (!h) return PLVS_EINVAL;
    std::lock_guard<std::mutex> lock(h->mu);
    for (int i = 0; i < KernelTimer::kSlots; ++i) { if (ms) ms[i] = h->timer.ms[i]; if (launches) launches[i] = h->timer.count[i]; }
    if (reset) h->timer.reset();
    return PLVS_OK;
}

int plvs_match_last_walks(const plvs_match* h) { return h ? h->last_walks : -1; }

int plvs_match_last_phase_cycles(const plvs_match* h, int32_t out[4])
{
    if (!h || !out) return PLVS_EINVAL;
    for (int i = 0; i < 4; ++i) out[i] = h->last_phase[i];
    return PLVS_OK;
}

int plvs_match_last_stats(const plvs_match* h, int* rounds, int* kernel_launches)
{
    if (!h) return PLVS_EINVAL;
    if (rounds) *rounds = h->last_rounds;
    if (kernel_launches) *kernel_launches = h->last_launches;
    return PLVS_OK;
}

int plvs_match_projection_map(plvs_match* h, const plvs_frame_view* F, const plvs_mp_query* q, int nq, float th, float nn_ratio,
                              int far_points, float th_far, const uint8_t* claimed_in, int32_t* assign, int* nmatches)
{
    return run_projection<0>(h, F, q, sizeof(plvs_mp_query), nq, th, nn_ratio, far_points, th_far, 0, 0, 0, claimed_in, assign, nmatches);
}

int plvs_match_projection_last(plvs_match* h, const plvs_frame_view* cur, const plvs_last_query* q, int nq, float th,
                               int forward, int backward, int check_orientation, const uint8_t* claimed_in, int32_t* assign, int* nmatches)
{
    return run_projection<1>(h, cur, q, sizeof(plvs_last_query), nq, th, 0.f, 0, 0.f, forward, backward, check_orientation, claimed_in, assign, nmatches);
}

int plvs_match_projection_reloc(plvs_match* h, const plvs_frame_view* cur, const plvs_last_query* q, int nq, float th, int orb_dist,
                                int check_orientation, const uint8_t* claimed_in, int32_t* assign, int* nmatches)
{
    // the relocalisation search is the last-frame search with: level window [l-1, l+1] around the PREDICTED level, no right-coordinate
    // gate (src/ORBmatcher.cc:2062-2076 has none), every non-null mvpMapPoints entry blocking, and ORBdist instead of TH_HIGH
    if (!cur) { set_error("null argument"); return PLVS_EINVAL; }
    if (orb_dist < 0 || orb_dist > 256) { set_error("ORBdist out of range"); return PLVS_EINVAL; }
    for (int i = 0; i < nq; ++i) if (q && !(q[i].flags & PLVS_Q_OBS_POSITIVE)) { set_error("reloc query %d: PLVS_Q_OBS_POSITIVE must be set (any claim blocks)", i); return PLVS_EINVAL; }
    plvs_frame_view v = *cur;
    v.uright = nullptr;
    return run_projection<1>(h, &v, q, sizeof(plvs_last_query), nq, th, 0.f, 0, 0.f, 0, 0, check_orientation, claimed_in, assign, nmatches, orb_dist);
}

int plvs_match_in_frustum(plvs_match* h, const plvs_frustum* fr, const plvs_map_point* pts, int n, plvs_mp_query* queries, uint8_t* in_view, int* n_in_view)
{
    if (!h || !fr || !n_in_view || n < 0 || (n && (!pts || !queries || !in_view))) { set_error("null argument"); return PLVS_EINVAL; }
    if (fr->nlevels < 1 || fr->nlevels > PLVS_MAX_LEVELS || !(fr->scale_factor > 1.0f)) { set_error("bad pyramid parameters"); return PLVS_EINVAL; }
    *n_in_view = 0;
    if (n == 0) return PLVS_OK;
    std::lock_guard<std::mutex> lock(h->mu);
    PLVS_CUDA(cudaSetDevice(h->device));
    FrustumDev D{};
    D.f = *fr;
    {   // T[k] = smallest float ratio with ceil(logf(ratio) / logf(scaleFactor)) > k: bisection over the ordered positive floats with the HOST's logf
        const float L = std::log(fr->scale_factor);
        for (int k = 0; k + 1 < fr->nlevels; ++k) {
            uint32_t lo = 0x00800000u, hi = 0x7f7fffffu;
            while (hi - lo > 1) {
                const uint32_t mid = lo + (hi - lo) / 2;
                float r; std::memcpy(&r, &mid, 4);
                if (std::ceil(std::log(r) / L) > (float)k) hi = mid; else lo = mid;
            }
            std::memcpy(&D.T[k], &hi, 4);
        }
    }
    int rc;
    cudaStream_t st = h->stream;
    const size_t qb = sizeof(plvs_mp_query) * (size_t)n;
    if ((rc = h->d_mpts.alloc(sizeof(plvs_map_point) * (size_t)n)) || (rc = h->d_query.alloc(qb + n + 8)) || (rc = h->p_frustum.alloc(qb + n + 8))) return rc;
    uint8_t* d_q = h->d_query.p; uint8_t* d_iv = d_q + qb; int* d_cnt = reinterpret_cast<int*>(d_q + ((qb + n + 3) & ~(size_t)3));
    PLVS_CUDA(cudaMemcpyAsync(h->d_mpts.p, pts, sizeof(plvs_map_point) * (size_t)n, cudaMemcpyHostToDevice, st));
    PLVS_CUDA(cudaMemsetAsync(d_cnt, 0, sizeof(int), st));
    k_in_frustum<<<div_up(n, 256), 256, 0, st>>>(D, reinterpret_cast<const plvs_map_point*>(h->d_mpts.p), n, reinterpret_cast<plvs_mp_query*>(d_q), d_iv, d_cnt);
    // the queries and flags stay on the device for plvs_match_projection_map_resident (which compacts them on first use)
    if ((rc = h->d_res_raw.alloc(qb + n))) return rc;
    PLVS_CUDA(cudaMemcpyAsync(h->d_res_raw.p, d_q, qb + n, cudaMemcpyDeviceToDevice, st));
    PLVS_CUDA(cudaMemcpyAsync(h->p_frustum.h, d_q, ((qb + n + 3) & ~(size_t)3) + 4, cudaMemcpyDeviceToHost, st));
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaStreamSynchronize(st));
    std::memcpy(queries, h->p_frustum.h, qb);
    std::memcpy(in_view, h->p_frustum.h + qb, (size_t)n);
    std::memcpy(n_in_view, h->p_frustum.h + ((qb + n + 3) & ~(size_t)3), sizeof(int));
    h->res_n = *n_in_view; h->res_total = n; h->res_compacted = false;
    h->last_launches = 1;
    return PLVS_OK;
}

int plvs_match_projection_map_resident(plvs_match* h, const plvs_frame_view* F, float th, float nn_ratio, int far_points, float th_far,
                                       const uint8_t* claimed_in, int32_t* assign, int* nmatches)
{
    if (!h || !F || !assign || !nmatches) { set_error("null argument"); return PLVS_EINVAL; }
    std::lock_guard<std::mutex> lock(h->mu);
    if (h->res_n < 0) { set_error("no resident queries: call plvs_match_in_frustum on this handle first"); return PLVS_ESTATE; }
    if (h->res_n == 0) { *nmatches = 0; for (int i = 0; i < F->n; ++i) assign[i] = -1; return PLVS_OK; }
    if (!h->res_compacted) {       // stable compaction of the in-view queries, once per plvs_match_in_frustum call
        PLVS_CUDA(cudaSetDevice(h->device));
        const int n = h->res_total;
        const size_t qb = sizeof(plvs_mp_query) * (size_t)n;
        int rc0;
        if ((rc0 = h->d_res_q.alloc(qb)) || (rc0 = h->d_res_src.alloc(n)) || (rc0 = h->p_res_src.alloc(n))) return rc0;
        k_compact_queries<<<1, 1024, 0, h->stream>>>(reinterpret_cast<const plvs_mp_query*>(h->d_res_raw.p), h->d_res_raw.p + qb, n,
                                                     reinterpret_cast<plvs_mp_query*>(h->d_res_q.p), h->d_res_src.p);
        PLVS_CUDA(cudaMemcpyAsync(h->p_res_src.h, h->d_res_src.p, (size_t)h->res_n * 4, cudaMemcpyDeviceToHost, h->stream));
        PLVS_CUDA(cudaGetLastError());
        PLVS_CUDA(cudaStreamSynchronize(h->stream));
        h->res_compacted = true;
    }
    const int rc = run_projection<0>(h, F, h->d_res_q.p, sizeof(plvs_mp_query), h->res_n, th, nn_ratio, far_points, th_far, 0, 0, 1, claimed_in, assign, nmatches,
                                     TH_HIGH, true);
    if (rc) return rc;
    // assign[] holds positions in the compacted list: map them back to indices of the caller's map-point array
    for (int i = 0; i < F->n; ++i) if (assign[i] >= 0) assign[i] = h->p_res_src.h[assign[i]];
    return PLVS_OK;
}

int plvs_match_projection_sim3(plvs_match* h, const plvs_frame_view* kf, const plvs_last_query* q, int nq, float th, float ratio_hamming,
                               const uint8_t* matched_in, int32_t* assign, int* nmatches)
{
    if (!kf) { set_error("null argument"); return PLVS_EINVAL; }
    for (int i = 0; i < nq; ++i) if (q && !(q[i].flags & PLVS_Q_OBS_POSITIVE)) { set_error("sim3 query %d: PLVS_Q_OBS_POSITIVE must be set (any match blocks)", i); return PLVS_EINVAL; }
    plvs_frame_view v = *kf;
    v.uright = nullptr;
    const int th_high = (int)floorf((float)TH_LOW * ratio_hamming);           // `bestDist <= TH_LOW*ratioHamming` with an integer bestDist (:606)
    return run_projection<1>(h, &v, q, sizeof(plvs_last_query), nq, th, 0.f, 0, 0.f, 1, 1, 0, matched_in, assign, nmatches, th_high);
}

static int fuse_impl(plvs_match* h, const plvs_frame_view* kf, const float* inv_level_sigma2, const plvs_fuse_query* q, int nq, float th,
                     int32_t* best_idx, int32_t* best_dist, int* nfused, bool chi2)
{
    if (!h || !kf || (chi2 && !inv_level_sigma2) || !best_idx || !best_dist || !nfused || nq < 0 || (nq && !q)) { set_error("null/invalid argument"); return PLVS_EINVAL; }
    std::lock_guard<std::mutex> lock(h->mu);
    PLVS_CUDA(cudaSetDevice(h->device));
    ViewDev V;
    int rc = stage_view(h, 0, kf, &V);
    if (rc) return rc;
    *nfused = 0;
    for (int i = 0; i < nq; ++i) { best_idx[i] = -1; best_dist[i] = 256; }
    const int n = kf->n;
    if (n == 0 || nq == 0) return PLVS_OK;
    for (int i = 0; i < nq; ++i) if (q[i].level < 0 || q[i].level >= kf->nlevels) { set_error("fuse query %d: level %d outside the pyramid", i, q[i].level); return PLVS_EINVAL; }
    cudaStream_t st = h->stream;
    if ((rc = h->d_query.alloc(sizeof(plvs_fuse_query) * (size_t)nq)) || (rc = h->d_cell_start.alloc(GRID_CELLS + 1)) || (rc = h->d_sorted.alloc(n)) ||
        (rc = h->d_kp_cell.alloc(n)) || (rc = h->d_assign.alloc((size_t)2 * nq + 1)) || (rc = h->p_assign.alloc((size_t)2 * nq + 1))) return rc;
    PLVS_CUDA(cudaMemcpyAsync(h->d_query.p, q, sizeof(plvs_fuse_query) * (size_t)nq, cudaMemcpyHostToDevice, st));
    int launches = 0;
    if (!(kf->cache_key != 0 && kf->cache_key == h->grid_key && n == h->grid_n)) {
        h->timer.begin(PLVS_MATCH_K_GRID, st);
        k_build_grid<<<1, 1024, 0, st>>>(V.keys, n, V.gp, h->d_cell_start.p, h->d_sorted.p, h->d_kp_cell.p);
        h->timer.end(st);
        ++launches;
        h->grid_key = kf->cache_key; h->grid_n = n;
    }
    FuseSigma sg{};
    for (int i = 0; i < kf->nlevels; ++i) sg.inv[i] = chi2 ? inv_level_sigma2[i] : 0.f;
    int32_t* d_idx = h->d_assign.p; int32_t* d_dist = d_idx + nq; int* d_nf = reinterpret_cast<int*>(d_dist + nq);
    PLVS_CUDA(cudaMemsetAsync(d_nf, 0, sizeof(int), st));
    h->timer.begin(PLVS_MATCH_K_FUSE, st);
    if (chi2) k_fuse<true><<<div_up(nq, 8), 256, 0, st>>>(V, sg, h->d_cell_start.p, h->d_sorted.p, reinterpret_cast<const plvs_fuse_query*>(h->d_query.p), nq, th, d_idx, d_dist, d_nf);
    else k_fuse<false><<<div_up(nq, 8), 256, 0, st>>>(V, sg, h->d_cell_start.p, h->d_sorted.p, reinterpret_cast<const plvs_fuse_query*>(h->d_query.p), nq, th, d_idx, d_dist, d_nf);
    h->timer.end(st);
    ++launches;
    PLVS_CUDA(cudaMemcpyAsync(h->p_assign.h, h->d_assign.p, ((size_t)2 * nq + 1) * 4, cudaMemcpyDeviceToHost, st));
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaStreamSynchronize(st));
    h->timer.collect();
    std::memcpy(best_idx, h->p_assign.h, (size_t)nq * 4);
    std::memcpy(best_dist, h->p_assign.h + nq, (size_t)nq * 4);
    *nfused = h->p_assign.h[2 * nq];
    h->last_launches = launches;
    return PLVS_OK;
}

int plvs_match_triangulation(plvs_match* h, const plvs_frame_view* kf1, const plvs_frame_view* kf2,
                             const plvs_featvec* fv1, const plvs_featvec* fv2, const uint8_t* has_mp1, const uint8_t* has_mp2,
                             const float F12[9], const float ep[2], int only_stereo, int coarse, int check_orientation,
                             int32_t* match12, int* nmatches)
{
    if (!h || !kf1 || !kf2 || !fv1 || !fv2 || !has_mp1 || !has_mp2 || !F12 || !ep || !match12 || !nmatches) { set_error("null argument"); return PLVS_EINVAL; }
    std::lock_guard<std::mutex> lock(h->mu);
    PLVS_CUDA(cudaSetDevice(h->device));
    ViewDev V1, V2; FvDev D1, D2;
    int rc;
    if ((rc = stage_view(h, 0, kf1, &V1)) || (rc = stage_view(h, 1, kf2, &V2))) return rc;
    if ((rc = stage_fv(h, 0, fv1, kf1->on_device, &D1)) || (rc = stage_fv(h, 1, fv2, kf2->on_device, &D2))) return rc;
    const int n1 = kf1->n, n2 = kf2->n;
    *nmatches = 0;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    if (n1 == 0 || n2 == 0 || D1.n_nodes == 0 || D2.n_nodes == 0) return PLVS_OK;
    cudaStream_t st = h->stream;
    const uint8_t *dh1 = has_mp1, *dh2 = has_mp2;
    if (!kf1->on_device) {
        if ((rc = h->d_has[0].alloc(n1))) return rc;
        PLVS_CUDA(cudaMemcpyAsync(h->d_has[0].p, has_mp1, n1, cudaMemcpyHostToDevice, st)); dh1 = h->d_has[0].p;
    }
    if (!kf2->on_device) {
        if ((rc = h->d_has[1].alloc(n2))) return rc;
        PLVS_CUDA(cudaMemcpyAsync(h->d_has[1].p, has_mp2, n2, cudaMemcpyHostToDevice, st)); dh2 = h->d_has[1].p;
    }
    if ((rc = h->d_f12.alloc(16)) || (rc = h->p_assign.alloc(n1)) || (rc = h->d_assign.alloc(n1)) || (rc = h->p_result.alloc(4))) return rc;
    PLVS_CUDA(cudaMemcpyAsync(h->d_f12.p, F12, 9 * sizeof(float), cudaMemcpyHostToDevice, st));
    h->timer.begin(PLVS_MATCH_K_TRIANGULATE, st);
    k_fill_i32<<<div_up(n1, 256), 256, 0, st>>>(h->d_assign.p, n1, -1);
    const int total1 = n1;         // every feature sits in at most one vocabulary node
    if (total1 > 0)
        k_triangulate<<<div_up(total1, 8), 256, 0, st>>>(V1, V2, D1, D2, dh1, dh2, h->d_f12.p, ep[0], ep[1], only_stereo, coarse, total1, h->d_assign.p);
    k_tri_finish<<<1, 1024, 0, st>>>(V1.keys, V2.keys, n1, check_orientation, h->d_assign.p, h->p_assign.d, h->p_result.d);
    h->timer.end(st);
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaStreamSynchronize(st));
    h->timer.collect();
    std::memcpy(match12, h->p_assign.h, (size_t)n1 * 4);
    count_d2h((size_t)n1 * 4 + 16);
    *nmatches = h->p_result.h[0];
    h->last_launches = 3;
    return PLVS_OK;
}

int plvs_match_fuse(plvs_match* h, const plvs_frame_view* kf, const float* inv_level_sigma2, const plvs_fuse_query* q, int nq, float th,
                    int32_t* best_idx, int32_t* best_dist, int* nfused)
{
    return fuse_impl(h, kf, inv_level_sigma2, q, nq, th, best_idx, best_dist, nfused, true);
}

int plvs_match_fuse_sim3(plvs_match* h, const plvs_frame_view* kf, const plvs_fuse_query* q, int nq, float th,
                         int32_t* best_idx, int32_t* best_dist, int* nfused)
{
    return fuse_impl(h, kf, nullptr, q, nq, th, best_idx, best_dist, nfused, false);
}

int plvs_match_bow(plvs_match* h, const plvs_frame_view* kf, const plvs_frame_view* f, const plvs_featvec* fv_kf, const plvs_featvec* fv_f,
                   const uint8_t* has_mp_kf, float nn_ratio, int check_orientation, int32_t* match_f, int* nmatches)
{
    if (!h || !kf || !f || !fv_kf || !fv_f || !has_mp_kf || !match_f || !nmatches) { set_error("null argument"); return PLVS_EINVAL; }
    std::lock_guard<std::mutex> lock(h->mu);
    PLVS_CUDA(cudaSetDevice(h->device));
    ViewDev VK, VF; FvDev DK, DF;
    int rc;
    if ((rc = stage_view(h, 0, kf, &VK)) || (rc = stage_view(h, 1, f, &VF))) return rc;
    if ((rc = stage_fv(h, 0, fv_kf, kf->on_device, &DK)) || (rc = stage_fv(h, 1, fv_f, f->on_device, &DF))) return rc;
    const int nk = kf->n, nf = f->n;
    *nmatches = 0;
    for (int i = 0; i < nf; ++i) match_f[i] = -1;
    if (nk == 0 || nf == 0 || DK.n_nodes == 0 || DF.n_nodes == 0) return PLVS_OK;
    cudaStream_t st = h->stream;
    const uint8_t* dh = has_mp_kf;
    if (!kf->on_device) {
        if ((rc = h->d_has[0].alloc(nk))) return rc;
        PLVS_CUDA(cudaMemcpyAsync(h->d_has[0].p, has_mp_kf, nk, cudaMemcpyHostToDevice, st)); dh = h->d_has[0].p;
    }
    if ((rc = h->p_assign.alloc(nf)) || (rc = h->d_assign.alloc(nf)) || (rc = h->p_result.alloc(4))) return rc;
    h->timer.begin(PLVS_MATCH_K_BOW, st);
    k_fill_i32<<<div_up(nf, 256), 256, 0, st>>>(h->d_assign.p, nf, -1);
    k_bow<<<div_up(DK.n_nodes, 8), 256, 0, st>>>(VK, VF, DK, DF, dh, nn_ratio, h->d_assign.p, nullptr, 0, nullptr);
    k_tri_finish<<<1, 1024, 0, st>>>(VF.keys, VK.keys, nf, check_orientation, h->d_assign.p, h->p_assign.d, h->p_result.d, 1);
    h->timer.end(st);
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaStreamSynchronize(st));
    h->timer.collect();
    std::memcpy(match_f, h->p_assign.h, (size_t)nf * 4);
    *nmatches = h->p_result.h[0];
    h->last_launches = 3;
    return PLVS_OK;
}

int plvs_match_bow_kf(plvs_match* h, const plvs_frame_view* kf1, const plvs_frame_view* kf2, const plvs_featvec* fv1, const plvs_featvec* fv2,
                      const uint8_t* has_mp1, const uint8_t* has_mp2, float nn_ratio, int check_orientation, int32_t* match12, int* nmatches)
{
    if (!h || !kf1 || !kf2 || !fv1 || !fv2 || !has_mp1 || !has_mp2 || !match12 || !nmatches) { set_error("null argument"); return PLVS_EINVAL; }
    std::lock_guard<std::mutex> lock(h->mu);
    PLVS_CUDA(cudaSetDevice(h->device));
    ViewDev V1, V2; FvDev D1, D2;
    int rc;
    if ((rc = stage_view(h, 0, kf1, &V1)) || (rc = stage_view(h, 1, kf2, &V2))) return rc;
    if ((rc = stage_fv(h, 0, fv1, kf1->on_device, &D1)) || (rc = stage_fv(h, 1, fv2, kf2->on_device, &D2))) return rc;
    const int n1 = kf1->n, n2 = kf2->n;
    *nmatches = 0;
    for (int i = 0; i < n1; ++i) match12[i] = -1;
    if (n1 == 0 || n2 == 0 || D1.n_nodes == 0 || D2.n_nodes == 0) return PLVS_OK;
    cudaStream_t st = h->stream;
    const uint8_t *dh1 = has_mp1, *dh2 = has_mp2;
    if (!kf1->on_device) { if ((rc = h->d_has[0].alloc(n1))) return rc; PLVS_CUDA(cudaMemcpyAsync(h->d_has[0].p, has_mp1, n1, cudaMemcpyHostToDevice, st)); dh1 = h->d_has[0].p; }
    if (!kf2->on_device) { if ((rc = h->d_has[1].alloc(n2))) return rc; PLVS_CUDA(cudaMemcpyAsync(h->d_has[1].p, has_mp2, n2, cudaMemcpyHostToDevice, st)); dh2 = h->d_has[1].p; }
    if ((rc = h->p_assign.alloc(n1)) || (rc = h->d_assign.alloc((size_t)n1 + n2)) || (rc = h->p_result.alloc(4))) return rc;
    int32_t* d_out12 = h->d_assign.p; int32_t* d_claim2 = h->d_assign.p + n1;
    h->timer.begin(PLVS_MATCH_K_BOW, st);
    k_fill_i32<<<div_up(n1 + n2, 256), 256, 0, st>>>(h->d_assign.p, n1 + n2, -1);
    k_bow<<<div_up(D1.n_nodes, 8), 256, 0, st>>>(V1, V2, D1, D2, dh1, nn_ratio, d_claim2, dh2, 1, d_out12);
    k_tri_finish<<<1, 1024, 0, st>>>(V1.keys, V2.keys, n1, check_orientation, d_out12, h->p_assign.d, h->p_result.d, 0);
    h->timer.end(st);
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaStreamSynchronize(st));
    h->timer.collect();
    std::memcpy(match12, h->p_assign.h, (size_t)n1 * 4);
    count_d2h((size_t)n1 * 4 + 16);
    *nmatches = h->p_result.h[0];
    h->last_launches = 3;
    return PLVS_OK;
}

int plvs_match_initialization(plvs_match* h, const plvs_frame_view* f1, const plvs_frame_view* f2, float* prev_matched, int window_size,
                              float nn_ratio, int check_orientation, int32_t* matches12, int* nmatches)
{
    if (!h || !f1 || !f2 || !matches12 || !nmatches || window_size < 0 || (f1->n > 0 && !prev_matched)) { set_error("null/invalid argument"); return PLVS_EINVAL; }
    std::lock_guard<std::mutex> lock(h->mu);
    PLVS_CUDA(cudaSetDevice(h->device));
    ViewDev V1, V2;
    int rc;
    if ((rc = stage_view(h, 1, f1, &V1)) || (rc = stage_view(h, 0, f2, &V2))) return rc;
    const int n1 = f1->n, n2 = f2->n;
    *nmatches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || n2 == 0) return PLVS_OK;
    cudaStream_t st = h->stream;
    if ((rc = h->d_cell_start.alloc(GRID_CELLS + 1)) || (rc = h->d_sorted.alloc(n2)) || (rc = h->d_kp_cell.alloc(n2)) || (rc = h->d_cand_n.alloc(n1)) ||
        (rc = h->d_claim_a.alloc((size_t)3 * std::max(n1, n2))) || (rc = h->d_target.alloc(n1)) || (rc = h->d_assign.alloc(n1)) || (rc = h->p_assign.alloc(n1)) ||
        (rc = h->d_su.alloc((size_t)2 * n1)) || (rc = h->p_su.alloc((size_t)2 * n1)) || (rc = h->p_result.alloc(4)) || (rc = h->p_cand_n.alloc(n1)))
        return rc;
    std::memcpy(h->p_su.h, prev_matched, (size_t)n1 * 8);
    PLVS_CUDA(cudaMemcpyAsync(h->d_su.p, h->p_su.h, (size_t)n1 * 8, cudaMemcpyHostToDevice, st));
    int launches = 0;
    if (!(f2->cache_key != 0 && f2->cache_key == h->grid_key && n2 == h->grid_n)) {
        h->timer.begin(PLVS_MATCH_K_GRID, st);
        k_build_grid<<<1, 1024, 0, st>>>(V2.keys, n2, V2.gp, h->d_cell_start.p, h->d_sorted.p, h->d_kp_cell.p);
        h->timer.end(st);
        ++launches;
        h->grid_key = f2->cache_key; h->grid_n = n2;
    }
    const size_t smem = (size_t)n2 * sizeof(uint16_t);
    if (smem > 48 * 1024) PLVS_CUDA(cudaFuncSetAttribute(k_init_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (;;) {
        if ((rc = h->d_cand.alloc((size_t)n1 * h->cap))) return rc;
        h->timer.begin(PLVS_MATCH_K_INIT, st);
        k_init_candidates<<<div_up(n1, 8), 256, 0, st>>>(V1, V2, h->d_cell_start.p, h->d_sorted.p, reinterpret_cast<const float2*>(h->d_su.p), (float)window_size,
                                                          h->d_cand.p, h->d_cand_n.p, h->cap);
        PLVS_CUDA(cudaMemcpyAsync(h->p_cand_n.h, h->d_cand_n.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, st));
        k_init_resolve<<<1, 32, smem, st>>>(h->d_cand.p, h->d_cand_n.p, h->cap, V1.keys, V2.keys, n1, n2, nn_ratio, check_orientation, h->d_assign.p,
                                            h->d_claim_a.p, h->d_target.p, reinterpret_cast<float2*>(h->d_su.p), h->p_assign.d,
                                            reinterpret_cast<float2*>(h->p_su.d), h->p_result.d);
        h->timer.end(st);
        launches += 2;
        PLVS_CUDA(cudaGetLastError());
        PLVS_CUDA(cudaStreamSynchronize(st));
        h->timer.collect();
        int mx = 0;
        for (int i = 0; i < n1; ++i) mx = std::max(mx, h->p_cand_n.h[i]);
        if (mx <= h->cap) break;
        while (h->cap < mx) h->cap *= 2;       // a window held more candidates than reserved: redo with room (d_su still holds the caller's positions)
    }
    std::memcpy(matches12, h->p_assign.h, (size_t)n1 * 4);
    std::memcpy(prev_matched, h->p_su.h, (size_t)n1 * 8);
    *nmatches = h->p_result.h[0];
    h->last_launches = launches;
    return PLVS_OK;
}

int plvs_distinctive_descriptors(plvs_match* h, const uint8_t* desc, const int32_t* offsets, int n_points, int32_t* best)
{
    if (!h || n_points < 0 || (n_points && (!desc || !offsets || !best))) { set_error("null argument"); return PLVS_EINVAL; }
    if (n_points == 0) return PLVS_OK;
    std::lock_guard<std::mutex> lock(h->mu);
    PLVS_CUDA(cudaSetDevice(h->device));
    const int total = offsets[n_points];
    int nmax = 0;
    for (int i = 0; i < n_points; ++i) { const int n = offsets[i + 1] - offsets[i]; if (n < 0) { set_error("offsets must be non-decreasing"); return PLVS_EINVAL; } nmax = std::max(nmax, n); }
    if (nmax > 320) { set_error("more than 320 observations of one map point are not supported"); return PLVS_ECAP; }
    int rc;
    cudaStream_t st = h->stream;
    if ((rc = h->d_desc[0].alloc((size_t)std::max(total, 1) * 32)) || (rc = h->d_fv_off[0].alloc(n_points + 1)) || (rc = h->d_assign.alloc(n_points)) ||
        (rc = h->p_assign.alloc(n_points))) return rc;
    PLVS_CUDA(cudaMemcpyAsync(h->d_desc[0].p, desc, (size_t)total * 32, cudaMemcpyHostToDevice, st));
    PLVS_CUDA(cudaMemcpyAsync(h->d_fv_off[0].p, offsets, (size_t)(n_points + 1) * 4, cudaMemcpyHostToDevice, st));
    const size_t smem = (size_t)std::max(nmax, 1) * nmax * sizeof(uint16_t);
    if (smem > 48 * 1024) PLVS_CUDA(cudaFuncSetAttribute(k_distinctive, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->grid_key = 0; h->grid_n = -1;          // d_desc[0] was overwritten: a cached frame grid no longer matches its descriptors
    k_distinctive<<<n_points, 32, smem, st>>>(h->d_desc[0].p, h->d_fv_off[0].p, h->p_assign.d);
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaStreamSynchronize(st));
    std::memcpy(best, h->p_assign.h, (size_t)n_points * 4);
    h->last_launches = 1;
    return PLVS_OK;
}

// position of every 8-bit string in Mihasher::query's enumeration of the strings with the same number of ones
// (Thirdparty/line_descriptor/src/binary_descriptor_matcher_custom.cpp:683-760: its `power` / `bit` walk), replayed once on the host
static void line_enumeration_rank(int rank[256])
{
    const int curb = 8;
    for (int s = 0; s <= 8; ++s) {
        int power[16] = {0}; unsigned bitstr = 0;
        for (int i = 0; i < s; ++i) power[i] = i;
        power[s] = curb + 1;
        int bit = s - 1, r = 0;
        for (;;) {
            if (bit != -1) {
                bitstr ^= (power[bit] == bit) ? (1u << power[bit]) : (3u << (power[bit] - 1));
                power[bit]++; bit--;
            } else {
                rank[bitstr & 0xffu] = r++;
                while (++bit < s && power[bit] == power[bit + 1] - 1) { bitstr ^= 1u << (power[bit] - 1); power[bit] = bit; }
                if (bit == s) break;
            }
        }
    }
}

int plvs_line_knn2(plvs_match* h, const uint8_t* query, int nq, const uint8_t* train, int nt, const uint8_t* mask, float nn_ratio,
                   int32_t* query_idx, int32_t* train_idx, float* dist, uint8_t* valid, int* n_rows, int* n_valid)
{
    if (!h || !n_rows || nq < 0 || nt < 0 || (nq && (!query || !query_idx || !train_idx || !dist || !valid)) || (nt && !train)) { set_error("null argument"); return PLVS_EINVAL; }
    *n_rows = 0; if (n_valid) *n_valid = 0;
    if (nq == 0 || nt == 0) { set_error("descriptors matrices cannot be void"); return PLVS_EINVAL; }          // the reference prints this and returns no matches (:262-266)
    if (nt < 2) { set_error("fewer train descriptors than k = 2: the reference returns uninitialised indices"); return PLVS_EINVAL; }
    if (nt >= (1 << 20)) { set_error("more than 2^20 train descriptors are not supported"); return PLVS_ECAP; }
    std::lock_guard<std::mutex> lock(h->mu);
    PLVS_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    int rc;
    if ((rc = h->d_line_rank.alloc(256)) || (rc = h->d_line_best.alloc((size_t)2 * nq)) || (rc = h->d_line_u8.alloc((size_t)32 * (nq + nt) + 2 * (size_t)nq)) ||
        (rc = h->d_line_f32.alloc((size_t)2 * nq)) || (rc = h->d_assign.alloc((size_t)3 * nq + 2)) || (rc = h->p_result.alloc(8))) return rc;
    if (!h->line_rank_ready) {
        int rank[256];
        line_enumeration_rank(rank);
        PLVS_CUDA(cudaMemcpyAsync(h->d_line_rank.p, rank, sizeof(rank), cudaMemcpyHostToDevice, st));
        PLVS_CUDA(cudaStreamSynchronize(st));       // `rank` is a local
        h->line_rank_ready = true;
    }
    uint8_t* d_q = h->d_line_u8.p; uint8_t* d_t = d_q + (size_t)32 * nq; uint8_t* d_mask = d_t + (size_t)32 * nt; uint8_t* d_valid = d_mask + nq;
    PLVS_CUDA(cudaMemcpyAsync(d_q, query, (size_t)32 * nq, cudaMemcpyHostToDevice, st));
    PLVS_CUDA(cudaMemcpyAsync(d_t, train, (size_t)32 * nt, cudaMemcpyHostToDevice, st));
    if (mask) PLVS_CUDA(cudaMemcpyAsync(d_mask, mask, (size_t)nq, cudaMemcpyHostToDevice, st));
    int32_t* d_qidx = h->d_assign.p; int32_t* d_tidx = d_qidx + nq; int* d_res = d_tidx + 2 * (size_t)nq;
    h->timer.begin(PLVS_MATCH_K_LINES, st);
    k_line_knn2<<<div_up(nq, 8), 256, 0, st>>>(d_q, nq, d_t, nt, h->d_line_rank.p, h->d_line_best.p);
    k_line_rows<<<1, 1024, 0, st>>>(h->d_line_best.p, nq, mask ? d_mask : nullptr, nn_ratio, d_qidx, d_tidx, h->d_line_f32.p, d_valid, d_res);
    h->timer.end(st);
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaMemcpyAsync(h->p_result.h, d_res, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
    PLVS_CUDA(cudaStreamSynchronize(st));
    const int rows = h->p_result.h[0];
    if (rows > 0) {
        PLVS_CUDA(cudaMemcpyAsync(query_idx, d_qidx, (size_t)rows * 4, cudaMemcpyDeviceToHost, st));
        PLVS_CUDA(cudaMemcpyAsync(train_idx, d_tidx, (size_t)rows * 8, cudaMemcpyDeviceToHost, st));
        PLVS_CUDA(cudaMemcpyAsync(dist, h->d_line_f32.p, (size_t)rows * 8, cudaMemcpyDeviceToHost, st));
        PLVS_CUDA(cudaMemcpyAsync(valid, d_valid, (size_t)rows, cudaMemcpyDeviceToHost, st));
        PLVS_CUDA(cudaStreamSynchronize(st));
    }
    h->timer.collect();
    *n_rows = rows; if (n_valid) *n_valid = h->p_result.h[1];
    h->grid_key = 0; h->grid_n = -1;
    h->last_launches = 2;
    return PLVS_OK;
}

int plvs_stereo_match(plvs_match* h, const plvs_frame_view* left, const plvs_frame_view* right,
                      const plvs_pyramid_view* pyr_left, const plvs_pyramid_view* pyr_right, const float* inv_scale,
                      float mb, float mbf, float* uright, float* depth, int* n_valid)
{
    if (!h || !left || !right || !pyr_left || !pyr_right || !inv_scale || !uright || !depth) { set_error("null argument"); return PLVS_EINVAL; }
    if (!(mb > 0.f)) { set_error("baseline must be positive"); return PLVS_EINVAL; }
    std::lock_guard<std::mutex> lock(h->mu);
    PLVS_CUDA(cudaSetDevice(h->device));
    ViewDev VL, VR;
    int rc;
    if ((rc = stage_view(h, 0, left, &VL)) || (rc = stage_view(h, 1, right, &VR))) return rc;
    const int n = left->n, nr = right->n;
    for (int i = 0; i < n; ++i) { uright[i] = -1.f; depth[i] = -1.f; }
    if (n_valid) *n_valid = 0;
    if (n == 0 || nr == 0) return PLVS_OK;
    PyrDev PL{}, PR{}; StereoScales S{};
    const int nl = std::min(pyr_left->nlevels, (int)PLVS_MAX_LEVELS);
    for (int l = 0; l < nl; ++l) {
        PL.data[l] = pyr_left->data[l]; PL.w[l] = pyr_left->w[l]; PL.h[l] = pyr_left->h[l]; PL.pitch[l] = pyr_left->pitch[l];
        PR.data[l] = pyr_right->data[l]; PR.w[l] = pyr_right->w[l]; PR.h[l] = pyr_right->h[l]; PR.pitch[l] = pyr_right->pitch[l];
        S.scale[l] = left->scale_factors[l]; S.inv[l] = inv_scale[l];
    }
    if ((rc = h->d_su.alloc(n)) || (rc = h->d_sd.alloc(n)) || (rc = h->d_assign.alloc(n)) || (rc = h->p_su.alloc(n)) || (rc = h->p_sd.alloc(n)) ||
        (rc = h->p_result.alloc(4))) return rc;
    cudaStream_t st = h->stream;
    k_stereo_rows<<<div_up(n, 8), 256, 0, st>>>(VL.keys, VL.desc, n, VR.keys, VR.desc, nr, PL, PR, S, PL.h[0], mb, mbf, h->d_su.p, h->d_sd.p, h->d_assign.p);
    k_stereo_median<<<1, 1024, 0, st>>>(n, h->d_assign.p, h->d_su.p, h->d_sd.p, h->p_su.d, h->p_sd.d, h->p_result.d);
    PLVS_CUDA(cudaGetLastError());
    PLVS_CUDA(cudaStreamSynchronize(st));
    std::memcpy(uright, h->p_su.h, (size_t)n * 4);
    std::memcpy(depth, h->p_sd.h, (size_t)n * 4);
    if (n_valid) *n_valid = h->p_result.h[0];
    h->last_launches = 2;
    return PLVS_OK;
}

}  // extern "C"
