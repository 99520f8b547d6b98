// Device code only (see match_common.cuh): the monocular initialiser's search.
#pragma once
#ifndef PLVS_DYN_SMEM
#define PLVS_DYN_SMEM(T, name) extern __shared__ T name[]
#endif

// ---------------------------------------------------------------------------------------------
// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:732-852), monocular start-up.
// k_init_candidates: one warp per level-0 keypoint of F1 lists the level-0 keypoints of F2 inside the square window around its
// previously matched position, in GetFeaturesInArea order, with their Hamming distances (all of it independent of the other queries).
// k_init_resolve: what is left IS sequential -- a candidate is skipped when an earlier accepted match on it was at least as good
// (vMatchedDistance), and a winner takes a feature away from an earlier query (vnMatches21) -- so one warp walks the queries in
// order with the matched distances in shared memory: lanes over the candidates, best = first position of the smallest distance,
// second = second smallest of the multiset (what the if / else-if pair computes).  Mono initialisation handles ~10^3 level-0
// keypoints a few times per session; the sequential pass is a few hundred microseconds.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_init_candidates(ViewDev F1, ViewDev F2, const int* __restrict__ cell_start, const int* __restrict__ sorted, const float2* __restrict__ prev,
                  float window, uint32_t* __restrict__ cand, int* __restrict__ cand_n, int cap)
{
    const int q = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (q >= F1.n) return;
    const int level1 = F1.keys[q].octave;
    int count = 0, c0, c1, r0, r1;
    const float2 pm = prev[q];
    if (level1 <= 0 && cell_window(F2.gp, pm.x, pm.y, window, c0, c1, r0, r1)) {
        const bool check = (level1 > 0) || (level1 >= 0);               // GetFeaturesInArea(x, y, r, level1, level1)
        const uint8_t* qd = F1.desc + (size_t)q * 32;
        const uint4 a0 = *reinterpret_cast<const uint4*>(qd), a1 = *reinterpret_cast<const uint4*>(qd + 16);
        uint32_t* out = cand + (size_t)q * cap;
        for (int ix = c0; ix <= c1; ++ix) {
            const int pbeg = cell_start[ix * GRID_ROWS + r0], pend = cell_start[ix * GRID_ROWS + r1 + 1];
            for (int p = pbeg + lane; p < ((pend - pbeg + 31) / 32) * 32 + pbeg; p += 32) {
                bool ok = p < pend;
                int idx = 0, oct = 0, dist = 0;
                if (ok) {
                    idx = sorted[p];
                    const plvs_keypoint kp = F2.keys[idx];
                    oct = kp.octave;
                    if (check && (oct < level1 || oct > level1)) ok = false;
                    if (ok && !(fabsf(kp.x - pm.x) < window && fabsf(kp.y - pm.y) < window)) ok = false;
                    if (ok) dist = hamming256(a0, a1, F2.desc + (size_t)idx * 32);
                }
                const uint32_t m = __ballot_sync(0xffffffffu, ok);
                if (ok) { const int pos = count + __popc(m & ((1u << lane) - 1)); if (pos < cap) out[pos] = pack_cand(idx, dist, oct & 31); }
                count += __popc(m);
            }
        }
    }
    if (lane == 0) cand_n[q] = count;
}

__global__ void __launch_bounds__(32)
k_init_resolve(const uint32_t* __restrict__ cand, const int* __restrict__ cand_n, int cap, const plvs_keypoint* __restrict__ k1,
               const plvs_keypoint* __restrict__ k2, int n1, int n2, float ratio, int check_ori, int32_t* __restrict__ m12, int32_t* __restrict__ m21,
               int* __restrict__ bin_of, float2* __restrict__ prev, int32_t* __restrict__ m12_out, float2* __restrict__ prev_out, int* __restrict__ result)
{
    PLVS_DYN_SMEM(uint16_t, s_md);                      // vMatchedDistance: 0xffff = INT_MAX, else a Hamming distance
    __shared__ int s_hist[HISTO], s_keep[HISTO];
    const int lane = threadIdx.x;
    for (int i = lane; i < n2; i += 32) { s_md[i] = 0xffff; m21[i] = -1; }
    for (int i = lane; i < n1; i += 32) { m12[i] = -1; bin_of[i] = -1; }
    if (lane < HISTO) { s_hist[lane] = 0; s_keep[lane] = 1; }
    __syncwarp();
    int nmatches = 0;                                   // lane 0's copy is the one that counts
    for (int q = 0; q < n1; ++q) {
        const int cn = min(cand_n[q], cap);
        if (cn == 0) continue;
        int best = INT_MAX, bpos = INT_MAX, bidx = -1, second = INT_MAX;
        const uint32_t* list = cand + (size_t)q * cap;
        for (int p = lane; p < cn; p += 32) {
            const uint32_t c = list[p];
            const int i2 = cand_idx(c), d = cand_dist(c);
            if ((int)s_md[i2] <= d) continue;           // an earlier query holds this feature with a distance at least as small
            if (d < best) { second = best; best = d; bpos = p; bidx = i2; }
            else if (d < second) second = d;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best, o), op = __shfl_xor_sync(0xffffffffu, bpos, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bidx, o), os = __shfl_xor_sync(0xffffffffu, second, o);
            const int ns = min(min(second, os), max(best, ob));
            if (ob < best || (ob == best && op < bpos)) { best = ob; bpos = op; bidx = oi; }
            second = ns;
        }
        if (best <= TH_LOW && (float)best < (float)second * ratio) {
            if (lane == 0) {
                const int before = m21[bidx];
                if (before >= 0) { m12[before] = -1; --nmatches; }
                m12[q] = bidx; m21[bidx] = q; s_md[bidx] = (uint16_t)best; ++nmatches;
                if (check_ori) {
                    float rot = k1[q].angle - k2[bidx].angle;
                    if (rot < 0.0f) rot += 360.0f;
                    int bin = (int)roundf(rot * (HISTO / 360.0f));
                    if (bin == HISTO) bin = 0;
                    bin_of[q] = bin; ++s_hist[bin];      // stays in the histogram even if the match is stolen later
                }
            }
            __syncwarp();
        }
    }
    __syncwarp();
    int dropped = 0;
    if (check_ori) {
        if (lane == 0) {
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
        __syncwarp();
        for (int i = lane; i < n1; i += 32) {
            const int b = bin_of[i];
            if (b >= 0 && !s_keep[b] && m12[i] >= 0) { m12[i] = -1; ++dropped; }
        }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) dropped += __shfl_xor_sync(0xffffffffu, dropped, o);
    for (int i = lane; i < n1; i += 32) {
        const int j = m12[i];
        float2 pm = prev[i];
        if (j >= 0) pm = make_float2(k2[j].x, k2[j].y);                 // "update prev matched" (:847-849)
        m12_out[i] = j; prev_out[i] = pm;
    }
    if (lane == 0) result[0] = nmatches - dropped;
}
