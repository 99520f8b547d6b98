// Device code only (see match_common.cuh): Tracking::SearchLocalPoints' visibility pass and the compaction of its survivors.
#pragma once

// ---------------------------------------------------------------------------------------------
// Frame::isInFrustum (src/Frame.cc:955-1017) + Pinhole::project (Pinhole.cpp:61-67) + MapPoint::PredictScale (MapPoint.cc:598-613), one thread
// per map point.  fp32 in the reference's operation order (Eigen's e0 + (e1 + e2), IEEE sqrt and divisions, no contraction).  PredictScale's
// ceil(logf(ratio) / logf(scaleFactor)) is evaluated as a count of host-computed thresholds (see plvs_match_in_frustum), so no device logarithm
// has to match glibc's.
// ---------------------------------------------------------------------------------------------
struct FrustumDev { plvs_frustum f; float T[PLVS_MAX_LEVELS]; };

__global__ void __launch_bounds__(256)
k_in_frustum(FrustumDev D, const plvs_map_point* __restrict__ pts, int n, plvs_mp_query* __restrict__ q, uint8_t* __restrict__ in_view, int* __restrict__ count)
{
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const plvs_frustum& fr = D.f;
    const plvs_map_point p = pts[i];
    plvs_mp_query o;
    o.proj_x = -1.f; o.proj_y = -1.f; o.proj_xr = 0.f; o.track_depth = 0.f; o.view_cos = 0.f; o.level = 0; o.flags = p.flags;
#pragma unroll
    for (int k = 0; k < 32; ++k) o.desc[k] = p.desc[k];
    bool in = false;
    const float X = p.xw[0], Y = p.xw[1], Z = p.xw[2];
    const float pcx = (fr.Rcw[0] * X + (fr.Rcw[1] * Y + fr.Rcw[2] * Z)) + fr.tcw[0];
    const float pcy = (fr.Rcw[3] * X + (fr.Rcw[4] * Y + fr.Rcw[5] * Z)) + fr.tcw[1];
    const float pcz = (fr.Rcw[6] * X + (fr.Rcw[7] * Y + fr.Rcw[8] * Z)) + fr.tcw[2];
    if (!(pcz < 0.0f)) {
        const float u = fr.fx * pcx / pcz + fr.cx, v = fr.fy * pcy / pcz + fr.cy;
        if (!(u < fr.min_x || u > fr.max_x) && !(v < fr.min_y || v > fr.max_y)) {
            o.proj_x = u; o.proj_y = v;
            const float maxDistance = 1.2f * p.max_dist, minDistance = 0.8f * p.min_dist;
            const float pox = X - fr.Ow[0], poy = Y - fr.Ow[1], poz = Z - fr.Ow[2];
            const float dist = sqrtf(pox * pox + (poy * poy + poz * poz));
            if (!(dist < minDistance || dist > maxDistance)) {
                const float viewCos = (pox * p.normal[0] + (poy * p.normal[1] + poz * p.normal[2])) / dist;
                if (!(viewCos < fr.viewing_cos_limit)) {
                    const float ratio = p.max_dist / dist;
                    int lvl = 0;
                    for (int k = 0; k + 1 < fr.nlevels; ++k) lvl += ratio >= D.T[k];
                    o.level = lvl;
                    o.proj_xr = u - fr.bf * (1.0f / pcz);
                    o.track_depth = sqrtf(pcx * pcx + (pcy * pcy + pcz * pcz));
                    o.view_cos = viewCos;
                    in = true;
                }
            }
        }
    }
    q[i] = o;
    in_view[i] = in ? 1 : 0;
    if (in) atomicAdd(count, 1);
}

// stable compaction of the in-view queries (the order SearchByProjection walks vpMapPoints in): one CTA, block-wide scan in chunks of 1024
__global__ void __launch_bounds__(1024)
k_compact_queries(const plvs_mp_query* __restrict__ q, const uint8_t* __restrict__ in_view, int n, plvs_mp_query* __restrict__ out, int32_t* __restrict__ src_index)
{
    __shared__ int s_warp[32];
    __shared__ int s_base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int c0 = 0; c0 < n; c0 += 1024) {
        const int i = c0 + tid;
        const int f = (i < n && in_view[i]) ? 1 : 0;
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
        if (f) { out[pos] = q[i]; src_index[pos] = i; }
        __syncthreads();
        if (tid == 1023) s_base += s_warp[31];
        __syncthreads();
    }
}
