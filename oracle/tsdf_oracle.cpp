// TEST INFRASTRUCTURE ONLY -- CPU oracle for the TSDF rows of SURVEY.md §8a (a18-a25, a27).
//
// Restates open_chisel's depth-scan integration as PLVS reaches it:
//   Chisel::IntegrateDepthScan<T>                       Thirdparty/open_chisel/include/open_chisel/Chisel.h:68-131
//   Chisel::IntegrateDepthScanColorWithOneCameraModelBGR  Chisel.h:198-258
//   ProjectionIntegrator::Integrate / IntegrateColorWithOneCameraModelBGR   ProjectionIntegrator.h:58-108, 189-269
//   DistVoxel::Integrate/Carve/Reset (DistVoxel.h:91-117), ColorVoxel::IntegrateSimple (ColorVoxel.h:91-110)
//   QuadraticTruncator (truncation/QuadraticTruncator.h:45-50), ConstantWeighter (weighting/ConstantWeighter.h:43-46)
//   PinholeCamera::ProjectPoint/IsPointOnImage/SetupFrustum (src/camera/PinholeCamera.cpp:38-64)
//   Frustum::SetFromParams/SetFromVectors/ComputeBoundingBox/Intersects (src/geometry/Frustum.cpp:41-222)
//   Plane(a,b,c) (src/geometry/Plane.cpp:44-52), ChunkManager::GetChunkIDsIntersecting / GetIDAt / CacheCentroids
//   (src/ChunkManager.cpp:65-94,241-271, ChunkManager.h:192-201), DepthImage::GetStats/DepthAt (camera/DepthImage.h:54-108)
// Every chunk of the frustum's padded bounding box is created, every one of its 4096 voxels is
// evaluated, and chunks that were new and received no update are collected -- exactly the
// reference's (brute-force) control flow.  Eigen is not available here; 3-vector reductions follow
// Eigen's unrolled order e0 + (e1 + e2) and fp contraction is off (DESIGN.md "TSDF float order").
//
// Parity status: "parity unpinned" -- the reference has no test for this path and cannot be built
// in this image (needs Eigen/PCL); tolerance for the CUDA path is |d| <= 1e-4 on sdf/weight and an
// identical chunk-key set.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <unordered_map>
#include <vector>
#include <omp.h>

namespace {

struct V3 { float x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }       // Eigen redux order for size 3
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

struct Plane {
    V3 n; float d;
    Plane() : n{0, 0, 0}, d(0) {}
    Plane(V3 a, V3 b, V3 c)
    {
        V3 ab = b - a, ac = c - a;
        V3 cr = cross(ab, ac);
        float nn = dot(cr, cr);
        n = nn > 0 ? V3{cr.x / std::sqrt(nn), cr.y / std::sqrt(nn), cr.z / std::sqrt(nn)} : cr;
        d = -dot(cr, a);              // un-normalised on purpose: that is what the reference does (Plane.cpp:51)
    }
};

struct Params {
    float voxel_resolution, trunc_quad, trunc_linear, trunc_const, trunc_scale, weight;
    int32_t use_carving; float carving_dist; int32_t use_color; float near_plane, far_plane; int32_t max_blocks;
};

struct Key { int x, y, z; bool operator==(const Key& o) const { return x == o.x && y == o.y && z == o.z; } };
struct KeyHash { size_t operator()(const Key& k) const { return ((size_t)k.x * 73856093u) ^ ((size_t)k.y * 19349663u) ^ ((size_t)k.z * 83492791u); } };

struct Block {
    std::vector<float> sdf, w;
    std::vector<uint8_t> rgba;    // r,g,b,colour weight
    Block() : sdf(4096, 99999.f), w(4096, 0.f), rgba(4096 * 4, 0) {}
};

struct Map {
    Params p;
    float fx, fy, cx, cy; int width, height; bool got_camera = false;
    std::unordered_map<Key, std::unique_ptr<Block>, KeyHash> blocks;
    int n_range = 0, n_updated = 0, n_new = 0, n_collected = 0;
    int threads = 1;
};

struct Frustum { Plane far_, near_, top, bottom, left, right; V3 corners[8]; };

void frustum_from_params(Frustum& f, const float* Twc, float nearD, float farD, float fx, float fy, float cy, float w, float h)
{
    // view.linear() = R (columns = camera axes in world), translation = t
    V3 rightV{Twc[0], Twc[4], Twc[8]}, up{-Twc[1], -Twc[5], -Twc[9]}, fwd{Twc[2], Twc[6], Twc[10]}, pos{Twc[3], Twc[7], Twc[11]};
    float aspect = (fx * w) / (fy * h);
    float fov = (float)(std::atan2((double)cy, (double)fy) + std::atan2((double)(h - cy), (double)fy));
    const float tang = (float)std::tan((double)(fov / 2));
    const float hF = tang * farD, wF = hF * aspect, hN = tang * nearD, wN = hN * aspect;
    const V3 fc = pos + fwd * farD;
    const V3 ftl = fc + (up * hF) - (rightV * wF), ftr = fc + (up * hF) + (rightV * wF);
    const V3 fbl = fc - (up * hF) - (rightV * wF), fbr = fc - (up * hF) + (rightV * wF);
    const V3 nc = pos + fwd * nearD;
    const V3 ntl = nc + (up * hN) - (rightV * wN), ntr = nc + (up * hN) + (rightV * wN);
    const V3 nbl = nc - (up * hN) - (rightV * wN), nbr = nc - (up * hN) + (rightV * wN);
    f.near_ = Plane(nbl, ntl, nbr); f.far_ = Plane(ftr, ftl, fbr);
    f.left = Plane(ftl, ntl, fbl);  f.right = Plane(ntr, ftr, nbr);
    f.top = Plane(ntl, ftl, ntr);   f.bottom = Plane(nbr, fbl, nbl);
    f.corners[0] = ftl; f.corners[1] = ftr; f.corners[2] = fbl; f.corners[3] = fbr;
    f.corners[4] = nbr; f.corners[5] = ntl; f.corners[6] = ntr; f.corners[7] = nbl;
}

bool intersects(const Frustum& f, V3 mn, V3 mx)
{
    const Plane* planes[6] = {&f.far_, &f.near_, &f.top, &f.bottom, &f.left, &f.right};
    for (const Plane* pl : planes) {
        V3 v;
        v.x = pl->n.x < 0.0f ? mn.x : mx.x;
        v.y = pl->n.y < 0.0f ? mn.y : mx.y;
        v.z = pl->n.z < 0.0f ? mn.z : mx.z;
        if (dot(v, pl->n) + pl->d > 0.0f) return true;
    }
    return false;
}

}  // namespace

extern "C" {

void* orc_tsdf_create(const Params* p, int threads)
{
    Map* m = new Map();
    m->p = *p;
    m->threads = std::max(1, threads);
    return m;
}
void orc_tsdf_destroy(void* h) { delete (Map*)h; }
void orc_tsdf_reset(void* h) { ((Map*)h)->blocks.clear(); }

void orc_tsdf_set_camera(void* h, double fx, double fy, double cx, double cy, int w, int ht)
{
    Map* m = (Map*)h;
    m->fx = (float)fx; m->fy = (float)fy; m->cx = (float)cx; m->cy = (float)cy; m->width = w; m->height = ht;
    m->got_camera = true;
}

// enumerate the chunk range of the reference for this scan: ids[3*i..] (cap entries), returns count
int orc_tsdf_chunk_range(void* h, const float* depth, int w, int ht, const float* Twc, int mode, int32_t* lo, int32_t* hi)
{
    Map* m = (Map*)h;
    float nearD = m->p.near_plane, farD = m->p.far_plane;
    if (mode == 0) {     // IntegrateDepthScan: planes from DepthImage::GetStats (zeros and NaNs skipped)
        float mn = std::numeric_limits<float>::max(), mx = -std::numeric_limits<float>::max();
        for (int i = 0; i < w * ht; ++i) { float d = depth[i]; if (d == 0 || std::isnan(d)) continue; mn = std::min(d, mn); mx = std::max(d, mx); }
        nearD = mn; farD = mx;
    }
    Frustum f;
    frustum_from_params(f, Twc, nearD, farD, m->fy, m->fy, m->cy, (float)m->width, (float)m->height);   // fy twice: PinholeCamera.cpp:58
    const float big = std::numeric_limits<float>::max();
    V3 mn{big, big, big}, mx{-big, -big, -big};
    for (int i = 0; i < 8; ++i) {
        mn.x = std::min(mn.x, f.corners[i].x); mn.y = std::min(mn.y, f.corners[i].y); mn.z = std::min(mn.z, f.corners[i].z);
        mx.x = std::max(mx.x, f.corners[i].x); mx.y = std::max(mx.y, f.corners[i].y); mx.z = std::max(mx.z, f.corners[i].z);
    }
    const float rf = 1.0f / (16 * m->p.voxel_resolution);
    int minID[3] = {(int)std::floor(mn.x * rf), (int)std::floor(mn.y * rf), (int)std::floor(mn.z * rf)};
    int maxID[3] = {(int)std::floor(mx.x * rf) + 1, (int)std::floor(mx.y * rf) + 1, (int)std::floor(mx.z * rf) + 1};
    for (int a = 0; a < 3; ++a) { lo[a] = minID[a] - 1; hi[a] = maxID[a] + 1; }
    return 0;
}

int orc_tsdf_integrate(void* h, const float* depth, int w, int ht, const uint8_t* bgr, int nch, const float* Twc, int mode)
{
    Map* m = (Map*)h;
    if (!m->got_camera || !depth) return -5;
    const Params& P = m->p;
    const float res = P.voxel_resolution;
    int32_t lo[3], hi[3];
    orc_tsdf_chunk_range(h, depth, w, ht, Twc, mode, lo, hi);
    // frustum again for the (lax) intersection test of each chunk box
    float nearD = P.near_plane, farD = P.far_plane;
    if (mode == 0) {
        float mn = std::numeric_limits<float>::max(), mx = -std::numeric_limits<float>::max();
        for (int i = 0; i < w * ht; ++i) { float d = depth[i]; if (d == 0 || std::isnan(d)) continue; mn = std::min(d, mn); mx = std::max(d, mx); }
        nearD = mn; farD = mx;
    }
    Frustum fr;
    frustum_from_params(fr, Twc, nearD, farD, m->fy, m->fy, m->cy, (float)m->width, (float)m->height);
    std::vector<Key> list;
    for (int x = lo[0]; x <= hi[0]; ++x)
        for (int y = lo[1]; y <= hi[1]; ++y)
            for (int z = lo[2]; z <= hi[2]; ++z) {
                V3 mn{(float)(x * 16) * res, (float)(y * 16) * res, (float)(z * 16) * res};
                V3 mx = mn + V3{16.f * res, 16.f * res, 16.f * res};
                if (intersects(fr, mn, mx)) list.push_back(Key{x, y, z});
            }
    m->n_range = (int)list.size();

    const float diag = (float)(2.0 * (double)std::sqrt(3.0f) * (double)res);
    const float half = res * 0.5f;
    // R^T rows = columns of R
    const float r00 = Twc[0], r01 = Twc[1], r02 = Twc[2], r10 = Twc[4], r11 = Twc[5], r12 = Twc[6], r20 = Twc[8], r21 = Twc[9], r22 = Twc[10];
    const V3 t{Twc[3], Twc[7], Twc[11]};
    const float fx = m->fx, fy = m->fy, cx = m->cx, cy = m->cy;
    const int width = m->width, height = m->height;

    // create missing chunks first (sequential map mutation), then evaluate chunks in parallel
    std::vector<Block*> blk(list.size());
    std::vector<char> is_new(list.size(), 0), updated(list.size(), 0);
    for (size_t i = 0; i < list.size(); ++i) {
        auto it = m->blocks.find(list[i]);
        if (it == m->blocks.end()) { is_new[i] = 1; it = m->blocks.emplace(list[i], std::make_unique<Block>()).first; }
        blk[i] = it->second.get();
    }
#pragma omp parallel for schedule(dynamic, 16) num_threads(m->threads)
    for (long ci = 0; ci < (long)list.size(); ++ci) {
        Block& B = *blk[ci];
        const Key k = list[ci];
        const V3 origin{(float)(16 * k.x) * res, (float)(16 * k.y) * res, (float)(16 * k.z) * res};
        bool upd = false;
        int i = 0;
        for (int z = 0; z < 16; ++z)
            for (int y = 0; y < 16; ++y)
                for (int x = 0; x < 16; ++x, ++i) {
                    const V3 centroid{(float)x * res + half, (float)y * res + half, (float)z * res + half};
                    const V3 c = centroid + origin;
                    const V3 dv = c - t;
                    const V3 pc{r00 * dv.x + (r10 * dv.y + r20 * dv.z), r01 * dv.x + (r11 * dv.y + r21 * dv.z), r02 * dv.x + (r12 * dv.y + r22 * dv.z)};
                    const float invZ = 1.0f / pc.z;
                    const float u = fx * pc.x * invZ + cx, v = fy * pc.y * invZ + cy;
                    if (!(u >= 0 && v >= 0 && u < width && v < height) || pc.z < 0) continue;
                    const float d = depth[(int)u + (int)v * w];
                    if (std::isnan(d)) continue;
                    const float trunc = (P.trunc_quad * d * d + P.trunc_linear * d + P.trunc_const) * P.trunc_scale;
                    const float s = d - pc.z;
                    if (std::fabs(s) < trunc + diag) {
                        float wu = 1.0f;
                        if (mode == 1) {
                            uint8_t* cv = &B.rgba[(size_t)i * 4];
                            if (cv[3] < 5) {        // ColorVoxel::IntegrateSimple(…,1)
                                const uint8_t* px = bgr + ((size_t)(int)u + (size_t)(int)v * w) * nch;
                                const uint8_t nb = px[0], ng = px[1], nr = px[2];
                                if (!(cv[3] >= 255 - 1)) {
                                    const float inv = 1.f / (float)(1 + cv[3]);
                                    cv[0] = (uint8_t)((float)(cv[3] * cv[0] + 1 * nr) * inv);
                                    cv[1] = (uint8_t)((float)(cv[3] * cv[1] + 1 * ng) * inv);
                                    cv[2] = (uint8_t)((float)(cv[3] * cv[2] + 1 * nb) * inv);
                                    cv[3] = (uint8_t)(cv[3] + 1);
                                }
                            }
                            wu = P.weight / (2.0f * trunc);
                        }
                        const float ow = B.w[i], os = B.sdf[i];
                        B.sdf[i] = (ow * os + wu * s) / (wu + ow);
                        B.w[i] = ow + wu;
                        upd = true;
                    } else if (P.use_carving && s > trunc + P.carving_dist) {
                        if (B.w[i] > 0 && B.sdf[i] < 1e-5) {
                            if (mode == 1) { B.sdf[i] = 99999.f; B.w[i] = 0.f; }         // Reset()
                            else { const float ow = B.w[i], os = B.sdf[i]; B.sdf[i] = (ow * os + 1.5f * 0.0f) / (1.5f + ow); B.w[i] = ow + 1.5f; }   // Carve()
                            upd = true;
                        }
                    }
                }
        updated[ci] = upd;
    }
    m->n_updated = m->n_new = m->n_collected = 0;
    for (size_t i = 0; i < list.size(); ++i) {
        if (updated[i]) { ++m->n_updated; if (is_new[i]) ++m->n_new; }
        else if (is_new[i]) { m->blocks.erase(list[i]); ++m->n_collected; }
    }
    return 0;
}

void orc_tsdf_stats(void* h, int32_t* out /*n_blocks,n_range,n_updated,n_new,n_collected*/)
{
    Map* m = (Map*)h;
    out[0] = (int)m->blocks.size(); out[1] = m->n_range; out[2] = m->n_updated; out[3] = m->n_new; out[4] = m->n_collected;
}

// blocks sorted by (x,y,z) so two maps can be compared element-wise
int orc_tsdf_download(void* h, int32_t* keys, float* sdf, float* weight, uint8_t* rgba, int cap)
{
    Map* m = (Map*)h;
    std::map<std::tuple<int, int, int>, Block*> ord;
    for (auto& kv : m->blocks) ord[{kv.first.x, kv.first.y, kv.first.z}] = kv.second.get();
    int n = 0;
    for (auto& kv : ord) {
        if (n >= cap) break;
        if (keys) { keys[3 * n] = std::get<0>(kv.first); keys[3 * n + 1] = std::get<1>(kv.first); keys[3 * n + 2] = std::get<2>(kv.first); }
        if (sdf) std::memcpy(sdf + (size_t)n * 4096, kv.second->sdf.data(), 4096 * 4);
        if (weight) std::memcpy(weight + (size_t)n * 4096, kv.second->w.data(), 4096 * 4);
        if (rgba) std::memcpy(rgba + (size_t)n * 16384, kv.second->rgba.data(), 16384);
        ++n;
    }
    return (int)ord.size();
}

}  // extern "C"
