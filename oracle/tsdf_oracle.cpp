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
// Parity status: PINNED -- tests/test_oracle_vs_reference_tsdf.py checks this file bit-exactly (chunk keys, sdf, weight,
// colour, chunk-range size) against the reference's own open_chisel sources, compiled into oracle/_ref/libchisel_ref.so by
// oracle/ref_build.py against the Eigen stand-in of oracle/eigen_standin (Eigen is not installed); goldens recorded from it
// are in tests/golden/tsdf_ref_*.npz.  Tolerance for the CUDA path: |d| <= 1e-4 on sdf/weight and an identical chunk-key set.
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
    std::vector<uint32_t> kfid;   // DistVoxel::kfid (DistVoxel.h:64-86,107-117): written by the point-cloud route, zeroed by Reset()
    Block() : sdf(4096, 99999.f), w(4096, 0.f), rgba(4096 * 4, 0), kfid(4096, 0u) {}
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

    // The reference creates every missing chunk of the range first, evaluates all of them and garbage-collects the new ones that
    // stayed untouched (Chisel.h:86-129).  Same result with bounded memory: a missing chunk is evaluated in a scratch block that
    // only enters the map if a voxel changed (at 5 mm / 5 m the range holds ~400 k chunks = 25 GB if all were kept alive at once).
    std::vector<Block*> blk(list.size());
    std::vector<std::unique_ptr<Block>> fresh(list.size());
    std::vector<char> is_new(list.size(), 0), updated(list.size(), 0);
    for (size_t i = 0; i < list.size(); ++i) {
        auto it = m->blocks.find(list[i]);
        if (it == m->blocks.end()) { is_new[i] = 1; blk[i] = nullptr; }
        else blk[i] = it->second.get();
    }
#pragma omp parallel for schedule(dynamic, 16) num_threads(m->threads)
    for (long ci = 0; ci < (long)list.size(); ++ci) {
        std::unique_ptr<Block> scratch;
        if (!blk[ci]) { scratch = std::make_unique<Block>(); blk[ci] = scratch.get(); }
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
                            if (mode == 1) { B.sdf[i] = 99999.f; B.w[i] = 0.f; B.kfid[i] = 0u; }         // Reset()
                            else { const float ow = B.w[i], os = B.sdf[i]; B.sdf[i] = (ow * os + 1.5f * 0.0f) / (1.5f + ow); B.w[i] = ow + 1.5f; }   // Carve()
                            upd = true;
                        }
                    }
                }
        updated[ci] = upd;
        if (scratch && upd) fresh[ci] = std::move(scratch);
    }
    m->n_updated = m->n_new = m->n_collected = 0;
    for (size_t i = 0; i < list.size(); ++i) {
        if (updated[i]) { ++m->n_updated; if (is_new[i]) { ++m->n_new; m->blocks.emplace(list[i], std::move(fresh[i])); } }
        else if (is_new[i]) ++m->n_collected;
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

// DistVoxel::GetKfid of every voxel, blocks in the order of orc_tsdf_download
int orc_tsdf_download_kfid(void* h, uint32_t* kfid, int cap)
{
    Map* m = (Map*)h;
    std::map<std::tuple<int, int, int>, Block*> ord;
    for (auto& kv : m->blocks) ord[{kv.first.x, kv.first.y, kv.first.z}] = kv.second.get();
    int n = 0;
    for (auto& kv : ord) { if (n >= cap) break; std::memcpy(kfid + (size_t)n * 4096, kv.second->kfid.data(), 4096 * 4); ++n; }
    return (int)ord.size();
}

// ---------------------------------------------------------------------------------------------
// a26: Chisel::IntegratePointCloudWidthDepth (Thirdparty/open_chisel/src/Chisel.cpp:382-585) -- PLVS's default
// Chisel route (src/PointCloudMapping.cc:641-642): (i) ProjectionIntegrator::CarveWithDepth over the existing chunks
// of the camera frustum (ProjectionIntegrator.h:271-335), (ii) per cloud point an Amanatides-Woo voxel walk
// (src/geometry/Raycast.cpp:65-182) over +-max(trunc, diag) along the ray with
// u = |Pc| (depth/Pc.z - 1), w = weight/(2 trunc), DistVoxel::Integrate + ColorVoxel::IntegrateSimple,
// (iii) garbage collection of chunks that were created but never updated.  Points are processed in order, so the
// running means see the reference's sequence.  Eigen::Affine3f::inverse() is a general 3x3 inverse (cofactors /
// determinant), restated as such.
// ---------------------------------------------------------------------------------------------
static inline float sgn_f(int x) { return x > 0 ? 1.f : x < 0 ? -1.f : 0.f; }
static inline float mod1(float value, float modulus) { return std::fmod(std::fmod(value, modulus) + modulus, modulus); }
static float intbound(float s, int ds)
{
    // smallest positive t such that s + t*ds is an integer (Raycast.cpp)
    if (ds < 0) return intbound(-s, -ds);
    s = mod1(s, 1.f);
    return (1 - s) / ds;
}

static const uint32_t* g_cloud_kfids = nullptr;      // per-point keyframe ids of the cloud being integrated (cloud.GetKfids(), src/Chisel.cpp:470)
static uint32_t g_cloud_kfid_all = 0;

int orc_tsdf_integrate_cloud(void* h, const float* xyz, const float* rgb, int n, const float* depth, int w, int ht, const float* Twc);
// the same with keyframe ids: kfids[n], or NULL for one id for the whole cloud
int orc_tsdf_integrate_cloud_kf(void* h, const float* xyz, const float* rgb, const uint32_t* kfids, uint32_t kfid_all, int n, const float* depth, int w, int ht, const float* Twc)
{
    g_cloud_kfids = kfids; g_cloud_kfid_all = kfid_all;
    const int rc = orc_tsdf_integrate_cloud(h, xyz, rgb, n, depth, w, ht, Twc);
    g_cloud_kfids = nullptr; g_cloud_kfid_all = 0;
    return rc;
}

int orc_tsdf_integrate_cloud(void* h, const float* xyz, const float* rgb, int n, const float* depth, int w, int ht, const float* Twc)
{
    Map* m = (Map*)h;
    const Params& P = m->p;
    const float res = P.voxel_resolution;
    const float diagD = (float)(2.0 * (double)std::sqrt(3.0f) * (double)res);
    const float r00 = Twc[0], r01 = Twc[1], r02 = Twc[2], r10 = Twc[4], r11 = Twc[5], r12 = Twc[6], r20 = Twc[8], r21 = Twc[9], r22 = Twc[10];
    const V3 t{Twc[3], Twc[7], Twc[11]};
    int n_carved = 0;
    if (P.use_carving && depth && m->got_camera) {
        int32_t lo[3], hi[3];
        orc_tsdf_chunk_range(h, depth, w, ht, Twc, 1, lo, hi);       // camera near/far planes (SetupFrustum of the stored camera)
        const float half = res * 0.5f;
        Frustum fr;
        frustum_from_params(fr, Twc, P.near_plane, P.far_plane, m->fy, m->fy, m->cy, (float)m->width, (float)m->height);
        for (auto& kv : m->blocks) {
            const Key k = kv.first;
            if (k.x < lo[0] || k.x > hi[0] || k.y < lo[1] || k.y > hi[1] || k.z < lo[2] || k.z > hi[2]) continue;   // ChunkManager::GetChunkIDsIntersecting
            {
                const V3 bmn{(float)(k.x * 16) * res, (float)(k.y * 16) * res, (float)(k.z * 16) * res};
                if (!intersects(fr, bmn, bmn + V3{16.f * res, 16.f * res, 16.f * res})) continue;
            }
            Block& B = *kv.second;
            const V3 origin{(float)(16 * k.x) * res, (float)(16 * k.y) * res, (float)(16 * k.z) * res};
            bool upd = false;
            int i = 0;
            for (int z = 0; z < 16; ++z) for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x, ++i) {
                if (B.w[i] <= 1e-15) continue;
                const V3 c = V3{(float)x * res + half, (float)y * res + half, (float)z * res + half} + origin;
                const V3 dv = c - t;
                const V3 pc{r00 * dv.x + (r10 * dv.y + r20 * dv.z), r01 * dv.x + (r11 * dv.y + r21 * dv.z), r02 * dv.x + (r12 * dv.y + r22 * dv.z)};
                const float invZ = 1.0f / pc.z;
                const float u = m->fx * pc.x * invZ + m->cx, v = m->fy * pc.y * invZ + m->cy;
                if (pc.z < 0 || !(u >= 0 && v >= 0 && u < m->width && v < m->height)) continue;
                const float d = depth[(int)u + (int)v * w];
                if (std::isnan(d)) continue;
                const float trunc = std::max((P.trunc_quad * d * d + P.trunc_linear * d + P.trunc_const) * P.trunc_scale, diagD);
                const float s = d - pc.z;
                if (s > trunc + P.carving_dist && B.sdf[i] < 1e-5) { B.sdf[i] = 99999.f; B.w[i] = 0.f; B.kfid[i] = 0u; upd = true; }
            }
            n_carved += upd;
        }
    }
    // inverse pose, Eigen style
    float inv[9], tinv[3];
    {
        const float a[9] = {r00, r01, r02, r10, r11, r12, r20, r21, r22};
        auto cof = [&](int i, int j) { return a[((i + 1) % 3) * 3 + (j + 1) % 3] * a[((i + 2) % 3) * 3 + (j + 2) % 3] - a[((i + 1) % 3) * 3 + (j + 2) % 3] * a[((i + 2) % 3) * 3 + (j + 1) % 3]; };
        const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
        const float det = c0 * a[0] + (c1 * a[3] + c2 * a[6]);
        const float invdet = 1.0f / det;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) inv[j * 3 + i] = cof(i, j) * invdet;
        for (int i = 0; i < 3; ++i) tinv[i] = -(inv[i * 3] * t.x + (inv[i * 3 + 1] * t.y + inv[i * 3 + 2] * t.z));
    }
    const float roundToVoxel = 1.0f / res;
    const float half = 0.5f * res;          // Vec3(0.5,0.5,0.5) * resolution
    const float rf = 1.0f / (16 * res);
    std::unordered_map<Key, bool, KeyHash> updated, created;
    m->n_updated = m->n_new = m->n_collected = 0;
    std::vector<Key> walk;
    for (int i = 0; i < n; ++i) {
        const V3 p{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        const float dpt = p.z;
        if (dpt < 0.01f) continue;
        const V3 wp{r00 * p.x + (r01 * p.y + r02 * p.z) + t.x, r10 * p.x + (r11 * p.y + r12 * p.z) + t.y, r20 * p.x + (r21 * p.y + r22 * p.z) + t.z};
        V3 dir = wp - t;
        const float nn = std::sqrt(dir.x * dir.x + (dir.y * dir.y + dir.z * dir.z));
        dir = V3{dir.x / nn, dir.y / nn, dir.z / nn};
        const float trunc = std::max((P.trunc_quad * dpt * dpt + P.trunc_linear * dpt + P.trunc_const) * P.trunc_scale, diagD);
        const V3 swp = wp * roundToVoxel;
        const V3 sdt = (dir * trunc) * roundToVoxel;
        const V3 start = swp - sdt, end = swp + sdt;
        // Raycast
        walk.clear();
        {
            int x = (int)std::floor(start.x), y = (int)std::floor(start.y), z = (int)std::floor(start.z);
            const int endX = (int)std::floor(end.x), endY = (int)std::floor(end.y), endZ = (int)std::floor(end.z);
            const V3 direction = end - start;
            const float maxDist = direction.x * direction.x + (direction.y * direction.y + direction.z * direction.z);
            const float dx = (float)(endX - x), dy = (float)(endY - y), dz = (float)(endZ - z);
            const int stepX = (int)sgn_f((int)dx), stepY = (int)sgn_f((int)dy), stepZ = (int)sgn_f((int)dz);
            float tMaxX = intbound(start.x, (int)dx), tMaxY = intbound(start.y, (int)dy), tMaxZ = intbound(start.z, (int)dz);
            const float tDeltaX = ((float)stepX) / dx, tDeltaY = ((float)stepY) / dy, tDeltaZ = ((float)stepZ) / dz;
            if (!(stepX == 0 && stepY == 0 && stepZ == 0)) {
                for (;;) {
                    walk.push_back(Key{x, y, z});
                    const V3 d3{(float)x - start.x, (float)y - start.y, (float)z - start.z};
                    const float dist = d3.x * d3.x + (d3.y * d3.y + d3.z * d3.z);
                    if (dist > maxDist) break;
                    if (x == endX && y == endY && z == endZ) break;
                    if (tMaxX < tMaxY) { if (tMaxX < tMaxZ) { x += stepX; tMaxX += tDeltaX; } else { z += stepZ; tMaxZ += tDeltaZ; } }
                    else { if (tMaxY < tMaxZ) { y += stepY; tMaxY += tDeltaY; } else { z += stepZ; tMaxZ += tDeltaZ; } }
                }
            }
        }
        for (const Key& vc : walk) {
            const V3 center{(float)vc.x * res + half, (float)vc.y * res + half, (float)vc.z * res + half};
            const Key cid{(int)std::floor(center.x * rf), (int)std::floor(center.y * rf), (int)std::floor(center.z * rf)};
            auto it = m->blocks.find(cid);
            if (it == m->blocks.end()) { it = m->blocks.emplace(cid, std::make_unique<Block>()).first; created[cid] = true; updated[cid] = false; }
            const int lx = vc.x - cid.x * 16, ly = vc.y - cid.y * 16, lz = vc.z - cid.z * 16;
            const int id = (lz * 16 + ly) * 16 + lx;
            if (!(id >= 0 && id < 4096)) continue;
            Block& B = *it->second;
            const V3 cc{inv[0] * center.x + (inv[1] * center.y + inv[2] * center.z) + tinv[0], inv[3] * center.x + (inv[4] * center.y + inv[5] * center.z) + tinv[1],
                        inv[6] * center.x + (inv[7] * center.y + inv[8] * center.z) + tinv[2]};
            const float length = std::sqrt(cc.x * cc.x + (cc.y * cc.y + cc.z * cc.z));
            const float u = length * (dpt / cc.z - 1);
            const float weight = P.weight / (2.0f * trunc);
            if (std::fabs(u) < trunc) {
                const float ow = B.w[id], os = B.sdf[id];
                B.sdf[id] = (ow * os + weight * u) / (weight + ow);
                B.w[id] = ow + weight;
                B.kfid[id] = g_cloud_kfids ? g_cloud_kfids[i] : g_cloud_kfid_all;         // distVoxel.SetKfid(kfid) (src/Chisel.cpp:534)
                if (rgb) {
                    uint8_t* cv = &B.rgba[(size_t)id * 4];
                    if (!(cv[3] >= 255 - 1)) {
                        const uint8_t nr = (uint8_t)(rgb[3 * i] * 255.0f), ng = (uint8_t)(rgb[3 * i + 1] * 255.0f), nb = (uint8_t)(rgb[3 * i + 2] * 255.0f);
                        const float invw = 1.f / (float)(1 + cv[3]);
                        cv[0] = (uint8_t)((float)(cv[3] * cv[0] + 1 * nr) * invw);
                        cv[1] = (uint8_t)((float)(cv[3] * cv[1] + 1 * ng) * invw);
                        cv[2] = (uint8_t)((float)(cv[3] * cv[2] + 1 * nb) * invw);
                        cv[3] = (uint8_t)(cv[3] + 1);
                    }
                }
                updated[cid] = true;
            }
        }
    }
    for (auto& kv : updated) if (kv.second) ++m->n_updated;
    for (auto& kv : created) {
        if (!updated[kv.first]) { m->blocks.erase(kv.first); ++m->n_collected; }
        else ++m->n_new;
    }
    m->n_range = n_carved;
    return 0;
}


// ---------------------------------------------------------------------------------------------
// §8f rank 3: ChunkManager::Deform (Thirdparty/open_chisel/src/ChunkManager.cpp:920-1062; ChiselServer::Deform, PointCloudMapChisel.cc:406-489).
// Every known voxel (weight > 1e-15) whose keyframe id has an entry in the deformation map moves to R * pos + t; the first voxel that lands
// in a new cell is copied (distance, weight, kfid, colour), later ones are folded in with DistVoxel::Integrate(sdf, weight) + SetKfid and
// ColorVoxel::Integrate(r, g, b, 1) (division form, saturated), in the order the reference visits them: its chunk map's iteration order, then
// voxel id.  That map is a std::unordered_map -- its order is a property of libstdc++ and of the whole insert/erase history -- so the order
// is an INPUT here (`order` = m chunk keys; chunks not listed follow in (x,y,z) key order; NULL = key order, what the product uses);
// tests/test_oracle_vs_reference_deform.py passes the compiled reference's own order and gets its result bit for bit.
// The two float roundings that pick the new chunk (floor(pos * 1/(16 res))) and the new voxel (floor(pos * 1/res) - 16 * chunk) can disagree
// at a chunk face; the reference then indexes outside the chunk (undefined behaviour): such voxels are dropped here.
// ---------------------------------------------------------------------------------------------
int orc_tsdf_deform(void* h, const uint32_t* kfids, const float* Rt /* n x 12: R row-major, then t in the 4th column */, int n, const int32_t* order, int m_order)
{
    Map* m = (Map*)h;
    const float res = m->p.voxel_resolution;
    const float half = res * 0.5f, inv_res = 1.f / res, rf = 1.0f / (16 * res);
    std::unordered_map<uint32_t, int> which;
    for (int i = 0; i < n; ++i) which[kfids[i]] = i;          // later entries of the same id win, like repeated operator[] assignments
    std::vector<Key> visit;
    std::unordered_map<Key, bool, KeyHash> listed;
    for (int i = 0; i < m_order; ++i) { const Key k{order[3 * i], order[3 * i + 1], order[3 * i + 2]}; if (m->blocks.count(k) && !listed[k]) { visit.push_back(k); listed[k] = true; } }
    {
        std::map<std::tuple<int, int, int>, Key> rest;
        for (auto& kv : m->blocks) if (!listed.count(kv.first)) rest[{kv.first.x, kv.first.y, kv.first.z}] = kv.first;
        for (auto& kv : rest) visit.push_back(kv.second);
    }
    std::unordered_map<Key, std::unique_ptr<Block>, KeyHash> fresh;
    int dropped = 0;
    for (const Key& k : visit) {
        const Block& B = *m->blocks[k];
        const V3 origin{(float)(16 * k.x) * res, (float)(16 * k.y) * res, (float)(16 * k.z) * res};
        int id = 0;
        for (int z = 0; z < 16; ++z) for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x, ++id) {
            if (B.w[id] <= 1e-15) continue;
            auto it = which.find(B.kfid[id]);
            if (it == which.end()) { ++dropped; continue; }
            const float* T = Rt + 12 * (size_t)it->second;
            const V3 pos = V3{(float)x * res + half, (float)y * res + half, (float)z * res + half} + origin;
            const V3 np{T[0] * pos.x + (T[1] * pos.y + T[2] * pos.z) + T[3], T[4] * pos.x + (T[5] * pos.y + T[6] * pos.z) + T[7],
                        T[8] * pos.x + (T[9] * pos.y + T[10] * pos.z) + T[11]};
            const Key nk{(int)std::floor(np.x * rf), (int)std::floor(np.y * rf), (int)std::floor(np.z * rf)};
            const int lx = (int)std::floor(np.x * inv_res) - 16 * nk.x, ly = (int)std::floor(np.y * inv_res) - 16 * nk.y, lz = (int)std::floor(np.z * inv_res) - 16 * nk.z;
            if (lx < 0 || lx > 15 || ly < 0 || ly > 15 || lz < 0 || lz > 15) { ++dropped; continue; }
            auto fit = fresh.find(nk);
            if (fit == fresh.end()) fit = fresh.emplace(nk, std::make_unique<Block>()).first;
            Block& N = *fit->second;
            const int nid = (lz * 16 + ly) * 16 + lx;
            if (N.w[nid] <= 1e-15) {
                N.sdf[nid] = B.sdf[id]; N.w[nid] = B.w[id]; N.kfid[nid] = B.kfid[id];
                std::memcpy(&N.rgba[(size_t)nid * 4], &B.rgba[(size_t)id * 4], 4);
            } else {
                const float ow = N.w[nid], os = N.sdf[nid];
                N.sdf[nid] = (ow * os + B.w[id] * B.sdf[id]) / (B.w[id] + ow);
                N.w[nid] = ow + B.w[id];
                N.kfid[nid] = B.kfid[id];
                if (m->p.use_color) {
                    uint8_t* cv = &N.rgba[(size_t)nid * 4];
                    const uint8_t* sc = &B.rgba[(size_t)id * 4];
                    if (!(cv[3] >= 255 - 1)) {             // ColorVoxel::Integrate(r, g, b, 1) (ColorVoxel.h:68-89)
                        const float wsum = (float)(1 + cv[3]);
                        for (int ch = 0; ch < 3; ++ch) {
                            const float upd = std::min(std::max((float)(cv[3] * (float)cv[ch] + 1 * sc[ch]) / wsum, 0.0f), 255.0f);
                            cv[ch] = (uint8_t)upd;
                        }
                        cv[3] = (uint8_t)(cv[3] + 1);
                    }
                }
            }
        }
    }
    m->blocks.swap(fresh);
    m->n_collected = dropped;
    return 0;
}

// ---------------------------------------------------------------------------------------------
// §8f rank 3 ("PLY load"): PointCloudMapChisel::LoadMap (src/PointCloudMapChisel.cc:527-549) reads the saved cloud with PCL (host I/O) and
// hands it to ChiselServer::IntegrateWorldPointCloud -> Chisel::IntegrateWorldPointCloudWithNormals (Thirdparty/open_chisel/src/Chisel.cpp:
// 238-379): per point a ray of +-4 voxels along its NORMAL, u = (centre - point) . dir, |u| < 4 res -> DistVoxel::Integrate(u, weight / (8 res)),
// SetKfid, ColorVoxel::Integrate(r, g, b, 1); no carving; new chunks that stayed untouched are collected.
// ---------------------------------------------------------------------------------------------
int orc_tsdf_integrate_world_cloud(void* h, const float* xyz, const float* rgb, const float* normals, const uint32_t* kfids, uint32_t kfid_all, int n, const float* Twc)
{
    Map* m = (Map*)h;
    const Params& P = m->p;
    const float res = P.voxel_resolution;
    const float r00 = Twc[0], r01 = Twc[1], r02 = Twc[2], r10 = Twc[4], r11 = Twc[5], r12 = Twc[6], r20 = Twc[8], r21 = Twc[9], r22 = Twc[10];
    const V3 t{Twc[3], Twc[7], Twc[11]};
    const float roundToVoxel = 1.0f / res, half = 0.5f * res, rf = 1.0f / (16 * res);
    const float truncation = 4 * res;
    std::unordered_map<Key, bool, KeyHash> updated, created;
    m->n_updated = m->n_new = m->n_collected = 0;
    std::vector<Key> walk;
    for (int i = 0; i < n; ++i) {
        const V3 p{xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]};
        const V3 wp{r00 * p.x + (r01 * p.y + r02 * p.z) + t.x, r10 * p.x + (r11 * p.y + r12 * p.z) + t.y, r20 * p.x + (r21 * p.y + r22 * p.z) + t.z};
        V3 dir{normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]};
        const float nn = dir.x * dir.x + (dir.y * dir.y + dir.z * dir.z);
        if (nn > 0) { const float sq = std::sqrt(nn); dir = V3{dir.x / sq, dir.y / sq, dir.z / sq}; }
        const V3 swp = wp * roundToVoxel;
        const V3 sdt = (dir * truncation) * roundToVoxel;
        const V3 start = swp - sdt, end = swp + sdt;
        walk.clear();
        {
            int x = (int)std::floor(start.x), y = (int)std::floor(start.y), z = (int)std::floor(start.z);
            const int endX = (int)std::floor(end.x), endY = (int)std::floor(end.y), endZ = (int)std::floor(end.z);
            const V3 direction = end - start;
            const float maxDist = direction.x * direction.x + (direction.y * direction.y + direction.z * direction.z);
            const float dx = (float)(endX - x), dy = (float)(endY - y), dz = (float)(endZ - z);
            const int stepX = (int)sgn_f((int)dx), stepY = (int)sgn_f((int)dy), stepZ = (int)sgn_f((int)dz);
            float tMaxX = intbound(start.x, (int)dx), tMaxY = intbound(start.y, (int)dy), tMaxZ = intbound(start.z, (int)dz);
            const float tDeltaX = ((float)stepX) / dx, tDeltaY = ((float)stepY) / dy, tDeltaZ = ((float)stepZ) / dz;
            if (!(stepX == 0 && stepY == 0 && stepZ == 0)) {
                for (;;) {
                    walk.push_back(Key{x, y, z});
                    const V3 d3{(float)x - start.x, (float)y - start.y, (float)z - start.z};
                    const float dist = d3.x * d3.x + (d3.y * d3.y + d3.z * d3.z);
                    if (dist > maxDist) break;
                    if (x == endX && y == endY && z == endZ) break;
                    if (tMaxX < tMaxY) { if (tMaxX < tMaxZ) { x += stepX; tMaxX += tDeltaX; } else { z += stepZ; tMaxZ += tDeltaZ; } }
                    else { if (tMaxY < tMaxZ) { y += stepY; tMaxY += tDeltaY; } else { z += stepZ; tMaxZ += tDeltaZ; } }
                }
            }
        }
        for (const Key& vc : walk) {
            const V3 center{(float)vc.x * res + half, (float)vc.y * res + half, (float)vc.z * res + half};
            const Key cid{(int)std::floor(center.x * rf), (int)std::floor(center.y * rf), (int)std::floor(center.z * rf)};
            auto it = m->blocks.find(cid);
            if (it == m->blocks.end()) { it = m->blocks.emplace(cid, std::make_unique<Block>()).first; created[cid] = true; updated[cid] = false; }
            const int lx = vc.x - cid.x * 16, ly = vc.y - cid.y * 16, lz = vc.z - cid.z * 16;
            const int id = (lz * 16 + ly) * 16 + lx;
            if (!(id >= 0 && id < 4096)) continue;
            Block& B = *it->second;
            const V3 dcp = center - wp;
            const float u = dcp.x * dir.x + (dcp.y * dir.y + dcp.z * dir.z);
            const float weight = P.weight / (2.0f * truncation);
            if (std::fabs(u) < truncation) {
                const float ow = B.w[id], os = B.sdf[id];
                B.sdf[id] = (ow * os + weight * u) / (weight + ow);
                B.w[id] = ow + weight;
                B.kfid[id] = kfids ? kfids[i] : kfid_all;
                uint8_t* cv = &B.rgba[(size_t)id * 4];
                if (P.use_color && !(cv[3] >= 255 - 1)) {
                    const uint8_t nc[3] = {(uint8_t)((rgb ? rgb[3 * i] : 0.f) * 255.0f), (uint8_t)((rgb ? rgb[3 * i + 1] : 0.f) * 255.0f), (uint8_t)((rgb ? rgb[3 * i + 2] : 0.f) * 255.0f)};
                    const float wsum = (float)(1 + cv[3]);
                    for (int ch = 0; ch < 3; ++ch) cv[ch] = (uint8_t)std::min(std::max((float)(cv[3] * (float)cv[ch] + 1 * nc[ch]) / wsum, 0.0f), 255.0f);
                    cv[3] = (uint8_t)(cv[3] + 1);
                }
                updated[cid] = true;
            }
        }
    }
    for (auto& kv : updated) if (kv.second) ++m->n_updated;
    for (auto& kv : created) {
        if (!updated[kv.first]) { m->blocks.erase(kv.first); ++m->n_collected; }
        else ++m->n_new;
    }
    return 0;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// SURVEY.md §8f rank 3 -- read-out: ChunkManager::RecomputeMesh (Thirdparty/open_chisel/src/ChunkManager.cpp:116-172) for every chunk:
//   GenerateMesh (:577-664: inside voxels in z,y,x order, then the max-X, max-Y and max-Z planes, whose cube corners reach into the
//   +x/+y/+z neighbour chunks, ExtractInside/BorderVoxelMeshKfid :438-575), MarchingCubes::MeshCube / InterpolateEdgeVertices /
//   InterpolateVertex (include/open_chisel/marching_cubes/MarchingCubes.h:76-271: triangle vertices in reversed table order, face
//   normal, `vertex1 + 0.5f * vertex2` for |sdf1 - sdf2| < 1e-6 as written), ColorizeMesh / InterpolateColor (:715-806, INCLUDING its
//   quirk: the eight GetColorVoxel look-ups are made with integer voxel indices passed as metric positions, so in practice the
//   nearest-voxel fallback Chunk::GetColorAt (src/Chunk.cpp:136-155) supplies the colour), ComputeNormalsFromGradients /
//   GetSDFAndGradient / GetSDF (:666-713, :838-856).  Stateless: the reference re-meshes the 27-neighbourhood of every updated chunk
//   after each integration (Chisel.h:108-118, Chisel::UpdateMeshes), which leaves exactly the meshes a full pass over the current voxels
//   produces; tests/test_oracle_vs_reference_mesh.py checks both flows of the compiled reference against this function.
//   kfids are not carried (the depth-scan path never sets them).  PINNED bit-exactly (vertex order, positions, normals, colours).
// ---------------------------------------------------------------------------------------------
namespace {

const uint64_t kTriTable[256] = {
#include "../plvs_b200/csrc/mc_tables.inc"
};
const int kEdgePairs[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
const int kCubeOff[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};

struct MeshCtx {
    const Map* m; float res, inv, half, rounding;
    const Block* find(int x, int y, int z) const { auto it = m->blocks.find(Key{x, y, z}); return it == m->blocks.end() ? nullptr : it->second.get(); }
    // ChunkManager::GetIDAt (ChunkManager.h:192-201)
    Key id_at(V3 p) const { return Key{(int)std::floor(p.x * rounding), (int)std::floor(p.y * rounding), (int)std::floor(p.z * rounding)}; }
    static V3 origin(const Key& k, float res) { return V3{(float)(16 * k.x) * res, (float)(16 * k.y) * res, (float)(16 * k.z) * res}; }
    // chunk->GetVoxelID(rel) with Chunk::GetVoxelCoords (src/Chunk.cpp:80-94): no range check on the coordinates, only on the id
    long voxel_id(V3 rel) const
    {
        const int x = (int)std::floor(rel.x * inv), y = (int)std::floor(rel.y * inv), z = (int)std::floor(rel.z * inv);
        return ((long)z * 16 + y) * 16 + x;
    }
    // ChunkManager::GetSDF (:692-713)
    bool sdf_at(V3 posf, double* dist) const
    {
        const Key k = id_at(posf);
        const Block* b = find(k.x, k.y, k.z);
        if (!b) return false;
        const long id = voxel_id(posf - origin(k, res));
        if (id >= 0 && id < 4096 && b->w[id] > 1e-12) { *dist = b->sdf[id]; return true; }
        return false;
    }
    // ChunkManager::GetColorVoxel (:822-841)
    const uint8_t* color_voxel(V3 pos) const
    {
        const Key k = id_at(pos);
        const Block* b = find(k.x, k.y, k.z);
        if (!b) return nullptr;
        const long id = voxel_id(pos - origin(k, res));
        return (id >= 0 && id < 4096) ? &b->rgba[4 * id] : nullptr;
    }
    V3 interpolate_color(V3 p) const
    {
        const float x = p.x, y = p.y, z = p.z;
        const int x_0 = (int)std::floor(x * inv), y_0 = (int)std::floor(y * inv), z_0 = (int)std::floor(z * inv);
        const int x_1 = x_0 + 1, y_1 = y_0 + 1, z_1 = z_0 + 1;
        const uint8_t* v_000 = color_voxel(V3{(float)x_0, (float)y_0, (float)z_0});
        const uint8_t* v_001 = color_voxel(V3{(float)x_0, (float)y_0, (float)z_1});
        const uint8_t* v_011 = color_voxel(V3{(float)x_0, (float)y_1, (float)z_1});
        const uint8_t* v_111 = color_voxel(V3{(float)x_1, (float)y_1, (float)z_1});
        const uint8_t* v_110 = color_voxel(V3{(float)x_1, (float)y_1, (float)z_0});
        const uint8_t* v_100 = color_voxel(V3{(float)x_1, (float)y_0, (float)z_0});
        const uint8_t* v_010 = color_voxel(V3{(float)x_0, (float)y_1, (float)z_0});
        const uint8_t* v_101 = color_voxel(V3{(float)x_1, (float)y_0, (float)z_1});
        if (!v_000 || !v_001 || !v_011 || !v_111 || !v_110 || !v_100 || !v_010 || !v_101) {
            const Key k = id_at(p);
            const Block* b = find(k.x, k.y, k.z);
            if (!b) return V3{0, 0, 0};
            // Chunk::GetColorAt (src/Chunk.cpp:136-155)
            const V3 o = origin(k, res);
            const float size = (float)16 * res;
            if (p.x >= o.x && p.y >= o.y && p.z >= o.z && p.x <= o.x + size && p.y <= o.y + size && p.z <= o.z + size) {
                const V3 cp = (p - o) * inv;
                const int cx = (int)cp.x, cy = (int)cp.y, cz = (int)cp.z;
                if (cx >= 0 && cx < 16 && cy >= 0 && cy < 16 && cz >= 0 && cz < 16) {
                    const uint8_t* c = &b->rgba[4 * ((cz * 16 + cy) * 16 + cx)];
                    const float invMax = 1.f / 255.f;
                    return V3{(float)c[0] * invMax, (float)c[1] * invMax, (float)c[2] * invMax};
                }
            }
            return V3{0, 0, 0};
        }
        const float xd = (x - x_0) / (x_1 - x_0), yd = (y - y_0) / (y_1 - y_0), zd = (z - z_0) / (z_1 - z_0);
        float out[3];
        for (int ch = 0; ch < 3; ++ch) {
            const float c_00 = v_000[ch] * (1 - xd) + v_100[ch] * xd;
            const float c_10 = v_010[ch] * (1 - xd) + v_110[ch] * xd;
            const float c_01 = v_001[ch] * (1 - xd) + v_101[ch] * xd;
            const float c_11 = v_011[ch] * (1 - xd) + v_111[ch] * xd;
            const float c_0 = c_00 * (1 - yd) + c_10 * yd;
            const float c_1 = c_01 * (1 - yd) + c_11 * yd;
            const float c = c_0 * (1 - zd) + c_1 * zd;
            out[ch] = c / 255.0f;
        }
        return V3{out[0], out[1], out[2]};
    }
    // ChunkManager::GetSDFAndGradient (:666-690) + the renormalisation of ComputeNormalsFromGradients (:838-856)
    bool gradient_normal(V3 pos, V3* n) const
    {
        const V3 posf{std::floor(pos.x * inv) * res + half, std::floor(pos.y * inv) * res + half, std::floor(pos.z * inv) * res + half};
        double dist, xp, yp, zp, xm, ym, zm;
        if (!sdf_at(posf, &dist)) return false;
        if (!sdf_at(posf + V3{res, 0, 0}, &xp)) return false;
        if (!sdf_at(posf + V3{0, res, 0}, &yp)) return false;
        if (!sdf_at(posf + V3{0, 0, res}, &zp)) return false;
        if (!sdf_at(posf - V3{res, 0, 0}, &xm)) return false;
        if (!sdf_at(posf - V3{0, res, 0}, &ym)) return false;
        if (!sdf_at(posf - V3{0, 0, res}, &zm)) return false;
        V3 g{(float)(xp - xm), (float)(yp - ym), (float)(zp - zm)};
        const float sq = dot(g, g);
        if (sq > 0.f) { const float s = std::sqrt(sq); g = V3{g.x / s, g.y / s, g.z / s}; }      // Eigen normalize()
        const float mag = std::sqrt(dot(g, g));
        if (mag > 1e-12) { *n = g * (1.0f / mag); return true; }
        return false;
    }
};

void mesh_voxel(const MeshCtx& c, const Key& key, const Block* blk, int ix, int iy, int iz, bool border, std::vector<V3>& verts, std::vector<V3>& normals,
                std::vector<uint32_t>& kfids)
{
    const float res = c.res;
    const V3 halfVoxel = V3{res, res, res} * 0.5f;
    const V3 centroid = V3{(float)ix, (float)iy, (float)iz} * res + halfVoxel;
    const V3 coords = centroid + MeshCtx::origin(key, res);
    V3 cc[8]; float sdf[8];
    uint32_t corner_kfid = 0;            // USE_KFID_VERTICES == 0: the cube takes the keyframe id of its first corner (ChunkManager.cpp:464-468)
    for (int i = 0; i < 8; ++i) {
        int x = ix + kCubeOff[i][0], y = iy + kCubeOff[i][1], z = iz + kCubeOff[i][2];
        const Block* b = blk;
        if (border && !(x < 16 && y < 16 && z < 16)) {
            int off[3] = {0, 0, 0};
            if (x >= 16) { off[0] = 1; x = 0; }
            if (y >= 16) { off[1] = 1; y = 0; }
            if (z >= 16) { off[2] = 1; z = 0; }
            b = c.find(key.x + off[0], key.y + off[1], key.z + off[2]);
            if (!b) return;
        }
        const int id = (z * 16 + y) * 16 + x;
        if (b->w[id] <= 1e-15) return;
        cc[i] = coords + V3{(float)kCubeOff[i][0] * res, (float)kCubeOff[i][1] * res, (float)kCubeOff[i][2] * res};
        sdf[i] = b->sdf[id];
        if (i == 0) corner_kfid = b->kfid[id];
    }
    int cfg = 0;
    for (int i = 0; i < 8; ++i) if (sdf[i] < 0) cfg |= 1 << i;
    if (cfg == 0) return;
    V3 edge[12];
    for (int e = 0; e < 12; ++e) {
        const int a = kEdgePairs[e][0], b = kEdgePairs[e][1];
        if ((sdf[a] < 0 && sdf[b] >= 0) || (sdf[a] >= 0 && sdf[b] < 0)) {
            const float diff = sdf[a] - sdf[b];
            if (std::fabs(diff) < 1e-6f) edge[e] = cc[a] + cc[b] * 0.5f;
            else { const float t = sdf[a] / diff; edge[e] = cc[a] + (cc[b] - cc[a]) * t; }
        }
    }
    const uint64_t row = kTriTable[cfg];
    const int ntri = (int)(row >> 60);
    for (int t = 0; t < ntri; ++t) {
        const V3 p0 = edge[(row >> (4 * (3 * t + 2))) & 15], p1 = edge[(row >> (4 * (3 * t + 1))) & 15], p2 = edge[(row >> (4 * (3 * t))) & 15];
        V3 n = cross(p1 - p0, p2 - p0);
        const float sq = dot(n, n);
        if (sq > 0.f) { const float s = std::sqrt(sq); n = V3{n.x / s, n.y / s, n.z / s}; }
        verts.push_back(p0); verts.push_back(p1); verts.push_back(p2);
        normals.push_back(n); normals.push_back(n); normals.push_back(n);
        kfids.push_back(corner_kfid); kfids.push_back(corner_kfid); kfids.push_back(corner_kfid);
    }
}

}  // namespace

extern "C" {

// meshes of all chunks in (x,y,z) key order; chunks without triangles are left out like ChunkManager::RecomputeMesh does (:165-167).
// keys[3*i], counts[i] = vertices of mesh i; verts / normals / colors = 3 floats per vertex, concatenated.  Returns the number of
// meshes, *total_verts the number of vertices; arrays may be null (counting pass) and are filled up to the caps.
static uint32_t* g_mesh_kfids = nullptr;          // optional per-vertex keyframe ids of the next orc_tsdf_extract_mesh (Mesh::kfids)

int orc_tsdf_extract_mesh(void* h, int32_t* keys, int32_t* counts, int cap_meshes, float* verts, float* normals, float* colors, long cap_verts, long* total_verts);
int orc_tsdf_extract_mesh_kfids(void* h, uint32_t* kfids, long cap_verts)
{
    g_mesh_kfids = kfids;
    long total = 0;
    std::vector<float> dummy;
    const int nm = orc_tsdf_extract_mesh(h, nullptr, nullptr, 0, nullptr, nullptr, nullptr, cap_verts, &total);
    g_mesh_kfids = nullptr;
    return nm;
}

int orc_tsdf_extract_mesh(void* h, int32_t* keys, int32_t* counts, int cap_meshes, float* verts, float* normals, float* colors, long cap_verts, long* total_verts)
{
    Map* m = (Map*)h;
    MeshCtx c{m, m->p.voxel_resolution, 1.f / m->p.voxel_resolution, 0.5f * m->p.voxel_resolution, 1.0f / ((float)16 * m->p.voxel_resolution)};
    std::map<std::tuple<int, int, int>, const Block*> ord;
    for (auto& kv : m->blocks) ord[{kv.first.x, kv.first.y, kv.first.z}] = kv.second.get();
    int nm = 0; long nv = 0;
    std::vector<V3> V, N;
    std::vector<uint32_t> Kf;
    for (auto& kv : ord) {
        const Key key{std::get<0>(kv.first), std::get<1>(kv.first), std::get<2>(kv.first)};
        const Block* b = kv.second;
        V.clear(); N.clear(); Kf.clear();
        for (int z = 0; z < 15; ++z) for (int y = 0; y < 15; ++y) for (int x = 0; x < 15; ++x) mesh_voxel(c, key, b, x, y, z, false, V, N, Kf);
        for (int z = 0; z < 15; ++z) for (int y = 0; y < 16; ++y) mesh_voxel(c, key, b, 15, y, z, true, V, N, Kf);
        for (int z = 0; z < 15; ++z) for (int x = 0; x < 15; ++x) mesh_voxel(c, key, b, x, 15, z, true, V, N, Kf);
        for (int y = 0; y < 16; ++y) for (int x = 0; x < 16; ++x) mesh_voxel(c, key, b, x, y, 15, true, V, N, Kf);
        if (V.empty()) continue;
        if (nm < cap_meshes) {
            if (keys) { keys[3 * nm] = key.x; keys[3 * nm + 1] = key.y; keys[3 * nm + 2] = key.z; }
            if (counts) counts[nm] = (int)V.size();
        }
        for (size_t i = 0; i < V.size(); ++i) {
            V3 col{0, 0, 0};
            if (m->p.use_color) col = c.interpolate_color(V[i]);
            V3 n = N[i];
            c.gradient_normal(V[i], &n);
            if (nv < cap_verts) {
                if (g_mesh_kfids) g_mesh_kfids[nv] = Kf[i];
                if (verts) { verts[3 * nv] = V[i].x; verts[3 * nv + 1] = V[i].y; verts[3 * nv + 2] = V[i].z; }
                if (normals) { normals[3 * nv] = n.x; normals[3 * nv + 1] = n.y; normals[3 * nv + 2] = n.z; }
                if (colors) { colors[3 * nv] = col.x; colors[3 * nv + 1] = col.y; colors[3 * nv + 2] = col.z; }
            }
            ++nv;
        }
        ++nm;
    }
    if (total_verts) *total_verts = nv;
    return nm;
}

}  // extern "C"
