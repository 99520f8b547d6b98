// TEST INFRASTRUCTURE ONLY -- C driver around the REFERENCE's own open_chisel sources.
//
// oracle/ref_build.py compiles /root/reference/Thirdparty/open_chisel/src/**/*.cpp where they lie (nothing is copied)
// together with this file into oracle/_ref/libchisel_ref.so.  Eigen is absent from the image, so the sources are
// compiled against oracle/eigen_standin/ (see its header for exactly what is restated there).  This file only does
// what chisel_server::ChiselServer does around the library (ChiselServer.cpp is PCL-bound and cannot be compiled):
//   ctor                         Thirdparty/chisel_server/src/ChiselServer.cpp:100-110   (Chisel(16^3, res, useColor) +
//                                SetupProjectionIntegrator :623-631 incl. the static_cast<uint16_t>(weight) of :108)
//   SetDepthCameraInfo           ChiselServer.cpp:444-450 + Conversions.h:494-506 (near/far planes from the params)
//   IntegrateLastDepthImage      ChiselServer.cpp:632-662 (colour -> IntegrateDepthScanColorWithOneCameraModelBGR,
//                                else IntegrateDepthScan)
//   IntegrateLastPointCloud      ChiselServer.cpp:664-705 (IntegratePointCloudWidthDepth)
// and exposes the same C entry points as oracle/tsdf_oracle.cpp with the prefix ref_ so one Python class drives both.
#include <open_chisel/Chisel.h>
#include <open_chisel/ProjectionIntegrator.h>
#include <open_chisel/truncation/QuadraticTruncator.h>
#include <open_chisel/weighting/ConstantWeighter.h>
#include <open_chisel/camera/DepthImage.h>
#include <open_chisel/camera/ColorImage.h>
#include <open_chisel/camera/PinholeCamera.h>
#include <open_chisel/pointcloud/PointCloud.h>

#include <cstdint>
#include <cstdio>
#include <csignal>
#include <sys/wait.h>
#include <unistd.h>
#include <cstring>
#include <map>
#include <memory>
#include <tuple>

namespace {

struct Params {
    float voxel_resolution, trunc_quad, trunc_linear, trunc_const, trunc_scale, weight;
    int32_t use_carving; float carving_dist; int32_t use_color; float near_plane, far_plane; int32_t max_blocks;
};

struct Ref {
    Params p;
    std::unique_ptr<chisel::Chisel> map;
    chisel::ProjectionIntegrator integrator;
    chisel::PinholeCamera depthCam, colorCam;
    bool got_camera = false;
    int n_range = 0;
};

void make_map(Ref* r)
{
    const Params& p = r->p;
    r->map.reset(new chisel::Chisel(Eigen::Vector3i(16, 16, 16), p.voxel_resolution, p.use_color != 0));
    chisel::Vec4 truncation(p.trunc_quad, p.trunc_linear, p.trunc_const, p.trunc_scale);
    const uint16_t weight = static_cast<uint16_t>(p.weight);
    r->integrator.SetCentroids(r->map->GetChunkManager().GetCentroids());
    r->integrator.SetTruncator(chisel::TruncatorPtr(new chisel::QuadraticTruncator(truncation(0), truncation(1), truncation(2), truncation(3))));
    r->integrator.SetWeighter(chisel::WeighterPtr(new chisel::ConstantWeighter(weight)));
    r->integrator.SetCarvingDist(p.carving_dist);
    r->integrator.SetCarvingEnabled(p.use_carving != 0);
}

chisel::PinholeCamera to_camera(double fx, double fy, double cx, double cy, int w, int h)
{
    chisel::PinholeCamera cam;
    chisel::Intrinsics in;
    in.SetFx(fx); in.SetFy(fy); in.SetCx(cx); in.SetCy(cy);
    cam.SetIntrinsics(in);
    cam.SetWidth(w); cam.SetHeight(h);
    return cam;
}

chisel::Transform to_transform(const float* Twc)
{
    chisel::Transform T = chisel::Transform::Identity();
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) T(i, j) = Twc[4 * i + j]; T(i, 3) = Twc[4 * i + 3]; }
    return T;
}

}  // namespace

extern "C" {

void* ref_tsdf_create(const Params* p, int /*threads*/)
{
    Ref* r = new Ref();
    r->p = *p;
    make_map(r);
    return r;
}
void ref_tsdf_destroy(void* h) { delete (Ref*)h; }
void ref_tsdf_reset(void* h) { ((Ref*)h)->map->Reset(); }

void ref_tsdf_set_camera(void* h, double fx, double fy, double cx, double cy, int w, int ht)
{
    Ref* r = (Ref*)h;
    r->depthCam = to_camera(fx, fy, cx, cy, w, ht);
    r->depthCam.SetNearPlane(r->p.near_plane); r->depthCam.SetFarPlane(r->p.far_plane);
    r->colorCam = r->depthCam;
    r->got_camera = true;
}

int ref_tsdf_integrate(void* h, const float* depth, int w, int ht, const uint8_t* bgr, int nch, const float* Twc, int mode)
{
    Ref* r = (Ref*)h;
    if (!r->got_camera || !depth) return -5;
    std::shared_ptr<chisel::DepthImage<float>> d(new chisel::DepthImage<float>(w, ht));
    std::memcpy(d->GetMutableData(), depth, sizeof(float) * (size_t)w * ht);
    const chisel::Transform T = to_transform(Twc);
    {   // statistic only: size of the reference's chunk list for this scan (same calls as the integrate functions make)
        chisel::Frustum fr; chisel::PinholeCamera c = r->depthCam;
        if (mode == 0) { float mn, mx, mean; d->GetStats(mn, mx, mean); c.SetNearPlane(mn); c.SetFarPlane(mx); }
        c.SetupFrustum(T, &fr);
        chisel::ChunkIDList ids; r->map->GetMutableChunkManager().GetChunkIDsIntersecting(fr, &ids);
        r->n_range = (int)ids.size();
    }
    if (mode == 1) {
        std::shared_ptr<chisel::ColorImage<uint8_t>> c(new chisel::ColorImage<uint8_t>(w, ht, nch));
        std::memcpy(c->GetMutableData(), bgr, (size_t)w * ht * nch);
        r->map->IntegrateDepthScanColorWithOneCameraModelBGR<float, uint8_t>(r->integrator, d, T, r->depthCam, c, T, r->colorCam);
    } else {
        r->map->IntegrateDepthScan<float>(r->integrator, d, T, r->depthCam);
    }
    return 0;
}

static const uint32_t* g_kfids = nullptr;
static uint32_t g_kfid_all = 0;

int ref_tsdf_integrate_cloud(void* h, const float* xyz, const float* rgb, int n, const float* depth, int w, int ht, const float* Twc);
// the same with the cloud's keyframe ids (PointCloud::GetKfids): kfids[n], or NULL for one id for every point
int ref_tsdf_integrate_cloud_kf(void* h, const float* xyz, const float* rgb, const uint32_t* kfids, uint32_t kfid_all, int n, const float* depth, int w, int ht, const float* Twc)
{
    g_kfids = kfids; g_kfid_all = kfid_all;
    const int rc = ref_tsdf_integrate_cloud(h, xyz, rgb, n, depth, w, ht, Twc);
    g_kfids = nullptr; g_kfid_all = 0;
    return rc;
}

int ref_tsdf_integrate_cloud(void* h, const float* xyz, const float* rgb, int n, const float* depth, int w, int ht, const float* Twc)
{
    Ref* r = (Ref*)h;
    chisel::PointCloud cloud;
    for (int i = 0; i < n; ++i) {
        cloud.AddPoint(chisel::Vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
        cloud.AddColor(rgb ? chisel::Vec3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]) : chisel::Vec3(0.f, 0.f, 0.f));
        cloud.GetMutableKfids().push_back(g_kfids ? g_kfids[i] : g_kfid_all);
    }
    std::shared_ptr<chisel::DepthImage<float>> d;
    if (depth) { d.reset(new chisel::DepthImage<float>(w, ht)); std::memcpy(d->GetMutableData(), depth, sizeof(float) * (size_t)w * ht); }
    const chisel::Transform T = to_transform(Twc);
    const bool carve = r->integrator.IsCarvingEnabled();
    if (!(depth && r->got_camera)) r->integrator.SetCarvingEnabled(false);      // the oracle/product skip carving without a depth image
    r->map->IntegratePointCloudWidthDepth<float>(r->integrator, cloud, T, d, r->depthCam, r->p.far_plane);
    r->integrator.SetCarvingEnabled(carve);
    return 0;
}

void ref_tsdf_stats(void* h, int32_t* out)
{
    Ref* r = (Ref*)h;
    out[0] = (int)r->map->GetChunkManager().GetChunks().size(); out[1] = r->n_range; out[2] = out[3] = out[4] = -1;
}

int ref_tsdf_download(void* h, int32_t* keys, float* sdf, float* weight, uint8_t* rgba, int cap)
{
    Ref* r = (Ref*)h;
    std::map<std::tuple<int, int, int>, chisel::ChunkPtr> ord;
    for (auto& kv : r->map->GetChunkManager().GetChunks()) ord[{kv.first(0), kv.first(1), kv.first(2)}] = kv.second;
    int n = 0;
    for (auto& kv : ord) {
        if (n >= cap) break;
        const chisel::Chunk& c = *kv.second;
        if (keys) { keys[3 * n] = std::get<0>(kv.first); keys[3 * n + 1] = std::get<1>(kv.first); keys[3 * n + 2] = std::get<2>(kv.first); }
        for (int i = 0; i < 4096; ++i) {
            const chisel::DistVoxel& v = c.GetDistVoxel(i);
            if (sdf) sdf[(size_t)n * 4096 + i] = v.GetSDF();
            if (weight) weight[(size_t)n * 4096 + i] = v.GetWeight();
            if (rgba) {
                uint8_t* o = rgba + ((size_t)n * 4096 + i) * 4;
                if (c.HasColors()) { const chisel::ColorVoxel& cv = c.GetColorVoxel(i); o[0] = cv.GetRed(); o[1] = cv.GetGreen(); o[2] = cv.GetBlue(); o[3] = cv.GetWeight(); }
                else o[0] = o[1] = o[2] = o[3] = 0;
            }
        }
        ++n;
    }
    return (int)ord.size();
}

// Chisel::UpdateMeshes (src/Chisel.cpp:57-65: the chunks flagged by the integrations since the last call) or, with all != 0,
// ChunkManager::RecomputeMeshes over every chunk
void ref_tsdf_update_meshes(void* h, int all)
{
    Ref* r = (Ref*)h;
    if (!all) { r->map->UpdateMeshes(); return; }
    chisel::ChunkSet every;
    for (auto& kv : r->map->GetChunkManager().GetChunks()) every[kv.first] = true;
    r->map->GetMutableChunkManager().RecomputeMeshes(every);
}

// ChunkManager::GetAllMeshes, non-empty ones, in (x,y,z) key order; same layout as orc_tsdf_extract_mesh
int ref_tsdf_mesh_download(void* h, int32_t* keys, int32_t* counts, int cap_meshes, float* verts, float* normals, float* colors, long cap_verts, long* total_verts)
{
    Ref* r = (Ref*)h;
    std::map<std::tuple<int, int, int>, chisel::MeshPtr> ord;
    for (auto& kv : r->map->GetChunkManager().GetAllMeshes())
        if (kv.second && !kv.second->vertices.empty()) ord[{kv.first(0), kv.first(1), kv.first(2)}] = kv.second;
    int nm = 0; long nv = 0;
    for (auto& kv : ord) {
        const chisel::Mesh& m = *kv.second;
        if (nm < cap_meshes) {
            if (keys) { keys[3 * nm] = std::get<0>(kv.first); keys[3 * nm + 1] = std::get<1>(kv.first); keys[3 * nm + 2] = std::get<2>(kv.first); }
            if (counts) counts[nm] = (int)m.vertices.size();
        }
        for (size_t i = 0; i < m.vertices.size(); ++i) {
            if (nv < cap_verts) {
                for (int k = 0; k < 3; ++k) {
                    if (verts) verts[3 * nv + k] = m.vertices[i](k);
                    if (normals) normals[3 * nv + k] = m.normals[i](k);
                    if (colors) colors[3 * nv + k] = m.HasColors() ? m.colors[i](k) : 0.f;
                }
            }
            ++nv;
        }
        ++nm;
    }
    if (total_verts) *total_verts = nv;
    return nm;
}

// Chisel::SaveAllMeshesToPLY (src/Chisel.cpp:79-118).  SaveMeshPLYASCII (src/io/PLY.cpp:29-86) flows off its end without a return statement, which
// g++ turns into a trap instruction: the file is complete by then (every line ends with std::endl), so the call runs in a forked child whose
// death by a signal is expected.  Returns the child's wait status.
int ref_tsdf_save_ply(void* h, const char* path)
{
    std::fflush(nullptr);
    const pid_t pid = fork();
    if (pid == 0) { for (int sg : {SIGILL, SIGABRT, SIGSEGV, SIGBUS, SIGFPE}) signal(sg, SIG_DFL); ((Ref*)h)->map->SaveAllMeshesToPLY(path); _exit(0); }
    int status = 0;
    waitpid(pid, &status, 0);
    return status;
}

// DistVoxel::GetKfid of every voxel, chunks in the order of ref_tsdf_download
int ref_tsdf_download_kfid(void* h, uint32_t* kfid, int cap)
{
    Ref* r = (Ref*)h;
    std::map<std::tuple<int, int, int>, chisel::ChunkPtr> ord;
    for (auto& kv : r->map->GetChunkManager().GetChunks()) ord[{kv.first(0), kv.first(1), kv.first(2)}] = kv.second;
    int n = 0;
    for (auto& kv : ord) {
        if (n >= cap) break;
        for (int i = 0; i < 4096; ++i) kfid[(size_t)n * 4096 + i] = kv.second->GetDistVoxel(i).GetKfid();
        ++n;
    }
    return (int)ord.size();
}

// Mesh::kfids of the non-empty meshes, concatenated in the order of ref_tsdf_mesh_download
int ref_tsdf_extract_mesh_kfids(void* h, uint32_t* kfids, long cap_verts)
{
    Ref* r = (Ref*)h;
    std::map<std::tuple<int, int, int>, chisel::MeshPtr> ord;
    for (auto& kv : r->map->GetChunkManager().GetAllMeshes())
        if (kv.second && !kv.second->vertices.empty()) ord[{kv.first(0), kv.first(1), kv.first(2)}] = kv.second;
    long nv = 0;
    for (auto& kv : ord) for (size_t i = 0; i < kv.second->vertices.size(); ++i) { if (nv < cap_verts) kfids[nv] = kv.second->kfids[i]; ++nv; }
    return (int)ord.size();
}

// The chunk map's iteration order (std::unordered_map<ChunkID, ChunkPtr, ChunkHasher>): what ChunkManager::Deform walks.  keys may be NULL.
int ref_tsdf_chunk_order(void* h, int32_t* keys, int cap)
{
    Ref* r = (Ref*)h;
    int n = 0;
    for (auto& kv : r->map->GetChunkManager().GetChunks()) {
        if (keys && n < cap) { keys[3 * n] = kv.first(0); keys[3 * n + 1] = kv.first(1); keys[3 * n + 2] = kv.first(2); }
        ++n;
    }
    return n;
}

// ChiselServer::Deform (ChiselServer.cpp:616-620) -> Chisel::Deform -> ChunkManager::Deform.  Rt: n x 12 (R row-major, t in the 4th column).
int ref_tsdf_deform(void* h, const uint32_t* kfids, const float* Rt, int n)
{
    Ref* r = (Ref*)h;
    chisel::MapKfidRt map;
    for (int i = 0; i < n; ++i) {
        chisel::TransformRt T;
        for (int a = 0; a < 3; ++a) { for (int b = 0; b < 3; ++b) T.R(a, b) = Rt[12 * i + 4 * a + b]; T.t(a) = Rt[12 * i + 4 * a + 3]; }
        map[kfids[i]] = T;
    }
    r->map->Deform(map);
    return 0;
}

// ChiselServer::IntegrateWorldPointCloud (ChiselServer.cpp:588-614) -> Chisel::IntegrateWorldPointCloudWithNormals
int ref_tsdf_integrate_world_cloud(void* h, const float* xyz, const float* rgb, const float* normals, const uint32_t* kfids, uint32_t kfid_all, int n, const float* Twc)
{
    Ref* r = (Ref*)h;
    chisel::PointCloud cloud;
    for (int i = 0; i < n; ++i) {
        cloud.AddPoint(chisel::Vec3(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]));
        cloud.AddColor(rgb ? chisel::Vec3(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]) : chisel::Vec3(0.f, 0.f, 0.f));
        cloud.GetMutableNormals().push_back(chisel::Vec3(normals[3 * i], normals[3 * i + 1], normals[3 * i + 2]));
        cloud.GetMutableKfids().push_back(kfids ? kfids[i] : kfid_all);
    }
    const chisel::Transform T = to_transform(Twc);
    r->map->IntegrateWorldPointCloudWithNormals(r->integrator, cloud, T, r->p.far_plane);
    return 0;
}

}  // extern "C"
