// Runtime harness of the C++ drop-in shims (shim/plvs_shim.hpp): the reference-facing class surfaces PLVS2::ORBextractor (include/ORBextractor.h:86),
// PLVS2::ORBmatcher (include/ORBmatcher.h:68-97) and chisel_server::ChiselServer (Thirdparty/chisel_server/include/chisel_server/ChiselServer.h:190-286)
// are driven with real data through stand-in Frame / MapPoint / KeyFrame objects, the way Tracking / LocalMapping / PointCloudMapping call them, and
// what they leave behind (mvpMapPoints, vMatchedPairs, the voxel map, the point cloud) is handed back as flat arrays so tests/test_shim_runtime.py can
// compare it with the Python mirror and the oracle.  Test infrastructure: OpenCV C++ / Eigen / Sophus are absent in this image, so the object types are
// the minimal ones below (exactly the members the shim touches); the arithmetic the shim performs on them (pose * point, camera projection) is plain
// float code the test reproduces in numpy float32.
#define PLVS_SHIM_STANDIN
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <tuple>
#include <unordered_map>
#include <vector>
#include "../../shim/standin.hpp"

struct V3 { float v[3]; float operator()(int i) const { return v[i]; } };
struct V2 { float v[2]; float operator()(int i) const { return v[i]; } };
struct Pose {                       // translation-only rigid transform (enough to drive bForward / bBackward and Tcw * x)
    V3 t{};
    Pose inverse() const { Pose p; p.t = V3{{-t.v[0], -t.v[1], -t.v[2]}}; return p; }
    V3 translation() const { return t; }
    V3 operator*(const V3& p) const { return V3{{p.v[0] + t.v[0], p.v[1] + t.v[1], p.v[2] + t.v[2]}}; }
};
struct Camera {
    float fx = 500, fy = 500, cx = 320, cy = 240;
    V2 project(const V3& p) const { return V2{{fx * p.v[0] / p.v[2] + cx, fy * p.v[1] / p.v[2] + cy}}; }
};
struct KeyFrame;
struct MapPoint {
    bool mbTrackInView = true; float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackDepth = 1, mTrackViewCos = 1; int mnTrackScaleLevel = 0;
    cv::Mat desc{1, 32, CV_8U}; V3 pos{{0, 0, 2}};
    bool bad = false; int nobs = 1;
    bool isBad() const { return bad; }
    int Observations() const { return nobs; }
    cv::Mat GetDescriptor() const { return desc; }
    V3 GetWorldPos() const { return pos; }
    V3 GetNormal() const { return V3{{0, 0, 1}}; }
    float GetMinDistanceInvariance() const { return 0.f; }
    float GetMaxDistanceInvariance() const { return 1e9f; }
    bool IsInKeyFrame(const std::shared_ptr<KeyFrame>&) const { return false; }
    int PredictScale(float, const std::shared_ptr<KeyFrame>&) const { return 0; }
    int PredictScale(float, const void*) const { return 0; }
    void Replace(const std::shared_ptr<MapPoint>&) {}
    void AddObservation(const std::shared_ptr<KeyFrame>&, int) {}
    std::tuple<int, int> GetIndexInKeyFrame(const std::shared_ptr<KeyFrame>&) const { return std::tuple<int, int>(-1, -1); }
};
typedef std::shared_ptr<MapPoint> MapPointPtr;
static float standin_dist(const V3& p, const V3& o) { const float d[3] = {p.v[0] - o.v[0], p.v[1] - o.v[1], p.v[2] - o.v[2]}; return d[0] * d[0] + d[1] * d[1] + d[2] * d[2]; }
static bool standin_view_gate(const V3&, const V3&, const V3&, float) { return false; }
struct Frame {
    std::map<unsigned, std::vector<unsigned>> mFeatVec;
    int N = 0; std::vector<cv::KeyPoint> mvKeys, mvKeysUn; cv::Mat mDescriptors; std::vector<float> mvuRight, mvScaleFactors, mvLevelSigma2;
    std::vector<MapPointPtr> mvpMapPoints; std::vector<bool> mvbOutlier; float mbf = 40, mb = 0.08f; Camera* mpCamera = nullptr; Pose pose;
    static float mnMinX, mnMinY, mnMaxX, mnMaxY, mfGridElementWidthInv, mfGridElementHeightInv;
    Pose GetPose() const { return pose; }
};
float Frame::mnMinX = 0, Frame::mnMinY = 0, Frame::mnMaxX = 640, Frame::mnMaxY = 480, Frame::mfGridElementWidthInv = 0.1f, Frame::mfGridElementHeightInv = 0.1f;
struct KeyFrame : Frame {
    std::vector<MapPointPtr> GetMapPointMatches() const { return mvpMapPoints; }
    std::set<MapPointPtr> GetMapPointsUnordered() const { return std::set<MapPointPtr>(); }
    MapPointPtr GetMapPoint(int i) const { return mvpMapPoints[i]; }
    float fx = 500, fy = 500, cx = 320, cy = 240;
    float F12[9] = {0}, ep[2] = {0};          // what the reference derives from the two poses (the test passes it in: Pinhole.cpp:127-131 is out of the harness' reach)
    std::vector<float> mvInvLevelSigma2 = std::vector<float>(8, 1.f);
    V3 GetCameraCenter() const { return V3{{0, 0, 0}}; }
    bool IsInImage(float x, float y) const { return x >= 0 && x < 640 && y >= 0 && y < 480; }
    void AddMapPoint(const MapPointPtr& p, int i) { mvpMapPoints[i] = p; }
};
struct Sim3 { Pose pose; Sim3 inverse() const { return *this; } V3 operator*(const V3& p) const { return pose * p; } };
static float standin_norm(const V3& p) { return p.v[0] * p.v[0] + p.v[1] * p.v[1] + p.v[2] * p.v[2]; }
static Pose standin_se3_of_sim3(const Sim3& s) { return s.pose; }
static void standin_fundamental(const KeyFrame& k1, const KeyFrame&, float* F12, float* ep) { std::memcpy(F12, k1.F12, sizeof k1.F12); ep[0] = k1.ep[0]; ep[1] = k1.ep[1]; }

#include "../../shim/plvs_shim.hpp"

namespace {
struct FrameIn {              // flat description of a frame, as tests/test_shim_runtime.py fills it (ctypes mirror there)
    int32_t n; const plvs_keypoint* keys; const uint8_t* desc; const float* uright;
    float min_x, min_y, max_x, max_y, grid_inv_w, grid_inv_h; int32_t nlevels; const float* scale; const float* sigma2; float bf, b;
};
void fill_frame(Frame& F, const FrameIn& in)
{
    F.N = in.n;
    F.mvKeys.resize(in.n); F.mvKeysUn.resize(in.n);
    static_assert(sizeof(cv::KeyPoint) == sizeof(plvs_keypoint), "cv::KeyPoint layout");
    std::memcpy(F.mvKeys.data(), in.keys, sizeof(plvs_keypoint) * (size_t)in.n);
    F.mvKeysUn = F.mvKeys;
    F.mDescriptors.create(std::max(in.n, 1), 32, CV_8U);
    std::memcpy(F.mDescriptors.data, in.desc, (size_t)in.n * 32);
    if (in.uright) F.mvuRight.assign(in.uright, in.uright + in.n);
    F.mvScaleFactors.assign(in.scale, in.scale + in.nlevels); F.mvLevelSigma2.assign(in.sigma2, in.sigma2 + in.nlevels);
    F.mvpMapPoints.assign(in.n, MapPointPtr()); F.mvbOutlier.assign(in.n, false);
    F.mbf = in.bf; F.mb = in.b;
    Frame::mnMinX = in.min_x; Frame::mnMinY = in.min_y; Frame::mnMaxX = in.max_x; Frame::mnMaxY = in.max_y;
    Frame::mfGridElementWidthInv = in.grid_inv_w; Frame::mfGridElementHeightInv = in.grid_inv_h;
}
thread_local std::string g_err;
}  // namespace

extern "C" {

const char* shim_rt_error() { return g_err.c_str(); }

// PLVS2::ORBextractor::operator() (include/ORBextractor.h:86) + the pyramid mirror Frame::ComputeStereoMatches reads
int shim_rt_extract(const uint8_t* img, int w, int h, int nfeatures, int lap0, int lap1, plvs_keypoint* out_kp, uint8_t* out_desc, int cap, int* mono,
                    uint8_t* level1, int level1_cap, int* level1_wh)
{
    try {
        PLVS2::ORBextractor ex(nfeatures, 1.2f, 8, 20, 7);
        cv::Mat image(h, w, CV_8U), desc;
        std::memcpy(image.data, img, (size_t)w * h);
        std::vector<cv::KeyPoint> kps; std::vector<int> lap{lap0, lap1};
        *mono = ex(image, cv::Mat(), kps, desc, lap);
        if ((int)kps.size() > cap) { g_err = "capacity"; return -1; }
        std::memcpy(out_kp, kps.data(), kps.size() * sizeof(cv::KeyPoint));
        for (size_t i = 0; i < kps.size(); ++i) std::memcpy(out_desc + 32 * i, desc.data + i * desc.step, 32);
        ex.SyncImagePyramid(false);
        const cv::Mat& l1 = ex.mvImagePyramid[1];
        level1_wh[0] = l1.cols; level1_wh[1] = l1.rows;
        if (l1.cols * l1.rows <= level1_cap) for (int r = 0; r < l1.rows; ++r) std::memcpy(level1 + (size_t)r * l1.cols, l1.data + r * l1.step, l1.cols);
        return (int)kps.size();
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPointPtr>&, th, bFarPoints, thFarPoints) (include/ORBmatcher.h:68-69), as
// Tracking::SearchLocalPoints calls it.  `q` = one record per local map point in vpMapPoints order; `mp_state[i]`: bit0 = mbTrackInView,
// bit1 = isBad(), bit2 = Observations() > 0.  `pre_claim[k]`: 0 = F.mvpMapPoints[k] empty, 1 = holds a point with observations, 2 = holds one without.
// out_assign[k] = index into vpMapPoints now held by keypoint k, -1 = empty, -2 = still the pre-existing claim.
int shim_rt_search_map(const FrameIn* fin, const plvs_mp_query* q, const uint8_t* mp_state, int nq, const uint8_t* pre_claim, float th, float nnratio,
                       int far_points, float th_far, int32_t* out_assign)
{
    try {
        Frame F; fill_frame(F, *fin);
        std::vector<MapPointPtr> mps(nq);
        std::unordered_map<const MapPoint*, int> index;
        for (int i = 0; i < nq; ++i) {
            auto p = std::make_shared<MapPoint>();
            p->mbTrackInView = mp_state[i] & 1; p->bad = mp_state[i] & 2; p->nobs = (mp_state[i] & 4) ? 3 : 0;
            p->mTrackProjX = q[i].proj_x; p->mTrackProjY = q[i].proj_y; p->mTrackProjXR = q[i].proj_xr; p->mTrackDepth = q[i].track_depth;
            p->mTrackViewCos = q[i].view_cos; p->mnTrackScaleLevel = q[i].level;
            std::memcpy(p->desc.data, q[i].desc, 32);
            mps[i] = p; index[p.get()] = i;
        }
        std::vector<MapPointPtr> pre(F.N);
        for (int k = 0; k < F.N; ++k) if (pre_claim[k]) { pre[k] = std::make_shared<MapPoint>(); pre[k]->nobs = pre_claim[k] == 1 ? 2 : 0; F.mvpMapPoints[k] = pre[k]; }
        PLVS2::ORBmatcher m(nnratio, true);
        const int n = m.SearchByProjection(F, mps, th, far_points != 0, th_far);
        for (int k = 0; k < F.N; ++k) {
            const MapPointPtr& p = F.mvpMapPoints[k];
            out_assign[k] = !p ? -1 : (p == pre[k] ? -2 : index.at(p.get()));
        }
        return n;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono) (include/ORBmatcher.h:72), as TrackWithMotionModel calls it.
// last_mp[i]: 0 = LastFrame.mvpMapPoints[i] empty, 1 = point with observations, 2 = point without, 3 = point but mvbOutlier[i]; last_pc = the point's
// position in the CURRENT camera frame (the stand-in pose is a pure translation `cur_t`, so world = pc - cur_t).  out_assign[k] = index i of the last
// frame's keypoint whose map point keypoint k of the current frame now holds.
int shim_rt_search_last(const FrameIn* cur_in, const FrameIn* last_in, const uint8_t* last_mp, const float* last_pc, const uint8_t* last_mp_desc,
                        const float* cam4, const float* cur_t, const float* last_t, float th, int mono, int check_ori, int32_t* out_assign)
{
    try {
        Frame Cur, Last; fill_frame(Last, *last_in); fill_frame(Cur, *cur_in);
        Camera cam; cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
        Cur.mpCamera = &cam; Last.mpCamera = &cam;
        Cur.pose.t = V3{{cur_t[0], cur_t[1], cur_t[2]}}; Last.pose.t = V3{{last_t[0], last_t[1], last_t[2]}};
        std::unordered_map<const MapPoint*, int> index;
        for (int i = 0; i < Last.N; ++i) {
            if (!last_mp[i]) continue;
            auto p = std::make_shared<MapPoint>();
            p->nobs = last_mp[i] == 2 ? 0 : 1;
            p->pos = V3{{last_pc[3 * i] - cur_t[0], last_pc[3 * i + 1] - cur_t[1], last_pc[3 * i + 2] - cur_t[2]}};
            std::memcpy(p->desc.data, last_mp_desc + 32 * (size_t)i, 32);
            Last.mvpMapPoints[i] = p; Last.mvbOutlier[i] = last_mp[i] == 3; index[p.get()] = i;
        }
        PLVS2::ORBmatcher m(0.9f, check_ori != 0);
        const int n = m.SearchByProjection(Cur, Last, th, mono != 0);
        for (int k = 0; k < Cur.N; ++k) out_assign[k] = Cur.mvpMapPoints[k] ? index.at(Cur.mvpMapPoints[k].get()) : -1;
        return n;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// ORBmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, bOnlyStereo, bCoarse) (include/ORBmatcher.h:97), as LocalMapping::CreateNewMapFeatures calls it
int shim_rt_triangulation(const FrameIn* in1, const FrameIn* in2, const plvs_featvec* fv1, const plvs_featvec* fv2, const uint8_t* has1, const uint8_t* has2,
                          const float* F12, const float* ep, int only_stereo, int coarse, float nnratio, int check_ori, int32_t* out_pairs, int cap)
{
    try {
        auto k1 = std::make_shared<KeyFrame>(), k2 = std::make_shared<KeyFrame>();
        fill_frame(*k1, *in1); fill_frame(*k2, *in2);
        auto unflatten = [](const plvs_featvec& f, std::map<unsigned, std::vector<unsigned>>& out) {
            for (int i = 0; i < f.n_nodes; ++i) for (int j = f.offsets[i]; j < f.offsets[i + 1]; ++j) out[f.node_ids[i]].push_back((unsigned)f.features[j]);
        };
        unflatten(*fv1, k1->mFeatVec); unflatten(*fv2, k2->mFeatVec);
        auto any = std::make_shared<MapPoint>();
        for (int i = 0; i < k1->N; ++i) if (has1[i]) k1->mvpMapPoints[i] = any;
        for (int i = 0; i < k2->N; ++i) if (has2[i]) k2->mvpMapPoints[i] = any;
        std::memcpy(k1->F12, F12, sizeof k1->F12); k1->ep[0] = ep[0]; k1->ep[1] = ep[1];
        PLVS2::ORBmatcher m(nnratio, check_ori != 0);
        std::vector<std::pair<size_t, size_t>> pairs;
        const int n = m.SearchForTriangulation(k1, k2, pairs, only_stereo != 0, coarse != 0);
        if ((int)pairs.size() > cap) { g_err = "capacity"; return -1; }
        for (size_t i = 0; i < pairs.size(); ++i) { out_pairs[2 * i] = (int32_t)pairs[i].first; out_pairs[2 * i + 1] = (int32_t)pairs[i].second; }
        return n == (int)pairs.size() ? n : -2;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

// chisel_server::ChiselServer as PointCloudMapChisel drives it (src/PointCloudMapChisel.cc:100-131, ChiselServer.h:190-286): camera info, then per
// keyframe either {SetDepthPose, SetDepthImageMemorySharing, SetColorImageMemorySharing, IntegrateLastDepthImage} (route 0) or
// {SetDepthImageMemorySharing, SetPointCloud, IntegrateLastPointCloud} (route 1: points taken from the depth image every `step` pixels).  Finally the
// voxel map (blocks in key order) and, after UpdateMesh, the cloud of GetPointCloud.
struct PointXYZRGBA { float x, y, z; unsigned char b, g, r, a; };
struct PointOut { float x, y, z, normal_x, normal_y, normal_z; unsigned char r, g, b, a; };
struct CloudIn { std::vector<PointXYZRGBA> points; };
struct CloudOut { std::vector<PointOut> points; void clear() { points.clear(); } void push_back(const PointOut& q) { points.push_back(q); } };

int shim_rt_chisel(float voxel, float near_p, float far_p, int carving, float carving_dist, int color, int max_blocks, const double* cam4, int w, int h,
                   int nscans, const float* depth, const uint8_t* bgr, const float* poses /*nscans x 12*/, int route, int step,
                   int32_t* keys, float* sdf, float* weight, uint8_t* rgba, int cap_blocks, int* n_blocks,
                   float* cloud_xyz, float* cloud_nrm, uint8_t* cloud_rgba, int cap_cloud, int* n_cloud)
{
    try {
        chisel_server::ChiselServerParams p;
        p.voxelResolution = voxel; p.nearPlaneDist = near_p; p.farPlaneDist = far_p; p.useCarving = carving != 0; p.carvingDist = carving_dist;
        p.useColor = color != 0; p.maxBlocks = max_blocks;
        chisel_server::ChiselServer cs(p);
        cs.SetDepthCameraInfo(cam4[0], cam4[1], cam4[2], cam4[3], w, h);
        cs.SetColorCameraInfo(cam4[0], cam4[1], cam4[2], cam4[3], w, h);
        std::vector<float> d((size_t)w * h); std::vector<uint8_t> c((size_t)w * h * 3);
        for (int s = 0; s < nscans; ++s) {
            std::memcpy(d.data(), depth + (size_t)s * w * h, d.size() * 4);
            std::memcpy(c.data(), bgr + (size_t)s * w * h * 3, c.size());
            Eigen::Affine3f T;
            for (int r = 0; r < 3; ++r) { for (int k = 0; k < 3; ++k) T.R[r][k] = poses[12 * s + 4 * r + k]; T.t[r] = poses[12 * s + 4 * r + 3]; }
            cs.SetDepthImageMemorySharing(d.data(), w, h, w * 4, (uint64_t)s);
            if (route == 0) {
                cs.SetDepthPose(T);
                if (color) { cs.SetColorPose(T); cs.SetColorImageMemorySharing(c.data(), w, h, w * 3, 3, (uint64_t)s); }
                cs.IntegrateLastDepthImage(false);
            } else {
                CloudIn cloud;       // the camera-frame cloud PointCloudMapping builds from the keyframe's depth image (src/PointCloudMapping.cc:580-640)
                const float fx = (float)cam4[0], fy = (float)cam4[1], cx = (float)cam4[2], cy = (float)cam4[3];
                for (int v = 0; v < h; v += step)
                    for (int u = 0; u < w; u += step) {
                        const float z = d[(size_t)v * w + u];
                        if (!(z > 0.f)) continue;
                        PointXYZRGBA q; q.z = z; q.x = ((float)u - cx) * z / fx; q.y = ((float)v - cy) * z / fy;
                        const uint8_t* px = &c[((size_t)v * w + u) * 3]; q.b = px[0]; q.g = px[1]; q.r = px[2]; q.a = 255;
                        cloud.points.push_back(q);
                    }
                cs.SetPointCloud(cloud, T);
                cs.IntegrateLastPointCloud(false);
            }
        }
        int n = 0;
        plvs_shim::check(plvs_tsdf_download_blocks(cs.handle(), keys, sdf, weight, rgba, cap_blocks, &n), "plvs_tsdf_download_blocks");
        *n_blocks = n;
        cs.UpdateMesh();
        CloudOut out; cs.GetPointCloud(out);
        *n_cloud = (int)out.points.size();
        for (int i = 0; i < (int)out.points.size() && i < cap_cloud; ++i) {
            const PointOut& q = out.points[i];
            cloud_xyz[3 * i] = q.x; cloud_xyz[3 * i + 1] = q.y; cloud_xyz[3 * i + 2] = q.z;
            cloud_nrm[3 * i] = q.normal_x; cloud_nrm[3 * i + 1] = q.normal_y; cloud_nrm[3 * i + 2] = q.normal_z;
            cloud_rgba[4 * i] = q.r; cloud_rgba[4 * i + 1] = q.g; cloud_rgba[4 * i + 2] = q.b; cloud_rgba[4 * i + 3] = q.a;
        }
        return 0;
    } catch (const std::exception& e) { g_err = e.what(); return -1; }
}

}  // extern "C"
