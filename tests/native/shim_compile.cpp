// Compile-and-link check of shim/plvs_shim.hpp against stand-in types (no OpenCV/Eigen in this image).
// The structs below expose exactly the members the reference's Frame / MapPoint / KeyFrame offer to the matchers.
#define PLVS_SHIM_STANDIN
#include <map>
#include <memory>
#include <set>
#include <tuple>
#include <vector>
#include "../../shim/standin.hpp"

struct V3 { float v[3]; float operator()(int i) const { return v[i]; } };
struct V2 { float v[2]; float operator()(int i) const { return v[i]; } };
struct Pose {
    V3 t{};
    Pose inverse() const { return *this; }
    V3 translation() const { return t; }
    V3 operator*(const V3& p) const { return V3{{p.v[0] + t.v[0], p.v[1] + t.v[1], p.v[2] + t.v[2]}}; }
};
struct Camera { V2 project(const V3& p) const { return V2{{500.f * p.v[0] / p.v[2] + 320.f, 500.f * p.v[1] / p.v[2] + 240.f}}; } };
struct KeyFrame;
struct MapPoint {
    bool mbTrackInView = true; float mTrackProjX = 0, mTrackProjY = 0, mTrackProjXR = 0, mTrackDepth = 1, mTrackViewCos = 1; int mnTrackScaleLevel = 0;
    cv::Mat desc{1, 32, CV_8U}; V3 pos{{0, 0, 2}};
    bool isBad() const { return false; }
    int Observations() const { return 1; }
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
    std::vector<float> mvInvLevelSigma2 = std::vector<float>(8, 1.f);
    V3 GetCameraCenter() const { return V3{{0, 0, 0}}; }
    bool IsInImage(float x, float y) const { return x >= 0 && x < 640 && y >= 0 && y < 480; }
    void AddMapPoint(const MapPointPtr& p, int i) { mvpMapPoints[i] = p; }
};
struct Sim3 { Pose pose; Sim3 inverse() const { return *this; } V3 operator*(const V3& p) const { return pose * p; } };
static float standin_norm(const V3& p) { return p.v[0] * p.v[0] + p.v[1] * p.v[1] + p.v[2] * p.v[2]; }
static Pose standin_se3_of_sim3(const Sim3& s) { return s.pose; }
static void standin_fundamental(const KeyFrame&, const KeyFrame&, float* F12, float* ep) { for (int i = 0; i < 9; ++i) F12[i] = 0; ep[0] = ep[1] = 0; }

#include "../../shim/plvs_shim.hpp"

extern "C" int shim_instantiate(int run)
{
    // instantiating every template is the point; nothing executes without a GPU (run == 0)
    if (!run) return (int)sizeof(PLVS2::ORBextractor) + (int)sizeof(PLVS2::ORBmatcher) + (int)sizeof(chisel_server::ChiselServer);
    PLVS2::ORBextractor ex(1000, 1.2f, 8, 20, 7);
    cv::Mat img(480, 640, CV_8U), desc; std::vector<cv::KeyPoint> kps; std::vector<int> lap{0, 0};
    int mono = ex(img, cv::Mat(), kps, desc, lap);
    ex.SyncImagePyramid();
    ex.PrecomputeGaussianPyramid(img);
    PLVS2::ORBmatcher m(0.8f, true);
    Frame F, L; std::vector<MapPointPtr> mps;
    int a = m.SearchByProjection(F, mps, 3.f, false, 50.f);
    int b = m.SearchByProjection(F, L, 15.f, false);
    auto k1 = std::make_shared<KeyFrame>(), k2 = std::make_shared<KeyFrame>();
    std::vector<std::pair<size_t, size_t>> pairs;
    int c = m.SearchForTriangulation(k1, k2, pairs, false, false);
    c += m.Fuse(k1, mps, 3.0f, false);
    c += m.SearchByBoW(k1, F, mps);
    Sim3 scw; std::vector<MapPointPtr> repl(mps.size());
    c += m.Fuse(k1, scw, mps, 4.f, repl);
    std::vector<MapPointPtr> vm(k1->N);
    c += m.SearchBySim3(k1, k2, vm, scw, 7.5f);
    c += m.SearchByBoW(k1, k2, vm);
    c += m.SearchByProjection(k1, scw, mps, vm, 8, 1.5f);
    std::set<MapPointPtr> found;
    c += m.SearchByProjection(F, k1, found, 10.f, 100);
    struct P2f { float x, y; };
    std::vector<P2f> prevMatched(F.mvKeysUn.size()); std::vector<int> m12;
    c += m.SearchForInitialization(F, L, prevMatched, m12, 100);
    PLVS2::ORBVocabulary voc; voc.loadFromTextFile("ORBvoc.txt");
    std::map<unsigned, double> bowVec; std::map<unsigned, std::vector<unsigned>> featVec; std::vector<cv::Mat> vDesc(1, cv::Mat(1, 32, CV_8U));
    voc.transform(vDesc, bowVec, featVec, 4); c += (int)voc.size();
    struct DMatchS { int queryIdx, trainIdx, imgIdx; float distance; };
    PLVS2::LineDescriptorMatcher lmx(0.78f);
    cv::Mat lq(5, 32, CV_8U), lt(7, 32, CV_8U), lmask(5, 1, CV_8U);
    std::vector<std::vector<DMatchS> > lmatches; std::vector<bool> lvalid;
    c += lmx.ComputeDescriptorMatches(lq, lt, lmask, lmatches, lvalid);
    chisel_server::ChiselServerParams p; chisel_server::ChiselServer cs(p);
    cs.SetDepthCameraInfo(500, 500, 320, 240, 640, 480);
    Eigen::Affine3f T; cs.SetDepthPose(T);
    std::vector<float> d(640 * 480, 1.f); cs.SetDepthImageMemorySharing(d.data(), 640, 480, 640 * 4, 0);
    cs.IntegrateLastDepthImage(false);
    std::vector<uint16_t> d16(640 * 480, 5000); cs.SetRawDepthImageMemorySharing(d16.data(), 640, 480, 640 * 2, 1.0f / 5000.0f, 0);
    cs.IntegrateLastDepthImage(false);
    struct PointOut { float x, y, z, normal_x, normal_y, normal_z; unsigned char r, g, b, a; };
    struct CloudOut { std::vector<PointOut> points; void clear() { points.clear(); } void push_back(const PointOut& q) { points.push_back(q); } } out;
    cs.UpdateMesh(); cs.GetPointCloud(out);
    struct PointOutK { float x, y, z, normal_x, normal_y, normal_z; unsigned char r, g, b, a; unsigned kfid; };
    struct CloudOutK { std::vector<PointOutK> points; void clear() { points.clear(); } void push_back(const PointOutK& q) { points.push_back(q); } } outk;
    cs.GetPointCloud(outk);
    // a26: SetPointCloud + IntegrateLastPointCloud with a coloured and a colourless PCL-like cloud
    struct PointXYZRGBA { float x, y, z; unsigned char b, g, r, a; };
    struct PointXYZ { float x, y, z; };
    struct PointKf { float x, y, z; unsigned char b, g, r, a; unsigned kfid; };
    struct CloudK { std::vector<PointKf> points; } ck; ck.points.push_back({0.1f, 0.2f, 1.f, 1, 2, 3, 255, 7u});
    cs.SetPointCloud(ck, T); cs.IntegrateLastPointCloud(false);
    struct CloudC { std::vector<PointXYZRGBA> points; } cc; cc.points.push_back({0.1f, 0.2f, 1.f, 1, 2, 3, 255});
    struct CloudP { std::vector<PointXYZ> points; } cp; cp.points.push_back({0.1f, 0.2f, 1.f});
    cs.SetPointCloud(cc, T); cs.IntegrateLastPointCloud(false);
    cs.SetPointCloud(cp, T); cs.IntegrateLastPointCloud(false);
    // Deform + the map-loading route
    struct M3 { float m[3][3]; float operator()(int r, int c2) const { return m[r][c2]; } };
    struct Vt { float v[3]; float operator()(int r) const { return v[r]; } };
    struct Rt { M3 R; Vt t; };
    std::map<unsigned, Rt> deformation; deformation[7u] = Rt{{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}, {{0.f, 0.f, 0.f}}};
    cs.Deform(deformation);
    struct PointN { float x, y, z, normal_x, normal_y, normal_z; unsigned char b, g, r, a; unsigned kfid; };
    struct CloudN { std::vector<PointN> points; } cn; cn.points.push_back({0.1f, 0.2f, 1.f, 0.f, 0.f, -1.f, 1, 2, 3, 255, 7u});
    cs.IntegrateWorldPointCloud(cn, T);
    c += cs.LoadMapPLY("volumetric_map_out_0.ply") ? 1 : 0;
    return mono + a + b + c;
}
