// plvs_shim.hpp -- header-only C++ shims that keep the reference's class surfaces and forward to the
// C ABI of libplvs_b200.so (include/plvs_b200.h).  A PLVS checkout swaps its three translation units
// for these: Tracking / LocalMapping / PointCloudMapping call in unchanged.
//
//   PLVS2::ORBextractor              include/ORBextractor.h:59-170     (reference)
//   PLVS2::ORBmatcher (3 overloads)  include/ORBmatcher.h:61-97
//   chisel_server::ChiselServer      Thirdparty/chisel_server/include/chisel_server/ChiselServer.h:81-322 (integrate path)
//
// With OpenCV/Eigen present the real cv::/Eigen:: types are used.  This build image has neither, so the
// shim is also compilable against the minimal stand-ins of shim/standin.hpp (define PLVS_SHIM_STANDIN):
// tests/test_shim_compile.py compiles and links it that way; the member access patterns are the
// reference's (Frame::mvKeysUn, MapPoint::mTrackProjX, ...), expressed as templates so any type with
// those members works.
#pragma once
#include <cmath>
#include <type_traits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <set>
#include <tuple>
#include <vector>
#include "../include/plvs_b200.h"

#ifdef PLVS_SHIM_STANDIN
#include "standin.hpp"
#else
#include <opencv2/core/core.hpp>
#include <Eigen/Geometry>
#endif

namespace plvs_shim {
#ifndef PLVS_SHIM_STANDIN
template <class V> inline float norm3(const V& p) { return p.norm(); }
#else
template <class V> inline float norm3(const V& p) { return standin_norm(p); }
#endif
#ifndef PLVS_SHIM_STANDIN
template <class Sim3T> inline auto se3_of_sim3(const Sim3T& Scw) { return Sophus::SE3f(Scw.rotationMatrix(), Scw.translation() / Scw.scale()); }
#else
template <class Sim3T> inline auto se3_of_sim3(const Sim3T& Scw) { return standin_se3_of_sim3(Scw); }
#endif

// colours of a cloud point, when it has them (PointXYZ has none: SetPointsAndColors is a no-op for it, Conversions.h:63-67)
template <class P> auto point_rgb_impl(const P& pt, float k, float* out, int) -> decltype((void)pt.r, true) { out[0] = pt.r * k; out[1] = pt.g * k; out[2] = pt.b * k; return true; }
template <class P> bool point_rgb_impl(const P&, float, float*, long) { return false; }
template <class P> bool point_rgb(const P& pt, float k, float* out) { return point_rgb_impl(pt, k, out, 0); }
// keyframe id of a cloud point, when the point type has one (PclPointCloudToChisel fills PointCloud::kfids from it, Conversions.h)
template <class P> auto point_kfid_impl(const P& pt, uint32_t* out, int) -> decltype((void)pt.kfid, true) { *out = (uint32_t)pt.kfid; return true; }
template <class P> bool point_kfid_impl(const P&, uint32_t*, long) { return false; }
template <class P> bool point_kfid(const P& pt, uint32_t* out) { return point_kfid_impl(pt, out, 0); }
// normals / alpha of an output point, when the point type has them (the reference selects GetPointCloud overloads on these fields)
template <class P> auto point_set_normal_impl(P& pt, const float* n, int) -> decltype((void)pt.normal_x, void()) { pt.normal_x = n[0]; pt.normal_y = n[1]; pt.normal_z = n[2]; }
template <class P> void point_set_normal_impl(P&, const float*, long) {}
template <class P> void point_set_normal(P& pt, const float* n) { point_set_normal_impl(pt, n, 0); }
template <class P> auto point_set_kfid_impl(P& pt, uint32_t k, int) -> decltype((void)pt.kfid, true) { pt.kfid = k; return true; }
template <class P> bool point_set_kfid_impl(P&, uint32_t, long) { return false; }
template <class P> bool point_set_kfid(P& pt, uint32_t k) { return point_set_kfid_impl(pt, k, 0); }
template <class P> auto point_set_alpha_impl(P& pt, int) -> decltype((void)pt.a, void()) { pt.a = 255; }
template <class P> void point_set_alpha_impl(P&, long) {}
template <class P> void point_set_alpha(P& pt) { point_set_alpha_impl(pt, 0); }
inline void check(int rc, const char* what)
{
    if (rc != PLVS_OK) throw std::runtime_error(std::string(what) + ": " + plvs_last_error());
}
static_assert(sizeof(cv::KeyPoint) == sizeof(plvs_keypoint), "cv::KeyPoint must be the 28-byte POD the ABI mirrors");
}  // namespace plvs_shim

namespace PLVS2 {

// ------------------------------------------------------------------------------------------------------
class ORBextractor {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    ORBextractor(int nfeatures_, float scaleFactor_, int nlevels_, int iniThFAST_, int minThFAST_, int device = 0)
        : nfeatures(nfeatures_), scaleFactor(scaleFactor_), nlevels(nlevels_), iniThFAST(iniThFAST_), minThFAST(minThFAST_)
    {
        plvs_orb_params p{nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST};
        plvs_shim::check(plvs_orb_create(&p, device, &h_), "plvs_orb_create");
        mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels); mvInvLevelSigma2.resize(nlevels);
        mnFeaturesPerLevel.resize(nlevels);
        plvs_shim::check(plvs_orb_tables(h_, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(), mvInvLevelSigma2.data(),
                                         mnFeaturesPerLevel.data()), "plvs_orb_tables");
        mvImagePyramid.resize(nlevels); mvImagePyramidFiltered.resize(nlevels);
    }
    ~ORBextractor() { plvs_orb_destroy(h_); }
    ORBextractor(const ORBextractor&) = delete;
    ORBextractor& operator=(const ORBextractor&) = delete;

    // Compute the ORB features and descriptors on an image (mask is ignored, as in the reference).
    int operator()(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints, cv::OutputArray _descriptors,
                   std::vector<int>& vLappingArea)
    {
        if (_image.empty()) return -1;
        cv::Mat image = _image.getMat();
        const int cap = 2 * nfeatures + 64 * nlevels;
        _keypoints.resize(cap);
        cv::Mat desc(cap, 32, CV_8U);
        int n = 0, mono = 0;
        const int lap0 = vLappingArea.size() > 0 ? vLappingArea[0] : 0, lap1 = vLappingArea.size() > 1 ? vLappingArea[1] : 0;
        plvs_shim::check(plvs_orb_extract(h_, image.data, image.cols, image.rows, (int)image.step, 0, lap0, lap1,
                                          reinterpret_cast<plvs_keypoint*>(_keypoints.data()), desc.data, cap, &n, &mono), "plvs_orb_extract");
        _keypoints.resize(n);
        if (n == 0) _descriptors.release();
        else desc.rowRange(0, n).copyTo(_descriptors);
        pyramidFresh_ = false;
        return mono;
    }

    int inline GetLevels() { return nlevels; }
    float inline GetScaleFactor() { return scaleFactor; }
    std::vector<float> inline GetScaleFactors() { return mvScaleFactor; }
    std::vector<float> inline GetInverseScaleFactors() { return mvInvScaleFactor; }
    std::vector<float> inline GetScaleSigmaSquares() { return mvLevelSigma2; }
    std::vector<float> inline GetInverseScaleSigmaSquares() { return mvInvLevelSigma2; }
    int GetBorderX() const { return 19; }
    int GetBorderY() const { return 19; }

    // mvImagePyramid is a public member that Frame::ComputeStereoMatches reads (src/Frame.cc:1789): the device
    // pyramid is mirrored to the host on demand instead of on every frame.
    void SyncImagePyramid(bool filtered = false)
    {
        for (int l = 0; l < nlevels; ++l) {
            const uint8_t* d; int w, h, pitch;
            plvs_shim::check(plvs_orb_pyramid_level(h_, 0, l, filtered ? 1 : 0, &d, &w, &h, &pitch), "plvs_orb_pyramid_level");
            cv::Mat& m = filtered ? mvImagePyramidFiltered[l] : mvImagePyramid[l];
            m.create(h, w, CV_8U);
            plvs_shim::check(plvs_orb_download_level(h_, 0, l, filtered ? 1 : 0, m.data, (int)m.step), "plvs_orb_download_level");
        }
        pyramidFresh_ = true;
    }

    // void PrecomputeGaussianPyramid(const cv::Mat& image)  (include/ORBextractor.h:116, src/ORBextractor.cc:1401-1427): the line-feature front end
    // calls it to get mvImagePyramid and mvImagePyramidFiltered (each level cloned and blurred 7x7, sigma 2) before operator().  Both pyramids
    // are by-products of the device extraction, so this runs one extraction (keypoints discarded) and mirrors them; the following operator() on the
    // same image repeats the device work (about the cost the reference saves by its flag) and returns identical results.
    void PrecomputeGaussianPyramid(const cv::Mat& image)
    {
        if (image.empty()) return;
        const int cap = 2 * nfeatures + 64 * nlevels;
        std::vector<cv::KeyPoint> kps(cap);
        cv::Mat desc(cap, 32, CV_8U);
        int n = 0, mono = 0;
        plvs_shim::check(plvs_orb_extract(h_, image.data, image.cols, image.rows, (int)image.step, 0, 0, 0, reinterpret_cast<plvs_keypoint*>(kps.data()), desc.data, cap,
                                          &n, &mono), "plvs_orb_extract");
        SyncImagePyramid(false);
        SyncImagePyramid(true);
        mbPrecomputedGaussianPyramid = true;
    }

    std::vector<cv::Mat> mvImagePyramid;
    std::vector<cv::Mat> mvImagePyramidFiltered;
    bool mbPrecomputedGaussianPyramid = false;
    plvs_orb* handle() { return h_; }

protected:
    int nfeatures; float scaleFactor; int nlevels, iniThFAST, minThFAST;
    std::vector<int> mnFeaturesPerLevel;
    std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
    plvs_orb* h_ = nullptr;
    bool pyramidFresh_ = false;
};

// ------------------------------------------------------------------------------------------------------
// ORBmatcher: the three hot-path overloads.  FrameT / KeyFrameT / MapPointPtr are the reference's own types;
// only the members the reference's code reads are touched (gather), and only Frame::mvpMapPoints is written
// (scatter), in the reference's order.
class ORBmatcher {
public:
    ORBmatcher(float nnratio = 0.6f, bool checkOri = true, int device = 0) : mfNNratio(nnratio), mbCheckOrientation(checkOri)
    {
        plvs_shim::check(plvs_match_create(device, &h_), "plvs_match_create");
    }
    ~ORBmatcher() { plvs_match_destroy(h_); }
    ORBmatcher(const ORBmatcher&) = delete;

    static int DescriptorDistance(const cv::Mat& a, const cv::Mat& b) { return plvs_hamming256(a.data, b.data); }

    template <class FrameT>
    static plvs_frame_view view_of(const FrameT& F, const std::vector<cv::KeyPoint>& keysUn, const cv::Mat& descriptors)
    {
        plvs_frame_view v{};
        v.n = F.N;
        v.keys = reinterpret_cast<const plvs_keypoint*>(keysUn.data());
        v.desc = descriptors.data;
        v.uright = F.mvuRight.empty() ? nullptr : F.mvuRight.data();
        v.min_x = FrameT::mnMinX; v.min_y = FrameT::mnMinY; v.max_x = FrameT::mnMaxX; v.max_y = FrameT::mnMaxY;
        v.grid_inv_w = FrameT::mfGridElementWidthInv; v.grid_inv_h = FrameT::mfGridElementHeightInv;
        v.nlevels = (int)F.mvScaleFactors.size();
        for (int i = 0; i < v.nlevels && i < PLVS_MAX_LEVELS; ++i) { v.scale_factors[i] = F.mvScaleFactors[i]; v.level_sigma2[i] = F.mvLevelSigma2[i]; }
        v.bf = F.mbf;
        v.on_device = 0;
        return v;
    }

    // int SearchByProjection(Frame &F, const std::vector<MapPointPtr> &vpMapPoints, const float th=3, const bool bFarPoints=false, const float thFarPoints=50.f)
    template <class FrameT, class MapPointPtr>
    int SearchByProjection(FrameT& F, const std::vector<MapPointPtr>& vpMapPoints, const float th = 3, const bool bFarPoints = false,
                           const float thFarPoints = 50.f)
    {
        std::vector<plvs_mp_query> q;
        std::vector<size_t> src;
        q.reserve(vpMapPoints.size());
        for (size_t i = 0; i < vpMapPoints.size(); ++i) {
            const MapPointPtr& pMP = vpMapPoints[i];
            if (!pMP->mbTrackInView) continue;                 // RGB-D: the right-camera branch does not exist (Nleft == -1)
            if (pMP->isBad()) continue;
            plvs_mp_query e{};
            e.proj_x = pMP->mTrackProjX; e.proj_y = pMP->mTrackProjY; e.proj_xr = pMP->mTrackProjXR;
            e.track_depth = pMP->mTrackDepth; e.view_cos = pMP->mTrackViewCos; e.level = pMP->mnTrackScaleLevel;
            e.flags = pMP->Observations() > 0 ? PLVS_Q_OBS_POSITIVE : 0u;
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(e.desc, d.data, 32);
            q.push_back(e); src.push_back(i);
        }
        std::vector<uint8_t> claimed(F.N, 0);
        for (int i = 0; i < F.N; ++i) claimed[i] = (F.mvpMapPoints[i] && F.mvpMapPoints[i]->Observations() > 0) ? 1 : 0;
        std::vector<int32_t> assign(F.N, -1);
        int nmatches = 0;
        const plvs_frame_view v = view_of(F, F.mvKeysUn, F.mDescriptors);
        plvs_shim::check(plvs_match_projection_map(h_, &v, q.data(), (int)q.size(), th, mfNNratio, bFarPoints ? 1 : 0, thFarPoints, claimed.data(),
                                                   assign.data(), &nmatches), "plvs_match_projection_map");
        for (int i = 0; i < F.N; ++i) if (assign[i] >= 0) F.mvpMapPoints[i] = vpMapPoints[src[assign[i]]];
        return nmatches;
    }

    // int SearchByProjection(Frame &CurrentFrame, const Frame &LastFrame, const float th, const bool bMono)
    template <class FrameT>
    int SearchByProjection(FrameT& CurrentFrame, const FrameT& LastFrame, const float th, const bool bMono)
    {
        const auto Tcw = CurrentFrame.GetPose();
        const auto twc = Tcw.inverse().translation();
        const auto Tlw = LastFrame.GetPose();
        const auto tlc = Tlw * twc;
        const bool bForward = tlc(2) > CurrentFrame.mb && !bMono;
        const bool bBackward = -tlc(2) > CurrentFrame.mb && !bMono;
        std::vector<plvs_last_query> q;
        std::vector<int> src;
        for (int i = 0; i < LastFrame.N; ++i) {
            auto pMP = LastFrame.mvpMapPoints[i];
            if (!pMP || LastFrame.mvbOutlier[i]) continue;
            const auto x3Dw = pMP->GetWorldPos();
            const auto x3Dc = Tcw * x3Dw;                      // the reference's own expressions: projection stays bit-identical
            plvs_last_query e{};
            e.invz = (float)(1.0 / x3Dc(2));
            const auto uv = CurrentFrame.mpCamera->project(x3Dc);
            e.u = uv(0); e.v = uv(1);
            e.last_octave = LastFrame.mvKeys[i].octave;
            e.angle = LastFrame.mvKeysUn[i].angle;
            e.flags = pMP->Observations() > 0 ? PLVS_Q_OBS_POSITIVE : 0u;
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(e.desc, d.data, 32);
            q.push_back(e); src.push_back(i);
        }
        std::vector<uint8_t> claimed(CurrentFrame.N, 0);
        for (int i = 0; i < CurrentFrame.N; ++i)
            claimed[i] = (CurrentFrame.mvpMapPoints[i] && CurrentFrame.mvpMapPoints[i]->Observations() > 0) ? 1 : 0;
        std::vector<int32_t> assign(CurrentFrame.N, -1);
        int nmatches = 0;
        const plvs_frame_view v = view_of(CurrentFrame, CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors);
        plvs_shim::check(plvs_match_projection_last(h_, &v, q.data(), (int)q.size(), th, bForward, bBackward, mbCheckOrientation ? 1 : 0,
                                                    claimed.data(), assign.data(), &nmatches), "plvs_match_projection_last");
        for (int i = 0; i < CurrentFrame.N; ++i) if (assign[i] >= 0) CurrentFrame.mvpMapPoints[i] = LastFrame.mvpMapPoints[src[assign[i]]];
        return nmatches;
    }

    // int SearchForTriangulation(KeyFramePtr& pKF1, KeyFramePtr& pKF2, vector<pair<size_t,size_t>>& vMatchedPairs, bool bOnlyStereo, bool bCoarse=false)
    template <class KeyFramePtr>
    int SearchForTriangulation(KeyFramePtr& pKF1, KeyFramePtr& pKF2, std::vector<std::pair<size_t, size_t>>& vMatchedPairs, const bool bOnlyStereo,
                               const bool bCoarse = false)
    {
        struct Flat { std::vector<uint32_t> ids; std::vector<int32_t> off, feat; plvs_featvec fv; };
        auto flatten = [](const auto& featVec, Flat& f) {       // DBoW2::FeatureVector = std::map<NodeId, std::vector<unsigned>>
            f.off.push_back(0);
            for (const auto& kv : featVec) {
                f.ids.push_back(kv.first);
                for (unsigned i : kv.second) f.feat.push_back((int32_t)i);
                f.off.push_back((int32_t)f.feat.size());
            }
            f.fv = plvs_featvec{(int32_t)f.ids.size(), f.ids.data(), f.off.data(), f.feat.data()};
        };
        Flat f1, f2;
        flatten(pKF1->mFeatVec, f1); flatten(pKF2->mFeatVec, f2);
        std::vector<uint8_t> has1(pKF1->N), has2(pKF2->N);
        for (int i = 0; i < pKF1->N; ++i) has1[i] = pKF1->GetMapPoint(i) ? 1 : 0;
        for (int i = 0; i < pKF2->N; ++i) has2[i] = pKF2->GetMapPoint(i) ? 1 : 0;
        // F12 and the epipole with the reference's own Eigen expressions (src/CameraModels/Pinhole.cpp:127-131, src/ORBmatcher.cc:1006-1011)
        float F12[9], ep[2];
#ifdef PLVS_SHIM_STANDIN
        standin_fundamental(*pKF1, *pKF2, F12, ep);
#else
        {
            const Sophus::SE3f T1w = pKF1->GetPose(), T2w = pKF2->GetPose(), Tw2 = pKF2->GetPoseInverse();
            const Eigen::Vector3f Cw = pKF1->GetCameraCenter();
            const Eigen::Vector3f C2 = T2w * Cw;
            const Eigen::Vector2f e = pKF2->mpCamera->project(C2);
            const Sophus::SE3f T12 = T1w * Tw2;
            const Eigen::Matrix3f R12 = T12.rotationMatrix();
            const Eigen::Vector3f t12 = T12.translation();
            const Eigen::Matrix3f t12x = Sophus::SO3f::hat(t12);
            const Eigen::Matrix3f K1 = pKF1->mpCamera->toK_(), K2 = pKF2->mpCamera->toK_();
            const Eigen::Matrix3f F = K1.transpose().inverse() * t12x * R12 * K2.inverse();
            for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) F12[3 * r + c] = F(r, c);
            ep[0] = e(0); ep[1] = e(1);
        }
#endif
        const plvs_frame_view v1 = view_of(*pKF1, pKF1->mvKeysUn, pKF1->mDescriptors), v2 = view_of(*pKF2, pKF2->mvKeysUn, pKF2->mDescriptors);
        std::vector<int32_t> m12(pKF1->N, -1);
        int nmatches = 0;
        plvs_shim::check(plvs_match_triangulation(h_, &v1, &v2, &f1.fv, &f2.fv, has1.data(), has2.data(), F12, ep, bOnlyStereo ? 1 : 0, bCoarse ? 1 : 0,
                                                  mbCheckOrientation ? 1 : 0, m12.data(), &nmatches), "plvs_match_triangulation");
        vMatchedPairs.clear();
        vMatchedPairs.reserve(nmatches);
        for (size_t i = 0; i < m12.size(); ++i) if (m12[i] >= 0) vMatchedPairs.push_back(std::make_pair(i, (size_t)m12[i]));
        return nmatches;
    }

    // int SearchByProjection(Frame& CurrentFrame, KeyFramePtr& pKF, const std::set<MapPointPtr>& sAlreadyFound, const float th, const int ORBdist)
    // (src/ORBmatcher.cc:1996-2122, Tracking::Relocalization)
    template <class FrameT, class KeyFramePtr, class SetT>
    int SearchByProjection(FrameT& CurrentFrame, KeyFramePtr& pKF, const SetT& sAlreadyFound, const float th, const int ORBdist)
    {
        const auto Tcw = CurrentFrame.GetPose();
        const auto Ow = Tcw.inverse().translation();
        const auto vpMPs = pKF->GetMapPointMatches();
        std::vector<plvs_last_query> q;
        std::vector<size_t> src;
        for (size_t i = 0; i < vpMPs.size(); ++i) {
            auto pMP = vpMPs[i];
            if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) continue;
            const auto x3Dw = pMP->GetWorldPos();
            const auto x3Dc = Tcw * x3Dw;
            const auto uv = CurrentFrame.mpCamera->project(x3Dc);
            if (uv(0) < FrameT::mnMinX || uv(0) > FrameT::mnMaxX) continue;
            if (uv(1) < FrameT::mnMinY || uv(1) > FrameT::mnMaxY) continue;
#ifdef PLVS_SHIM_STANDIN
            const float dist3D = standin_dist(x3Dw, Ow);
#else
            const Eigen::Vector3f PO = x3Dw - Ow;
            const float dist3D = PO.norm();
#endif
            if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
            plvs_last_query e{};
            e.u = uv(0); e.v = uv(1); e.invz = 1.f;
            e.last_octave = pMP->PredictScale(dist3D, &CurrentFrame);
            e.angle = pKF->mvKeysUn[i].angle;
            e.flags = PLVS_Q_OBS_POSITIVE;                      // any map point written here blocks later ones (:2066)
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(e.desc, d.data, 32);
            q.push_back(e); src.push_back(i);
        }
        std::vector<uint8_t> claimed(CurrentFrame.N, 0);
        for (int i = 0; i < CurrentFrame.N; ++i) claimed[i] = CurrentFrame.mvpMapPoints[i] ? 1 : 0;
        std::vector<int32_t> assign(CurrentFrame.N + 1, -1);
        int nmatches = 0;
        const plvs_frame_view v = view_of(CurrentFrame, CurrentFrame.mvKeysUn, CurrentFrame.mDescriptors);
        plvs_shim::check(plvs_match_projection_reloc(h_, &v, q.data(), (int)q.size(), th, ORBdist, mbCheckOrientation ? 1 : 0, claimed.data(), assign.data(), &nmatches),
                         "plvs_match_projection_reloc");
        for (int i = 0; i < CurrentFrame.N; ++i) if (assign[i] >= 0) CurrentFrame.mvpMapPoints[i] = vpMPs[src[assign[i]]];
        return nmatches;
    }

    // int SearchByBoW(KeyFramePtr& pKF, Frame& F, std::vector<MapPointPtr>& vpMapPointMatches)   (src/ORBmatcher.cc:300-506)
    template <class KeyFramePtr, class FrameT, class MapPointPtr>
    int SearchByBoW(KeyFramePtr& pKF, FrameT& F, std::vector<MapPointPtr>& vpMapPointMatches)
    {
        struct Flat { std::vector<uint32_t> ids; std::vector<int32_t> off, feat; plvs_featvec fv; };
        auto flatten = [](const auto& featVec, Flat& f) {
            f.off.push_back(0);
            for (const auto& kv : featVec) {
                f.ids.push_back(kv.first);
                for (unsigned i : kv.second) f.feat.push_back((int32_t)i);
                f.off.push_back((int32_t)f.feat.size());
            }
            f.fv = plvs_featvec{(int32_t)f.ids.size(), f.ids.data(), f.off.data(), f.feat.data()};
        };
        Flat fk, ff;
        flatten(pKF->mFeatVec, fk); flatten(F.mFeatVec, ff);
        const std::vector<MapPointPtr> vpMapPointsKF = pKF->GetMapPointMatches();
        std::vector<uint8_t> has(pKF->N);
        for (int i = 0; i < pKF->N; ++i) has[i] = (vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad()) ? 1 : 0;
        const plvs_frame_view vk = view_of(*pKF, pKF->mvKeysUn, pKF->mDescriptors), vf = view_of(F, F.mvKeysUn, F.mDescriptors);
        std::vector<int32_t> m(F.N + 1, -1);
        int nmatches = 0;
        plvs_shim::check(plvs_match_bow(h_, &vk, &vf, &fk.fv, &ff.fv, has.data(), mfNNratio, mbCheckOrientation ? 1 : 0, m.data(), &nmatches), "plvs_match_bow");
        vpMapPointMatches = std::vector<MapPointPtr>(F.N, static_cast<MapPointPtr>(nullptr));
        for (int i = 0; i < F.N; ++i) if (m[i] >= 0) vpMapPointMatches[i] = vpMapPointsKF[m[i]];
        return nmatches;
    }

    // int Fuse(KeyFramePtr& pKF, const vector<MapPointPtr>& vpMapPoints, const float th=3.0, const bool bRight=false)
    // (src/ORBmatcher.cc:1244-1435).  The gates before the search run here with the reference's own expressions (:1277-1338),
    // the window search runs on the device (plvs_match_fuse, independent per map point), and the Replace / AddObservation /
    // AddMapPoint bookkeeping is applied here in the reference's order (:1409-1427).  RGB-D / rectified stereo: bRight == false.
    template <class KeyFramePtr, class MapPointPtr>
    int Fuse(KeyFramePtr& pKF, const std::vector<MapPointPtr>& vpMapPoints, const float th = 3.0f, const bool bRight = false)
    {
        (void)bRight;
        const auto Tcw = pKF->GetPose();
        const auto Ow = pKF->GetCameraCenter();
        const float bf = pKF->mbf;
        std::vector<plvs_fuse_query> q;
        std::vector<size_t> src;
        for (size_t i = 0; i < vpMapPoints.size(); ++i) {
            MapPointPtr pMP = vpMapPoints[i];
            if (!pMP) continue;
            if (pMP->isBad()) continue;
            else if (pMP->IsInKeyFrame(pKF)) continue;
            const auto p3Dw = pMP->GetWorldPos();
            const auto p3Dc = Tcw * p3Dw;
            if (p3Dc(2) < 0.0f) continue;
            const float invz = 1 / p3Dc(2);
            const auto uv = pKF->mpCamera->project(p3Dc);
            if (!pKF->IsInImage(uv(0), uv(1))) continue;
            const float ur = uv(0) - bf * invz;
            const float maxDistance = pMP->GetMaxDistanceInvariance();
            const float minDistance = pMP->GetMinDistanceInvariance();
#ifdef PLVS_SHIM_STANDIN
            const float dist3D = standin_dist(p3Dw, Ow);
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            if (standin_view_gate(p3Dw, Ow, pMP->GetNormal(), dist3D)) continue;
#else
            const Eigen::Vector3f PO = p3Dw - Ow;
            const float dist3D = PO.norm();
            if (dist3D < minDistance || dist3D > maxDistance) continue;
            const Eigen::Vector3f Pn = pMP->GetNormal();
            if (PO.dot(Pn) < 0.5 * dist3D) continue;
#endif
            plvs_fuse_query e{};
            e.u = uv(0); e.v = uv(1); e.ur = ur;
            e.level = pMP->PredictScale(dist3D, pKF);
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(e.desc, d.data, 32);
            q.push_back(e); src.push_back(i);
        }
        std::vector<int32_t> best_idx(q.size() + 1, -1), best_dist(q.size() + 1, 256);
        int nsearch = 0;
        const plvs_frame_view v = view_of(*pKF, pKF->mvKeysUn, pKF->mDescriptors);
        plvs_shim::check(plvs_match_fuse(h_, &v, pKF->mvInvLevelSigma2.data(), q.data(), (int)q.size(), th, best_idx.data(), best_dist.data(), &nsearch),
                         "plvs_match_fuse");
        int nFused = 0;
        for (size_t k = 0; k < q.size(); ++k) {
            if (best_dist[k] > TH_LOW) continue;
            MapPointPtr pMP = vpMapPoints[src[k]];
            const int bestIdx = best_idx[k];
            MapPointPtr pMPinKF = pKF->GetMapPoint(bestIdx);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) {
                    if (pMPinKF->Observations() > pMP->Observations()) pMP->Replace(pMPinKF);
                    else pMPinKF->Replace(pMP);
                }
            } else {
                pMP->AddObservation(pKF, bestIdx);
                pKF->AddMapPoint(pMP, bestIdx);
            }
            nFused++;
        }
        return nFused;
    }

    // int SearchByProjection(KeyFramePtr& pKF, Sophus::Sim3f& Scw, const vector<MapPointPtr>& vpPoints, vector<MapPointPtr>& vpMatched, int th, float ratioHamming=1.0)
    // (src/ORBmatcher.cc:509-615) and, with vpPointsKFs / vpMatchedKF non-null, its sibling (:617-730)
    template <class KeyFramePtr, class Sim3T, class MapPointPtr>
    int SearchByProjection(KeyFramePtr& pKF, Sim3T& Scw, const std::vector<MapPointPtr>& vpPoints, std::vector<MapPointPtr>& vpMatched, int th,
                           float ratioHamming = 1.0f, const std::vector<KeyFramePtr>* vpPointsKFs = nullptr, std::vector<KeyFramePtr>* vpMatchedKF = nullptr)
    {
        const auto Tcw = plvs_shim::se3_of_sim3(Scw);
        const auto Ow = Tcw.inverse().translation();
        std::set<MapPointPtr> spAlreadyFound(vpMatched.begin(), vpMatched.end());
        spAlreadyFound.erase(static_cast<MapPointPtr>(nullptr));
        std::vector<plvs_last_query> q;
        std::vector<size_t> src;
        for (size_t i = 0; i < vpPoints.size(); ++i) {
            MapPointPtr pMP = vpPoints[i];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            const auto p3Dw = pMP->GetWorldPos();
            const auto p3Dc = Tcw * p3Dw;
            if (p3Dc(2) < 0.0) continue;
            const auto uv = pKF->mpCamera->project(p3Dc);
            if (!pKF->IsInImage(uv(0), uv(1))) continue;
#ifdef PLVS_SHIM_STANDIN
            const float dist = standin_dist(p3Dw, Ow);
            if (dist < pMP->GetMinDistanceInvariance() || dist > pMP->GetMaxDistanceInvariance()) continue;
            if (standin_view_gate(p3Dw, Ow, pMP->GetNormal(), dist)) continue;
#else
            const Eigen::Vector3f PO = p3Dw - Ow;
            const float dist = PO.norm();
            if (dist < pMP->GetMinDistanceInvariance() || dist > pMP->GetMaxDistanceInvariance()) continue;
            const Eigen::Vector3f Pn = pMP->GetNormal();
            if (PO.dot(Pn) < 0.5 * dist) continue;
#endif
            plvs_last_query e{};
            e.u = uv(0); e.v = uv(1); e.invz = 1.f;
            e.last_octave = pMP->PredictScale(dist, pKF);
            e.flags = PLVS_Q_OBS_POSITIVE;
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(e.desc, d.data, 32);
            q.push_back(e); src.push_back(i);
        }
        std::vector<uint8_t> matched(pKF->N, 0);
        for (int i = 0; i < pKF->N; ++i) matched[i] = vpMatched[i] ? 1 : 0;
        std::vector<int32_t> assign(pKF->N + 1, -1);
        int nmatches = 0;
        const plvs_frame_view v = view_of(*pKF, pKF->mvKeysUn, pKF->mDescriptors);
        plvs_shim::check(plvs_match_projection_sim3(h_, &v, q.data(), (int)q.size(), (float)th, ratioHamming, matched.data(), assign.data(), &nmatches),
                         "plvs_match_projection_sim3");
        for (int i = 0; i < pKF->N; ++i) if (assign[i] >= 0) {
            vpMatched[i] = vpPoints[src[assign[i]]];
            if (vpPointsKFs && vpMatchedKF) (*vpMatchedKF)[i] = (*vpPointsKFs)[src[assign[i]]];
        }
        return nmatches;
    }

    // int SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize=10)
    // (src/ORBmatcher.cc:732-852): cv::Point2f is two floats, so the vector's storage is handed over as it is and updated in place
    template <class FrameT, class Point2fT>
    int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<Point2fT>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize = 10)
    {
        static_assert(sizeof(Point2fT) == 2 * sizeof(float), "cv::Point2f layout");
        const plvs_frame_view v1 = view_of(F1, F1.mvKeysUn, F1.mDescriptors), v2 = view_of(F2, F2.mvKeysUn, F2.mDescriptors);
        std::vector<int32_t> m(F1.mvKeysUn.size() + 1, -1);
        int nmatches = 0;
        plvs_shim::check(plvs_match_initialization(h_, &v1, &v2, reinterpret_cast<float*>(vbPrevMatched.data()), windowSize, mfNNratio,
                                                   mbCheckOrientation ? 1 : 0, m.data(), &nmatches), "plvs_match_initialization");
        vnMatches12.assign(m.begin(), m.begin() + F1.mvKeysUn.size());
        return nmatches;
    }

    // int SearchByBoW(KeyFramePtr& pKF1, KeyFramePtr& pKF2, std::vector<MapPointPtr>& vpMatches12)   (src/ORBmatcher.cc:853-997)
    template <class KeyFramePtr, class MapPointPtr>
    int SearchByBoW(KeyFramePtr& pKF1, KeyFramePtr& pKF2, std::vector<MapPointPtr>& vpMatches12)
    {
        struct Flat { std::vector<uint32_t> ids; std::vector<int32_t> off, feat; plvs_featvec fv; };
        auto flatten = [](const auto& featVec, Flat& f) {
            f.off.push_back(0);
            for (const auto& kv : featVec) {
                f.ids.push_back(kv.first);
                for (unsigned i : kv.second) f.feat.push_back((int32_t)i);
                f.off.push_back((int32_t)f.feat.size());
            }
            f.fv = plvs_featvec{(int32_t)f.ids.size(), f.ids.data(), f.off.data(), f.feat.data()};
        };
        Flat f1, f2;
        flatten(pKF1->mFeatVec, f1); flatten(pKF2->mFeatVec, f2);
        const std::vector<MapPointPtr> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
        std::vector<uint8_t> has1(pKF1->N), has2(pKF2->N);
        for (int i = 0; i < pKF1->N; ++i) has1[i] = (mp1[i] && !mp1[i]->isBad()) ? 1 : 0;
        for (int i = 0; i < pKF2->N; ++i) has2[i] = (mp2[i] && !mp2[i]->isBad()) ? 1 : 0;
        const plvs_frame_view v1 = view_of(*pKF1, pKF1->mvKeysUn, pKF1->mDescriptors), v2 = view_of(*pKF2, pKF2->mvKeysUn, pKF2->mDescriptors);
        std::vector<int32_t> m(pKF1->N + 1, -1);
        int nmatches = 0;
        plvs_shim::check(plvs_match_bow_kf(h_, &v1, &v2, &f1.fv, &f2.fv, has1.data(), has2.data(), mfNNratio, mbCheckOrientation ? 1 : 0, m.data(), &nmatches),
                         "plvs_match_bow_kf");
        vpMatches12 = std::vector<MapPointPtr>(mp1.size(), static_cast<MapPointPtr>(nullptr));
        for (int i = 0; i < pKF1->N; ++i) if (m[i] >= 0) vpMatches12[i] = mp2[m[i]];
        return nmatches;
    }

    // int SearchBySim3(KeyFramePtr& pKF1, KeyFramePtr& pKF2, std::vector<MapPointPtr>& vpMatches12, const Sophus::Sim3f& S12, const float th)
    // (src/ORBmatcher.cc:1555-1772): the two projection searches run on the device (plvs_match_fuse_sim3), the gates before them and the
    // mutual-agreement check after them here, with the reference's own expressions
    template <class KeyFramePtr, class Sim3T, class MapPointPtr>
    int SearchBySim3(KeyFramePtr& pKF1, KeyFramePtr& pKF2, std::vector<MapPointPtr>& vpMatches12, const Sim3T& S12, const float th)
    {
        const float fx = pKF1->fx, fy = pKF1->fy, cx = pKF1->cx, cy = pKF1->cy;
        const auto T1w = pKF1->GetPose();
        const auto T2w = pKF2->GetPose();
        const auto S21 = S12.inverse();
        const std::vector<MapPointPtr> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
        const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
        std::vector<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);
        for (int i = 0; i < N1; ++i) {
            MapPointPtr pMP = vpMatches12[i];
            if (pMP) {
                vbAlreadyMatched1[i] = true;
                const int idx2 = std::get<0>(pMP->GetIndexInKeyFrame(pKF2));
                if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
            }
        }
        auto pass = [&](const std::vector<MapPointPtr>& from, const std::vector<bool>& done, auto&& to_other, KeyFramePtr& other, std::vector<int>& vnMatch) {
            std::vector<plvs_fuse_query> q; std::vector<int> src;
            for (int i = 0; i < (int)from.size(); ++i) {
                MapPointPtr pMP = from[i];
                if (!pMP || done[i] || pMP->isBad()) continue;
                const auto p3Dc = to_other(pMP->GetWorldPos());
                if (p3Dc(2) < 0.0) continue;
                const float invz = 1.0 / p3Dc(2);
                const float x = p3Dc(0) * invz, y = p3Dc(1) * invz;
                const float u = fx * x + cx, v = fy * y + cy;
                if (!other->IsInImage(u, v)) continue;
                const float dist3D = plvs_shim::norm3(p3Dc);
                if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
                plvs_fuse_query e{};
                e.u = u; e.v = v; e.level = pMP->PredictScale(dist3D, other);
                const cv::Mat d = pMP->GetDescriptor();
                std::memcpy(e.desc, d.data, 32);
                q.push_back(e); src.push_back(i);
            }
            std::vector<int32_t> bi(q.size() + 1, -1), bd(q.size() + 1, 256);
            int n = 0;
            const plvs_frame_view v = view_of(*other, other->mvKeysUn, other->mDescriptors);
            plvs_shim::check(plvs_match_fuse_sim3(h_, &v, q.data(), (int)q.size(), th, bi.data(), bd.data(), &n), "plvs_match_fuse_sim3");
            for (size_t k = 0; k < q.size(); ++k) if (bd[k] <= TH_HIGH) vnMatch[src[k]] = bi[k];
        };
        std::vector<int> vnMatch1(N1, -1), vnMatch2(N2, -1);
        pass(vpMapPoints1, vbAlreadyMatched1, [&](const auto& p3Dw) { return S21 * (T1w * p3Dw); }, pKF2, vnMatch1);
        pass(vpMapPoints2, vbAlreadyMatched2, [&](const auto& p3Dw) { return S12 * (T2w * p3Dw); }, pKF1, vnMatch2);
        int nFound = 0;
        for (int i1 = 0; i1 < N1; ++i1) {
            const int idx2 = vnMatch1[i1];
            if (idx2 >= 0 && vnMatch2[idx2] == i1) { vpMatches12[i1] = vpMapPoints2[idx2]; nFound++; }
        }
        return nFound;
    }

    // int Fuse(KeyFramePtr& pKF, Sophus::Sim3f& Scw, const vector<MapPointPtr>& vpPoints, float th, vector<MapPointPtr>& vpReplacePoint)
    // (src/ORBmatcher.cc:1437-1553, LoopClosing::SearchAndFuse): gates here, search on the device (no chi-square gate), bookkeeping here
    template <class KeyFramePtr, class Sim3T, class MapPointPtr>
    int Fuse(KeyFramePtr& pKF, Sim3T& Scw, const std::vector<MapPointPtr>& vpPoints, float th, std::vector<MapPointPtr>& vpReplacePoint)
    {
        const auto Tcw = plvs_shim::se3_of_sim3(Scw);             // Sophus::SE3f(Scw.rotationMatrix(), Scw.translation()/Scw.scale()) (:1446)
        const auto Ow = Tcw.inverse().translation();
        const auto spAlreadyFound = pKF->GetMapPointsUnordered();
        std::vector<plvs_fuse_query> q;
        std::vector<size_t> src;
        for (size_t i = 0; i < vpPoints.size(); ++i) {
            MapPointPtr pMP = vpPoints[i];
            if (pMP->isBad() || spAlreadyFound.count(pMP)) continue;
            const auto p3Dw = pMP->GetWorldPos();
            const auto p3Dc = Tcw * p3Dw;
            if (p3Dc(2) < 0.0f) continue;
            const auto uv = pKF->mpCamera->project(p3Dc);
            if (!pKF->IsInImage(uv(0), uv(1))) continue;
#ifdef PLVS_SHIM_STANDIN
            const float dist3D = standin_dist(p3Dw, Ow);
            if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
            if (standin_view_gate(p3Dw, Ow, pMP->GetNormal(), dist3D)) continue;
#else
            const Eigen::Vector3f PO = p3Dw - Ow;
            const float dist3D = PO.norm();
            if (dist3D < pMP->GetMinDistanceInvariance() || dist3D > pMP->GetMaxDistanceInvariance()) continue;
            const Eigen::Vector3f Pn = pMP->GetNormal();
            if (PO.dot(Pn) < 0.5 * dist3D) continue;
#endif
            plvs_fuse_query e{};
            e.u = uv(0); e.v = uv(1); e.ur = 0.f;
            e.level = pMP->PredictScale(dist3D, pKF);
            const cv::Mat d = pMP->GetDescriptor();
            std::memcpy(e.desc, d.data, 32);
            q.push_back(e); src.push_back(i);
        }
        std::vector<int32_t> best_idx(q.size() + 1, -1), best_dist(q.size() + 1, 256);
        int nsearch = 0;
        const plvs_frame_view v = view_of(*pKF, pKF->mvKeysUn, pKF->mDescriptors);
        plvs_shim::check(plvs_match_fuse_sim3(h_, &v, q.data(), (int)q.size(), th, best_idx.data(), best_dist.data(), &nsearch), "plvs_match_fuse_sim3");
        int nFused = 0;
        for (size_t k = 0; k < q.size(); ++k) {
            if (best_dist[k] > TH_LOW) continue;
            MapPointPtr pMP = vpPoints[src[k]];
            MapPointPtr pMPinKF = pKF->GetMapPoint(best_idx[k]);
            if (pMPinKF) {
                if (!pMPinKF->isBad()) vpReplacePoint[src[k]] = pMPinKF;
            } else {
                pMP->AddObservation(pKF, best_idx[k]);
                pKF->AddMapPoint(pMP, best_idx[k]);
            }
            nFused++;
        }
        return nFused;
    }

    static const int TH_LOW = 50, TH_HIGH = 100, HISTO_LENGTH = 12;

protected:
    float mfNNratio;
    bool mbCheckOrientation;
    plvs_match* h_ = nullptr;
};

// ------------------------------------------------------------------------------------------------------
// ORBVocabulary (include/ORBVocabulary.h: DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) for the calls PLVS makes on the path:
// loadFromTextFile (src/System.cc) and transform(features, BowVector&, FeatureVector&, levelsup) from Frame::ComputeBoW / KeyFrame::ComputeBoW
// (src/Frame.cc:1498-1505).  BowVectorT / FeatureVectorT are DBoW2's own map types (std::map<WordId, WordValue>, std::map<NodeId,
// std::vector<unsigned>>): the device returns them flattened in ascending key order, so they are rebuilt with hinted insertions.
class ORBVocabulary {
public:
    explicit ORBVocabulary(int device = 0) : device_(device) {}
    ~ORBVocabulary() { if (h_) plvs_voc_destroy(h_); }
    ORBVocabulary(const ORBVocabulary&) = delete;
    bool loadFromTextFile(const std::string& filename)
    {
        if (h_) { plvs_voc_destroy(h_); h_ = nullptr; }
        return plvs_voc_load_text(filename.c_str(), device_, &h_) == PLVS_OK;
    }
    bool empty() const { return h_ == nullptr; }
    unsigned int size() const { return (unsigned int)plvs_voc_size(h_); }
    template <class BowVectorT, class FeatureVectorT>
    void transform(const std::vector<cv::Mat>& features, BowVectorT& v, FeatureVectorT& fv, int levelsup)
    {
        v.clear(); fv.clear();
        if (!h_) return;                                                        // empty(): "safe for subclasses" (TemplatedVocabulary.h:1147-1150)
        const int n = (int)features.size();
        desc_.resize((size_t)n * 32);
        for (int i = 0; i < n; ++i) std::memcpy(&desc_[(size_t)i * 32], features[i].data, 32);
        bowIds_.resize(n + 1); bowVals_.resize(n + 1); fvNodes_.resize(n + 1); fvOff_.resize(n + 2); fvFeat_.resize(n + 1);
        int nb = 0, nn = 0;
        plvs_shim::check(plvs_voc_transform(h_, desc_.data(), n, 0, levelsup, nullptr, nullptr, nullptr, bowIds_.data(), bowVals_.data(), &nb, fvNodes_.data(),
                                            fvOff_.data(), fvFeat_.data(), &nn, nullptr), "plvs_voc_transform");
        for (int i = 0; i < nb; ++i) v.insert(v.end(), typename BowVectorT::value_type(bowIds_[i], bowVals_[i]));
        for (int i = 0; i < nn; ++i) {
            auto it = fv.insert(fv.end(), typename FeatureVectorT::value_type(fvNodes_[i], typename FeatureVectorT::mapped_type()));
            it->second.assign(fvFeat_.begin() + fvOff_[i], fvFeat_.begin() + fvOff_[i + 1]);
        }
    }
private:
    plvs_voc* h_ = nullptr; int device_ = 0;
    std::vector<uint8_t> desc_; std::vector<uint32_t> bowIds_, fvNodes_; std::vector<double> bowVals_; std::vector<int32_t> fvOff_, fvFeat_;
};

// ------------------------------------------------------------------------------------------------------
// LineMatcher::ComputeDescriptorMatches (include/LineMatcher.h, src/LineMatcher.cc:2567-2615): the k = 2 nearest LBD descriptors per query line
// with the neighbour order of the vendored multi-index hashing (Thirdparty/line_descriptor), then the ratio test.  The rest of LineMatcher (the
// geometric gates of SearchByKnn / SearchForTriangulation) stays in the reference's file and calls this instead of mBdm->knnMatch + the loop.
// DMatchT = cv::DMatch (queryIdx, trainIdx, imgIdx, distance).
class LineDescriptorMatcher {
public:
    explicit LineDescriptorMatcher(float nnratio = 0.78f, int device = 0) : mfNNratio(nnratio) { plvs_shim::check(plvs_match_create(device, &h_), "plvs_match_create"); }
    ~LineDescriptorMatcher() { if (h_) plvs_match_destroy(h_); }
    LineDescriptorMatcher(const LineDescriptorMatcher&) = delete;
    template <class DMatchT>
    int ComputeDescriptorMatches(const cv::Mat& ldesc_q, const cv::Mat& ldesc_t, const cv::Mat& queryMask, std::vector<std::vector<DMatchT> >& lmatches, std::vector<bool>& vValidMatch)
    {
        lmatches.clear(); vValidMatch.clear();
        const int nq = ldesc_q.rows, nt = ldesc_t.rows;
        if (nq == 0 || nt < 2) return 0;            // the reference prints an error for empty matrices; with one train descriptor its result is undefined
        q_.resize((size_t)nq * 32); t_.resize((size_t)nt * 32); m_.assign((size_t)nq, 1);
        for (int i = 0; i < nq; ++i) std::memcpy(&q_[(size_t)i * 32], ldesc_q.data + (size_t)i * ldesc_q.step, 32);
        for (int i = 0; i < nt; ++i) std::memcpy(&t_[(size_t)i * 32], ldesc_t.data + (size_t)i * ldesc_t.step, 32);
        const bool masked = !queryMask.empty();
        if (masked) for (int i = 0; i < nq; ++i) m_[i] = queryMask.data[(size_t)i * queryMask.step];
        qi_.resize(nq); ti_.resize((size_t)2 * nq); d_.resize((size_t)2 * nq); v_.resize(nq);
        int rows = 0, nvalid = 0;
        plvs_shim::check(plvs_line_knn2(h_, q_.data(), nq, t_.data(), nt, masked ? m_.data() : nullptr, mfNNratio, qi_.data(), ti_.data(), d_.data(), v_.data(), &rows, &nvalid),
                         "plvs_line_knn2");
        lmatches.resize(rows); vValidMatch.resize(rows);
        for (int r = 0; r < rows; ++r) {
            lmatches[r].resize(2);
            for (int k = 0; k < 2; ++k) { DMatchT& dm = lmatches[r][k]; dm.queryIdx = qi_[r]; dm.trainIdx = ti_[2 * r + k]; dm.imgIdx = 0; dm.distance = d_[2 * r + k]; }
            vValidMatch[r] = v_[r] != 0;
        }
        return nvalid;
    }
    float mfNNratio;
private:
    plvs_match* h_ = nullptr;
    std::vector<uint8_t> q_, t_, m_, v_; std::vector<int32_t> qi_, ti_; std::vector<float> d_;
};

}  // namespace PLVS2

// ------------------------------------------------------------------------------------------------------
namespace chisel_server {

struct ChiselServerParams {        // Thirdparty/chisel_server/include/chisel_server/ChiselServer.h (same field names)
    int chunkSizeX = 16, chunkSizeY = 16, chunkSizeZ = 16;
    float voxelResolution = 0.015f;
    float truncationDistQuad = 0.0019f, truncationDistLinear = -0.00152f, truncationDistConst = 0.001504f, truncationDistScale = 6.0f;
    int weight = 1;
    bool useCarving = true, useColor = true, saveFile = true;
    float carvingDist = 0.05f, nearPlaneDist = 0.05f, farPlaneDist = 5.0f;
    int fusionMode = 1;
    int maxBlocks = 65536;         // extension: capacity of the device block pool
};

class ChiselServer {
public:
    enum class FusionMode { DepthImage, PointCloud };

    explicit ChiselServer(ChiselServerParams& params, int device = 0) : useColor(params.useColor)
    {
        if (params.chunkSizeX != 16 || params.chunkSizeY != 16 || params.chunkSizeZ != 16) throw std::runtime_error("chunk size must be 16^3");
        plvs_tsdf_params p{};
        p.voxel_resolution = params.voxelResolution;
        p.trunc_quad = params.truncationDistQuad; p.trunc_linear = params.truncationDistLinear; p.trunc_const = params.truncationDistConst;
        p.trunc_scale = params.truncationDistScale;
        p.weight = (float)static_cast<uint16_t>(params.weight);
        p.use_carving = params.useCarving; p.carving_dist = params.carvingDist; p.use_color = params.useColor;
        p.near_plane = params.nearPlaneDist; p.far_plane = params.farPlaneDist; p.max_blocks = params.maxBlocks;
        plvs_shim::check(plvs_tsdf_create(&p, device, &h_), "plvs_tsdf_create");
    }
    ~ChiselServer() { plvs_tsdf_destroy(h_); }
    ChiselServer(const ChiselServer&) = delete;

    // ChiselServer::UpdateMesh (ChiselServer.cpp:707-710): marching cubes, colours and gradient normals of every chunk, on the device
    void UpdateMesh() { plvs_shim::check(plvs_tsdf_update_meshes(h_, &nMeshes_, &nMeshVerts_), "plvs_tsdf_update_meshes"); }
    // ChiselServer::GetPointCloud (ChiselServer.cpp:872-1075): one point per mesh vertex; colour * 255 when the map has colour,
    // otherwise the reference's two-light Lambert shading of the normal (lightDir normalised, lightDir1 left as written, :907-911).
    // Chunks come in (x,y,z) key order (the reference walks an unordered_map).  CloudT: pcl::PointCloud<PointT>-like (clear, push_back).
    template <class CloudT>
    void GetPointCloud(CloudT& output_cloud)
    {
        output_cloud.clear();
        if (nMeshVerts_ <= 0) return;
        const size_t n = (size_t)nMeshVerts_;
        meshV_.resize(3 * n); meshN_.resize(3 * n); meshC_.resize(3 * n);
        plvs_shim::check(plvs_tsdf_get_meshes(h_, nullptr, nullptr, 0, meshV_.data(), meshN_.data(), meshC_.data(), nMeshVerts_, 0), "plvs_tsdf_get_meshes");
        const float l0[3] = {0.8f, -0.2f, 0.7f};
        const float l0n = std::sqrt(l0[0] * l0[0] + (l0[1] * l0[1] + l0[2] * l0[2]));
        const float lightDir[3] = {l0[0] / l0n, l0[1] / l0n, l0[2] / l0n}, lightDir1[3] = {-0.5f, 0.2f, 0.2f};
        typedef typename std::decay<decltype(output_cloud.points[0])>::type PointT;
        { PointT probe; if (plvs_shim::point_set_kfid(probe, 0u)) {            // the overload for point types with a kfid field (ChiselServer.cpp:1077-1180)
            meshK_.resize(n);
            plvs_shim::check(plvs_tsdf_get_mesh_kfids(h_, meshK_.data(), nMeshVerts_, 0), "plvs_tsdf_get_mesh_kfids"); } else meshK_.clear(); }
        for (size_t i = 0; i < n; ++i) {
            PointT point;
            if (!meshK_.empty()) plvs_shim::point_set_kfid(point, meshK_[i]);
            point.x = meshV_[3 * i]; point.y = meshV_[3 * i + 1]; point.z = meshV_[3 * i + 2];
            const float* nrm = &meshN_[3 * i];
            if (useColor) {
                point.r = meshC_[3 * i] * 255; point.g = meshC_[3 * i + 1] * 255; point.b = meshC_[3 * i + 2] * 255;
                plvs_shim::point_set_normal(point, nrm);
            } else {
                const float d0 = std::fmax(nrm[0] * lightDir[0] + (nrm[1] * lightDir[1] + nrm[2] * lightDir[2]), 0.0f);
                const float d1 = std::fmax(nrm[0] * lightDir1[0] + (nrm[1] * lightDir1[1] + nrm[2] * lightDir1[2]), 0.0f);
                const float lambert = (d0 * 0.5f + d1 * 0.5f) + 0.2f;
                point.r = std::fmin(lambert, 1.0) * 255; point.g = point.r; point.b = point.r;
                plvs_shim::point_set_alpha(point);
            }
            output_cloud.push_back(point);
        }
    }
    // ChiselServer::SaveMesh (ChiselServer.cpp) -> Chisel::SaveAllMeshesToPLY: the meshes of the last UpdateMesh, chunks in key order
    bool SaveMesh(const std::string& filename)
    {
        const size_t n = (size_t)std::max<long long>(nMeshVerts_, 0);
        meshV_.resize(3 * n); meshC_.resize(3 * n);
        if (n) plvs_shim::check(plvs_tsdf_get_meshes(h_, nullptr, nullptr, 0, meshV_.data(), nullptr, meshC_.data(), nMeshVerts_, 0), "plvs_tsdf_get_meshes");
        return plvs_mesh_save_ply(filename.c_str(), meshV_.data(), useColor ? meshC_.data() : nullptr, nMeshVerts_) == PLVS_OK;
    }
    void Reset() { plvs_shim::check(plvs_tsdf_reset(h_), "plvs_tsdf_reset"); }
    void SetDepthCameraInfo(const double fx, const double fy, const double cx, const double cy, const int width, const int height)
    {
        plvs_shim::check(plvs_tsdf_set_camera(h_, fx, fy, cx, cy, width, height), "plvs_tsdf_set_camera");
        gotInfo = true;
    }
    void SetColorCameraInfo(const double fx, const double fy, const double cx, const double cy, const int w, const int h) { SetDepthCameraInfo(fx, fy, cx, cy, w, h); }
    void SetDepthPose(const Eigen::Affine3f& tf)
    {
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) Twc_[4 * r + c] = tf.linear()(r, c); Twc_[4 * r + 3] = tf.translation()(r); }
        gotPose = true;
    }
    void SetColorPose(const Eigen::Affine3f& tf) { SetDepthPose(tf); }
    // borrowed buffers: must stay valid until IntegrateLastDepthImage returns (as in the reference)
    void SetDepthImageMemorySharing(float* img, int width, int height, int /*step*/, uint64_t /*timestamp*/) { depth_ = img; dw_ = width; dh_ = height; depth16_ = nullptr; }
    void SetColorImageMemorySharing(unsigned char* img, int width, int height, int step, int num_channels, uint64_t /*timestamp*/)
    {
        color_ = img; cw_ = width; ch_ = height; cstep_ = step; cn_ = num_channels;
    }
    // Extension: the raw 16-bit depth map of the sensor; `convertTo(CV_32F, mDepthMapFactor)` (src/Tracking.cc:1812-1813)
    // then runs on the device.  `depthMapFactor` is Tracking's member after its own `1.0f / factor`.
    void SetRawDepthImageMemorySharing(const uint16_t* img, int width, int height, int step_bytes, float depthMapFactor, uint64_t /*timestamp*/)
    {
        depth16_ = img; dw_ = width; dh_ = height; dstep16_ = step_bytes; dfactor_ = depthMapFactor; depth_ = nullptr;
    }
    // ChiselServer::IntegrateLastDepthImage(bool updateMesh = true) (ChiselServer.cpp:623-662): integrates, then chiselMap->UpdateMeshes() when asked
    void IntegrateLastDepthImage(bool updateMesh = true)
    {
        if (gotInfo && gotPose && depth16_ && !depth_) {
            const bool color16 = useColor && color_;
            plvs_shim::check(plvs_tsdf_integrate_depth_u16(h_, depth16_, dw_, dh_, dstep16_, dfactor_, color16 ? color_ : nullptr, cstep_, cn_, Twc_,
                                                           color16 ? PLVS_TSDF_SCAN_COLOR : PLVS_TSDF_SCAN), "plvs_tsdf_integrate_depth_u16");
            if (updateMesh) UpdateMesh();
            return;
        }
        if (!gotInfo || !gotPose || !depth_) { std::fprintf(stderr, "ChiselServer - PROBLEM in integrating depth scan!!! ************\n"); return; }
        const bool color = useColor && color_;
        plvs_shim::check(plvs_tsdf_integrate_depth(h_, depth_, dw_, dh_, color ? color_ : nullptr, cstep_, cn_, Twc_,
                                                   color ? PLVS_TSDF_SCAN_COLOR : PLVS_TSDF_SCAN, 0), "plvs_tsdf_integrate_depth");
        if (updateMesh) UpdateMesh();
    }
    // ChiselServer::SetPointCloud (ChiselServer.cpp:560-583): camera-frame cloud + its pose; colours become floats with
    // byteToFloat = 1.0f / 255.0f exactly as PclPointCloudToChisel does (Conversions.h:69-94).  CloudT is any
    // pcl::PointCloud<PointT>-like type whose points have x, y, z (and r, g, b when useColor).
    template <class CloudT>
    void SetPointCloud(const CloudT& cloud_camera, const Eigen::Affine3f& Twc)
    {
        const size_t n = cloud_camera.points.size();
        cloudXyz_.resize(3 * n);
        cloudRgb_.resize(useColor && useColorCloud_ ? 3 * n : 0);
        cloudKfid_.assign(n, 0u); bool hasKfid = n > 0;
        const float byteToFloat = 1.0f / 255.0f;
        size_t i = 0;
        for (const auto& pt : cloud_camera.points) {
            cloudXyz_[3 * i] = pt.x; cloudXyz_[3 * i + 1] = pt.y; cloudXyz_[3 * i + 2] = pt.z;
            if (!cloudRgb_.empty() && !plvs_shim::point_rgb(pt, byteToFloat, &cloudRgb_[3 * i])) { cloudRgb_.clear(); useColorCloud_ = false; }
            if (hasKfid && !plvs_shim::point_kfid(pt, &cloudKfid_[i])) hasKfid = false;
            ++i;
        }
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) cloudTwc_[4 * r + c] = Twc.linear()(r, c); cloudTwc_[4 * r + 3] = Twc.translation()(r); }
        if (!hasKfid) cloudKfid_.clear();
        gotCloudPose = true;
    }
    // ChiselServer::IntegrateLastPointCloud (ChiselServer.cpp:664-705) -> Chisel::IntegratePointCloudWidthDepth: the carve pass
    // uses the depth image registered with SetDepthImage[MemorySharing] when the depth camera info is known
    void IntegrateLastPointCloud(bool updateMesh = true)
    {
        if (!gotCloudPose) { std::fprintf(stderr, "ChiselServer - PROBLEM in integrating point cloud!!! ************\n"); return; }
        const bool with_depth = gotInfo && depth_;
        if (!cloudKfid_.empty())         // the cloud's points carry keyframe ids: they are stamped on the voxels like PointCloud::kfids (src/Chisel.cpp:470,534)
            plvs_shim::check(plvs_tsdf_integrate_cloud_kf(h_, cloudXyz_.data(), cloudRgb_.empty() ? nullptr : cloudRgb_.data(), cloudKfid_.data(), 0u,
                                                          (int)(cloudXyz_.size() / 3), with_depth ? depth_ : nullptr, with_depth ? dw_ : 0, with_depth ? dh_ : 0, cloudTwc_),
                             "plvs_tsdf_integrate_cloud_kf");
        else
        plvs_shim::check(plvs_tsdf_integrate_cloud(h_, cloudXyz_.data(), cloudRgb_.empty() ? nullptr : cloudRgb_.data(), (int)(cloudXyz_.size() / 3),
                                                   with_depth ? depth_ : nullptr, with_depth ? dw_ : 0, with_depth ? dh_ : 0, cloudTwc_), "plvs_tsdf_integrate_cloud");
        if (updateMesh) UpdateMesh();
    }
    // void ChiselServer::Deform(chisel::MapKfidRt& deformationMap) (ChiselServer.h:202, src/PointCloudMapChisel.cc:489): MapT is any map
    // kfid -> {R (3x3, operator()(r,c)), t (operator()(r))} such as chisel::MapKfidRt.  Chunks are visited in (x,y,z) key order (the reference
    // walks a std::unordered_map: see plvs_tsdf_deform).
    template <class MapT>
    void Deform(MapT& deformationMap)
    {
        std::vector<uint32_t> ids; std::vector<float> Rt;
        ids.reserve(deformationMap.size()); Rt.reserve(12 * deformationMap.size());
        for (const auto& kv : deformationMap) {
            ids.push_back((uint32_t)kv.first);
            for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) Rt.push_back(kv.second.R(r, c)); Rt.push_back(kv.second.t(r)); }
        }
        plvs_shim::check(plvs_tsdf_deform(h_, ids.data(), Rt.data(), (int)ids.size(), nullptr, 0), "plvs_tsdf_deform");
    }
    // template<PointType> void ChiselServer::IntegrateWorldPointCloud(const pcl::PointCloud<PointType>& cloud, chisel::Transform& Twc)
    // (ChiselServer.h:199, PointCloudMapChisel::LoadMap): points with x, y, z, normal_x/y/z, r/g/b (and kfid when the type has it)
    template <class CloudT>
    void IntegrateWorldPointCloud(const CloudT& cloud, const Eigen::Affine3f& Twc)
    {
        const size_t n = cloud.points.size();
        std::vector<float> xyz(3 * n), nrm(3 * n), rgb(useColor ? 3 * n : 0);
        std::vector<uint32_t> kf(n, 0u); bool hasKfid = n > 0;
        const float byteToFloat = 1.0f / 255.0f;
        size_t i = 0;
        for (const auto& pt : cloud.points) {
            xyz[3 * i] = pt.x; xyz[3 * i + 1] = pt.y; xyz[3 * i + 2] = pt.z;
            nrm[3 * i] = pt.normal_x; nrm[3 * i + 1] = pt.normal_y; nrm[3 * i + 2] = pt.normal_z;
            if (!rgb.empty() && !plvs_shim::point_rgb(pt, byteToFloat, &rgb[3 * i])) rgb.clear();
            if (hasKfid && !plvs_shim::point_kfid(pt, &kf[i])) hasKfid = false;
            ++i;
        }
        float T[12];
        for (int r = 0; r < 3; ++r) { for (int c = 0; c < 3; ++c) T[4 * r + c] = Twc.linear()(r, c); T[4 * r + 3] = Twc.translation()(r); }
        plvs_shim::check(plvs_tsdf_integrate_world_cloud(h_, xyz.data(), rgb.empty() ? nullptr : rgb.data(), nrm.data(), hasKfid ? kf.data() : nullptr, 0u, (int)n, T),
                         "plvs_tsdf_integrate_world_cloud");
    }
    // PointCloudMapChisel<PointT>::LoadMap (src/PointCloudMapChisel.cc:527-549) without PCL: PointCloudMap::LoadMap's file reading (src/PointCloudMap.cc:466-503,
    // plvs_map_load_ply), InvertColors (:505-514: swap red and blue), then IntegrateWorldPointCloud with the identity.  A file without normals makes the
    // reference estimate them with pcl::NormalEstimation, which is not replaced: such a file is refused (false), like a missing one.
    bool LoadMapPLY(const std::string& filename)
    {
        long long n = 0; int fields = 0;
        int rc = plvs_map_load_ply(filename.c_str(), nullptr, nullptr, nullptr, nullptr, nullptr, 0, &n, &fields);
        if ((rc != PLVS_OK && rc != PLVS_ECAP) || !(fields & 1) || !(fields & 4)) return false;
        std::vector<float> xyz(3 * (size_t)n + 3), nrm(3 * (size_t)n + 3), rgbf; std::vector<uint8_t> rgb(3 * (size_t)n + 3); std::vector<uint32_t> lab((size_t)n + 1), kf((size_t)n + 1);
        if (plvs_map_load_ply(filename.c_str(), xyz.data(), rgb.data(), nrm.data(), lab.data(), kf.data(), n, &n, &fields) != PLVS_OK) return false;
        if (useColor && (fields & 2)) {
            rgbf.resize(3 * (size_t)n);
            const float byteToFloat = 1.0f / 255.0f;
            for (long long i = 0; i < n; ++i) {       // after InvertColors: r = the file's blue, b = the file's red
                rgbf[3 * i] = rgb[3 * i + 2] * byteToFloat; rgbf[3 * i + 1] = rgb[3 * i + 1] * byteToFloat; rgbf[3 * i + 2] = rgb[3 * i] * byteToFloat;
            }
        }
        const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        plvs_shim::check(plvs_tsdf_integrate_world_cloud(h_, xyz.data(), rgbf.empty() ? nullptr : rgbf.data(), nrm.data(), (fields & 16) ? kf.data() : nullptr, 0u, (int)n, I),
                         "plvs_tsdf_integrate_world_cloud");
        return true;
    }
    plvs_tsdf* handle() { return h_; }

protected:
    bool useColor, gotInfo = false, gotPose = false, gotCloudPose = false;
    std::vector<float> cloudXyz_, cloudRgb_;
    bool useColorCloud_ = true;       // false once a colourless point type was seen
    float cloudTwc_[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    float Twc_[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    float* depth_ = nullptr; int dw_ = 0, dh_ = 0;
    const uint16_t* depth16_ = nullptr; int dstep16_ = 0; float dfactor_ = 1.0f;
    int nMeshes_ = 0; long long nMeshVerts_ = 0; std::vector<float> meshV_, meshN_, meshC_;
    std::vector<uint32_t> cloudKfid_, meshK_;
    unsigned char* color_ = nullptr; int cw_ = 0, ch_ = 0, cstep_ = 0, cn_ = 0;
    plvs_tsdf* h_ = nullptr;
};

}  // namespace chisel_server
