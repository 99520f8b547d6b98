// TEST INFRASTRUCTURE ONLY -- C driver around the REFERENCE's own src/ORBmatcher.cc.
//
// oracle/ref_build.py compiles /root/reference/src/ORBmatcher.cc where it lies (nothing is copied), DBoW2's FeatureVector.cpp
// and this file into oracle/_ref/libmatch_ref.so; the data-model classes come from oracle/plvs_standin/plvs_types.hpp (see
// its header for what is restated there: the frame grid and the epipolar line distance; everything in ORBmatcher.cc is the
// reference's own).  This file builds Frame / KeyFrame / MapPoint objects from the flat views of the drop-in C ABI
// (include/plvs_b200.h), calls the three searches, and flattens the side effects (Frame::mvpMapPoints, vMatchedPairs)
// back into the arrays the oracle and the CUDA path return.  Entry points mirror oracle/match_oracle.cpp with the prefix ref_.
#include "ORBmatcher.h"

#include <cstdint>
#include <cstring>
#include <memory>
#include <unordered_map>
#include <vector>

using namespace PLVS2;

namespace {

struct KeyPointC { float x, y, size, angle, response; int32_t octave, class_id; };
struct FrameView {
    int32_t n; const KeyPointC* keys; const uint8_t* desc; const float* uright;
    float min_x, min_y, max_x, max_y, grid_inv_w, grid_inv_h;
    float scale_factors[16], level_sigma2[16];
    int32_t nlevels; float bf; int32_t on_device; uint64_t cache_key;
};
struct MpQuery { float proj_x, proj_y, proj_xr, track_depth, view_cos; int32_t level; uint32_t flags; uint8_t desc[32]; };
struct LastQuery { float u, v, invz; int32_t last_octave; float angle; uint32_t flags; uint8_t desc[32]; };
struct FuseQuery { float u, v, ur; int32_t level; uint8_t desc[32]; };
struct FeatVec { int32_t n_nodes; const uint32_t* node_ids; const int32_t* offsets; const int32_t* features; };

void fill_base(FrameBase& f, const FrameView* v)
{
    f.N = v->n; f.Nleft = -1; f.NLeft = -1;
    f.mvKeysUn.resize(v->n);
    for (int i = 0; i < v->n; ++i) {
        const KeyPointC& k = v->keys[i];
        f.mvKeysUn[i] = cv::KeyPoint(k.x, k.y, k.size, k.angle, k.response, k.octave, k.class_id);
    }
    f.mvKeys = f.mvKeysUn;
    f.mvuRight.assign(v->n, -1.f);
    if (v->uright) for (int i = 0; i < v->n; ++i) f.mvuRight[i] = v->uright[i];
    f.mDescriptors = cv::Mat(v->n, 32, CV_8U, (void*)v->desc, 32);
    f.mvScaleFactors.assign(v->scale_factors, v->scale_factors + 16);
    f.mvLevelSigma2.assign(v->level_sigma2, v->level_sigma2 + 16);
    f.mnMinX = v->min_x; f.mnMinY = v->min_y; f.mnMaxX = v->max_x; f.mnMaxY = v->max_y;
    f.mfGridElementWidthInv = v->grid_inv_w; f.mfGridElementHeightInv = v->grid_inv_h;
    f.mbf = v->bf; f.mb = 0.08f;
}

void fill(Frame& f, const FrameView* v) { fill_base(f, v); f.AssignFeaturesToGrid(); }

// the keyframe takes its grid from the frame it is made of (KeyFrame::KeyFrame, src/KeyFrame.cc:173-183)
void fill(KeyFrame& k, const FrameView* v)
{
    fill_base(k, v);
    std::unique_ptr<Frame> F(new Frame());
    fill(*F, v);
    k.mGrid.resize(k.mnGridCols);
    for (int i = 0; i < k.mnGridCols; i++) {
        k.mGrid[i].resize(k.mnGridRows);
        for (int j = 0; j < k.mnGridRows; j++) k.mGrid[i][j] = F->mGrid[i][j];
    }
}

cv::Mat desc_mat(const uint8_t* d) { cv::Mat m(1, 32, CV_8U); std::memcpy(m.data, d, 32); return m; }

}  // namespace

extern "C" {

int ref_hamming256(const uint8_t* a, const uint8_t* b)
{
    return ORBmatcher::DescriptorDistance(cv::Mat(1, 32, CV_8U, (void*)a, 32), cv::Mat(1, 32, CV_8U, (void*)b, 32));
}

// ORBmatcher::SearchByProjection(Frame&, const vector<MapPointPtr>&, th, bFarPoints, thFarPoints)   src/ORBmatcher.cc:71-244
int ref_search_by_projection_map(const FrameView* Fv, const MpQuery* q, int nq, float th, float nn_ratio, int far_points, float th_far,
                                 const uint8_t* claimed_in, int32_t* assign)
{
    GeometricCamera cam;
    Frame F; fill(F, Fv); F.mpCamera = &cam;
    MapPoint pre; pre.nObs = 1;                                   // a map point that already has observations
    F.mvpMapPoints.assign(Fv->n, nullptr);
    if (claimed_in) for (int i = 0; i < Fv->n; ++i) if (claimed_in[i]) F.mvpMapPoints[i] = &pre;
    std::vector<std::unique_ptr<MapPoint>> mps(nq);
    std::vector<MapPointPtr> vp(nq);
    std::unordered_map<MapPointPtr, int> index;
    for (int i = 0; i < nq; ++i) {
        mps[i].reset(new MapPoint());
        MapPoint& m = *mps[i];
        m.mbTrackInView = true; m.mbTrackInViewR = false;
        m.mTrackProjX = q[i].proj_x; m.mTrackProjY = q[i].proj_y; m.mTrackProjXR = q[i].proj_xr; m.mTrackDepth = q[i].track_depth;
        m.mTrackViewCos = q[i].view_cos; m.mnTrackScaleLevel = q[i].level;
        m.nObs = (q[i].flags & 1u) ? 1 : 0;
        m.desc = desc_mat(q[i].desc);
        vp[i] = &m; index[&m] = i;
    }
    ORBmatcher matcher(nn_ratio, true);
    const int n = matcher.SearchByProjection(F, vp, th, far_points != 0, th_far);
    for (int i = 0; i < Fv->n; ++i) { auto it = index.find(F.mvpMapPoints[i]); assign[i] = it == index.end() ? -1 : it->second; }
    return n;
}

// ORBmatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)   src/ORBmatcher.cc:1774-1993
// z[i]: camera-frame depth of query i with (float)(1.0 / z[i]) == q[i].invz (the reference computes invzc from z)
int ref_search_by_projection_last(const FrameView* Cv, const LastQuery* q, const float* z, int nq, float th, int forward, int backward,
                                  int check_ori, const uint8_t* claimed_in, int32_t* assign)
{
    GeometricCamera cam;
    Frame Cur; fill(Cur, Cv); Cur.mpCamera = &cam;
    MapPoint pre; pre.nObs = 1;
    Cur.mvpMapPoints.assign(Cv->n, nullptr);
    if (claimed_in) for (int i = 0; i < Cv->n; ++i) if (claimed_in[i]) Cur.mvpMapPoints[i] = &pre;
    Frame Last; Last.N = nq; Last.Nleft = -1; Last.mpCamera = &cam;
    Last.mvKeys.resize(nq); Last.mvKeysUn.resize(nq); Last.mvbOutlier.assign(nq, false); Last.mvpMapPoints.assign(nq, nullptr);
    std::vector<std::unique_ptr<MapPoint>> mps(nq);
    std::unordered_map<MapPointPtr, int> index;
    for (int i = 0; i < nq; ++i) {
        mps[i].reset(new MapPoint());
        MapPoint& m = *mps[i];
        m.pos = Eigen::Vector3f(q[i].u, q[i].v, z[i]);             // identity pose + identity camera: project(Tcw * pos) = (u, v)
        m.nObs = (q[i].flags & 1u) ? 1 : 0;
        m.desc = desc_mat(q[i].desc);
        Last.mvpMapPoints[i] = &m; index[&m] = i;
        Last.mvKeys[i].octave = q[i].last_octave; Last.mvKeysUn[i].octave = q[i].last_octave;
        Last.mvKeys[i].angle = q[i].angle; Last.mvKeysUn[i].angle = q[i].angle;
    }
    // bForward = tlc(2) > mb, bBackward = -tlc(2) > mb with tlc = Tlw * twc = the last frame's translation here (:1788-1795)
    Last.mTcw.t = Eigen::Vector3f(0.f, 0.f, forward ? 10.f : backward ? -10.f : 0.f);
    ORBmatcher matcher(0.9f, check_ori != 0);
    const int n = matcher.SearchByProjection(Cur, Last, th, false);
    for (int i = 0; i < Cv->n; ++i) { auto it = index.find(Cur.mvpMapPoints[i]); assign[i] = it == index.end() ? -1 : it->second; }
    return n;
}

// ORBmatcher::SearchForTriangulation(KF1, KF2, pairs, bOnlyStereo, bCoarse)   src/ORBmatcher.cc:999-1242
int ref_search_for_triangulation(const FrameView* K1v, const FrameView* K2v, const FeatVec* fv1, const FeatVec* fv2,
                                 const uint8_t* has_mp1, const uint8_t* has_mp2, const float* F12, const float* ep,
                                 int only_stereo, int coarse, int check_ori, int32_t* match12)
{
    GeometricCamera cam1, cam2;
    std::memcpy(cam1.F12, F12, sizeof(cam1.F12));
    KeyFrame K1, K2; fill(K1, K1v); fill(K2, K2v);
    K1.mpCamera = &cam1; K2.mpCamera = &cam2;
    MapPoint some;
    K1.mvpMapPoints.assign(K1v->n, nullptr); K2.mvpMapPoints.assign(K2v->n, nullptr);
    for (int i = 0; i < K1v->n; ++i) if (has_mp1[i]) K1.mvpMapPoints[i] = &some;
    for (int i = 0; i < K2v->n; ++i) if (has_mp2[i]) K2.mvpMapPoints[i] = &some;
    auto load = [](DBoW2::FeatureVector& out, const FeatVec* fv) {
        for (int a = 0; a < fv->n_nodes; ++a)
            for (int k = fv->offsets[a]; k < fv->offsets[a + 1]; ++k) out.addFeature(fv->node_ids[a], (unsigned)fv->features[k]);
    };
    load(K1.mFeatVec, fv1); load(K2.mFeatVec, fv2);
    // epipole = project(T2w * Cw) (:1009-1012): camera centre of KF1 placed at (ep, 1), KF2 at the identity
    K1.mTcw.t = Eigen::Vector3f(-ep[0], -ep[1], -1.f);
    ORBmatcher matcher(0.6f, check_ori != 0);
    std::vector<std::pair<size_t, size_t>> pairs;
    KeyFramePtr p1 = &K1, p2 = &K2;
    const int n = matcher.SearchForTriangulation(p1, p2, pairs, only_stereo != 0, coarse != 0);
    for (int i = 0; i < K1v->n; ++i) match12[i] = -1;
    for (auto& pr : pairs) match12[pr.first] = (int32_t)pr.second;
    return n;
}

// ORBmatcher::SearchByProjection(Frame& Cur, KeyFramePtr& pKF, const set<MapPointPtr>& sAlreadyFound, th, ORBdist)   src/ORBmatcher.cc:1996-2122
// every third query is handed over as "already found" when skip_found != 0 (it must then not appear in the result)
int ref_search_by_projection_reloc(const FrameView* Cv, const LastQuery* q, int nq, float th, int orb_dist, int check_ori,
                                   const uint8_t* claimed_in, int32_t* assign)
{
    GeometricCamera cam;
    Frame Cur; fill(Cur, Cv); Cur.mpCamera = &cam;
    MapPoint pre;
    Cur.mvpMapPoints.assign(Cv->n, nullptr);
    if (claimed_in) for (int i = 0; i < Cv->n; ++i) if (claimed_in[i]) Cur.mvpMapPoints[i] = &pre;
    KeyFrame K; K.N = nq; K.mpCamera = &cam;
    K.mvKeysUn.resize(nq); K.mvpMapPoints.assign(nq, nullptr);
    std::vector<std::unique_ptr<MapPoint>> mps(nq);
    std::unordered_map<MapPointPtr, int> index;
    for (int i = 0; i < nq; ++i) {
        mps[i].reset(new MapPoint());
        MapPoint& m = *mps[i];
        m.pos = Eigen::Vector3f(q[i].u, q[i].v, 1.f);
        m.predictedLevel = q[i].last_octave;
        m.desc = desc_mat(q[i].desc);
        K.mvpMapPoints[i] = &m; index[&m] = i;
        K.mvKeysUn[i].angle = q[i].angle;
    }
    ORBmatcher matcher(0.9f, check_ori != 0);
    std::set<MapPointPtr> found;
    KeyFramePtr pk = &K;
    const int n = matcher.SearchByProjection(Cur, pk, found, th, orb_dist);
    for (int i = 0; i < Cv->n; ++i) { auto it = index.find(Cur.mvpMapPoints[i]); assign[i] = it == index.end() ? -1 : it->second; }
    return n;
}

// ORBmatcher::SearchByBoW(KeyFramePtr& pKF, Frame& F, vector<MapPointPtr>& vpMapPointMatches)   src/ORBmatcher.cc:300-506
int ref_search_by_bow(const FrameView* Kv, const FrameView* Fv, const FeatVec* fvK, const FeatVec* fvF, const uint8_t* has_mp,
                      float nn_ratio, int check_ori, int32_t* match_f)
{
    GeometricCamera cam;
    KeyFrame K; fill(K, Kv); K.mpCamera = &cam;
    Frame F; fill(F, Fv); F.mpCamera = &cam;
    std::vector<std::unique_ptr<MapPoint>> mps(Kv->n);
    std::unordered_map<MapPointPtr, int> index;
    K.mvpMapPoints.assign(Kv->n, nullptr);
    for (int i = 0; i < Kv->n; ++i) if (has_mp[i]) { mps[i].reset(new MapPoint()); K.mvpMapPoints[i] = mps[i].get(); index[mps[i].get()] = i; }
    auto load = [](DBoW2::FeatureVector& out, const FeatVec* fv) {
        for (int a = 0; a < fv->n_nodes; ++a)
            for (int k = fv->offsets[a]; k < fv->offsets[a + 1]; ++k) out.addFeature(fv->node_ids[a], (unsigned)fv->features[k]);
    };
    load(K.mFeatVec, fvK); load(F.mFeatVec, fvF);
    ORBmatcher matcher(nn_ratio, check_ori != 0);
    std::vector<MapPointPtr> matches;
    KeyFramePtr pk = &K;
    const int n = matcher.SearchByBoW(pk, F, matches);
    for (int i = 0; i < Fv->n; ++i) { auto it = index.find(matches[i]); match_f[i] = it == index.end() ? -1 : it->second; }
    return n;
}

// ORBmatcher::SearchByBoW(KeyFramePtr& pKF1, KeyFramePtr& pKF2, vector<MapPointPtr>& vpMatches12)   src/ORBmatcher.cc:853-997
int ref_search_by_bow_kf(const FrameView* K1v, const FrameView* K2v, const FeatVec* fv1, const FeatVec* fv2, const uint8_t* has1, const uint8_t* has2,
                         float nn_ratio, int check_ori, int32_t* match12)
{
    GeometricCamera cam;
    KeyFrame K1, K2; fill(K1, K1v); fill(K2, K2v);
    K1.mpCamera = &cam; K2.mpCamera = &cam;
    std::vector<std::unique_ptr<MapPoint>> m1(K1v->n), m2(K2v->n);
    std::unordered_map<MapPointPtr, int> idx2;
    K1.mvpMapPoints.assign(K1v->n, nullptr); K2.mvpMapPoints.assign(K2v->n, nullptr);
    for (int i = 0; i < K1v->n; ++i) if (has1[i]) { m1[i].reset(new MapPoint()); K1.mvpMapPoints[i] = m1[i].get(); }
    for (int i = 0; i < K2v->n; ++i) if (has2[i]) { m2[i].reset(new MapPoint()); K2.mvpMapPoints[i] = m2[i].get(); idx2[m2[i].get()] = i; }
    auto load = [](DBoW2::FeatureVector& out, const FeatVec* fv) {
        for (int a = 0; a < fv->n_nodes; ++a)
            for (int k = fv->offsets[a]; k < fv->offsets[a + 1]; ++k) out.addFeature(fv->node_ids[a], (unsigned)fv->features[k]);
    };
    load(K1.mFeatVec, fv1); load(K2.mFeatVec, fv2);
    ORBmatcher matcher(nn_ratio, check_ori != 0);
    std::vector<MapPointPtr> matches;
    KeyFramePtr p1 = &K1, p2 = &K2;
    const int n = matcher.SearchByBoW(p1, p2, matches);
    for (int i = 0; i < K1v->n; ++i) { auto it = idx2.find(matches[i]); match12[i] = it == idx2.end() ? -1 : it->second; }
    return n;
}

// ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight=false)   src/ORBmatcher.cc:1244-1435
// z[i]: camera-frame depth of query i; the harness camera has bf = `bf` and the reference computes ur = u - bf * (1/z), which the
// caller made equal to q[i].ur.  fused_idx[i] = keypoint the map point was fused into (bestDist <= TH_LOW), else -1.
int ref_fuse(const FrameView* Kv, const float* inv_level_sigma2, const FuseQuery* q, const float* z, float bf, int nq, float th, int32_t* fused_idx)
{
    GeometricCamera cam;
    KeyFrame K; fill(K, Kv); K.mpCamera = &cam; K.mbf = bf;
    K.mvInvLevelSigma2.assign(inv_level_sigma2, inv_level_sigma2 + Kv->nlevels);
    K.mvInvLevelSigma2.resize(16, 1.f);
    K.mvpMapPoints.assign(Kv->n, nullptr);
    std::vector<std::unique_ptr<MapPoint>> mps(nq);
    std::vector<MapPointPtr> vp(nq);
    for (int i = 0; i < nq; ++i) {
        mps[i].reset(new MapPoint());
        MapPoint& m = *mps[i];
        m.pos = Eigen::Vector3f(q[i].u, q[i].v, z[i]);
        m.normal = m.pos.normalized();                            // passes the viewing-angle gate (:1331-1338), a caller-side gate of the C ABI
        m.predictedLevel = q[i].level;
        m.nObs = 1 + (i % 3);
        m.desc = desc_mat(q[i].desc);
        vp[i] = &m;
    }
    ORBmatcher matcher(0.6f, true);
    KeyFramePtr pk = &K;
    const int n = matcher.Fuse(pk, vp, th, false);
    std::unordered_map<MapPointPtr, int> where;                   // map point held by the keyframe -> keypoint index
    for (int i = 0; i < Kv->n; ++i) if (K.mvpMapPoints[i]) where[K.mvpMapPoints[i]] = i;
    for (int i = 0; i < nq; ++i) {
        const MapPoint& m = *mps[i];
        if (m.addedIdx >= 0) fused_idx[i] = m.addedIdx;
        else if (m.fusedWith && where.count(m.fusedWith)) fused_idx[i] = where[m.fusedWith];
        else fused_idx[i] = -1;
    }
    return n;
}

// ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming)   src/ORBmatcher.cc:509-615   (Scw = identity)
int ref_search_by_projection_sim3(const FrameView* Kv, const LastQuery* q, int nq, float th, float ratio_hamming, const uint8_t* matched_in, int32_t* assign)
{
    GeometricCamera cam;
    KeyFrame K; fill(K, Kv); K.mpCamera = &cam;
    K.mvpMapPoints.assign(Kv->n, nullptr);
    MapPoint pre;
    std::vector<MapPointPtr> matched(Kv->n, nullptr);
    if (matched_in) for (int i = 0; i < Kv->n; ++i) if (matched_in[i]) matched[i] = &pre;
    std::vector<std::unique_ptr<MapPoint>> mps(nq);
    std::vector<MapPointPtr> vp(nq);
    std::unordered_map<MapPointPtr, int> index;
    for (int i = 0; i < nq; ++i) {
        mps[i].reset(new MapPoint());
        MapPoint& m = *mps[i];
        m.pos = Eigen::Vector3f(q[i].u, q[i].v, 1.f);
        m.normal = m.pos.normalized();
        m.predictedLevel = q[i].last_octave;
        m.desc = desc_mat(q[i].desc);
        vp[i] = &m; index[&m] = i;
    }
    ORBmatcher matcher(0.75f, true);
    Sophus::Sim3f Scw;
    KeyFramePtr pk = &K;
    const int n = matcher.SearchByProjection(pk, Scw, vp, matched, (int)th, ratio_hamming);
    for (int i = 0; i < Kv->n; ++i) { auto it = index.find(matched[i]); assign[i] = it == index.end() ? -1 : it->second; }
    return n;
}

// ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint)   src/ORBmatcher.cc:1437-1553   (Scw = identity similarity)
// pre_mp[i] != 0: the keyframe already holds a map point at keypoint i (=> the vpReplacePoint branch)
int ref_fuse_sim3(const FrameView* Kv, const FuseQuery* q, const float* z, int nq, float th, const uint8_t* pre_mp, int32_t* fused_idx)
{
    GeometricCamera cam;
    KeyFrame K; fill(K, Kv); K.mpCamera = &cam;
    K.mvpMapPoints.assign(Kv->n, nullptr);
    std::vector<std::unique_ptr<MapPoint>> held(Kv->n);
    std::unordered_map<MapPointPtr, int> where;
    if (pre_mp) for (int i = 0; i < Kv->n; ++i) if (pre_mp[i]) { held[i].reset(new MapPoint()); K.mvpMapPoints[i] = held[i].get(); where[held[i].get()] = i; }
    std::vector<std::unique_ptr<MapPoint>> mps(nq);
    std::vector<MapPointPtr> vp(nq), repl(nq, nullptr);
    for (int i = 0; i < nq; ++i) {
        mps[i].reset(new MapPoint());
        MapPoint& m = *mps[i];
        m.pos = Eigen::Vector3f(q[i].u, q[i].v, z[i]);
        m.normal = m.pos.normalized();
        m.predictedLevel = q[i].level;
        m.desc = desc_mat(q[i].desc);
        vp[i] = &m;
    }
    ORBmatcher matcher(0.8f, true);
    Sophus::Sim3f Scw;
    KeyFramePtr pk = &K;
    const int n = matcher.Fuse(pk, Scw, vp, th, repl);
    for (int i = 0; i < Kv->n; ++i) if (K.mvpMapPoints[i] && !where.count(K.mvpMapPoints[i])) where[K.mvpMapPoints[i]] = i;   // added during the call
    for (int i = 0; i < nq; ++i) {
        if (mps[i]->addedIdx >= 0) fused_idx[i] = mps[i]->addedIdx;
        else if (repl[i] && where.count(repl[i])) fused_idx[i] = where[repl[i]];
        else fused_idx[i] = -1;
    }
    return n;
}

// ORBmatcher::SearchBySim3(pKF1, pKF2, vpMatches12, S12, th)   src/ORBmatcher.cc:1555-1772   (identity poses / similarity, fx = fy = 1,
// cx = cy = 0 and depth 1, so a map point's position IS its projection into the other keyframe).  q12[i1] describes map point i1 of
// KF1 projected into KF2 (u, v, level, desc), q21[i2] the reverse; has1 / has2: the keyframe holds a map point there.
// match12[i1] = index in KF2 of the mutually consistent match or -1.
int ref_search_by_sim3(const FrameView* K1v, const FrameView* K2v, const FuseQuery* q12, const FuseQuery* q21, const uint8_t* has1, const uint8_t* has2,
                       float th, int32_t* match12)
{
    GeometricCamera cam;
    KeyFrame K1, K2; fill(K1, K1v); fill(K2, K2v);
    K1.mpCamera = &cam; K2.mpCamera = &cam;
    K1.fx = K1.fy = 1.f; K1.cx = K1.cy = 0.f;
    std::vector<std::unique_ptr<MapPoint>> m1(K1v->n), m2(K2v->n);
    K1.mvpMapPoints.assign(K1v->n, nullptr); K2.mvpMapPoints.assign(K2v->n, nullptr);
    std::unordered_map<MapPointPtr, int> idx2;
    for (int i = 0; i < K1v->n; ++i) if (has1[i]) {
        m1[i].reset(new MapPoint()); MapPoint& m = *m1[i];
        m.pos = Eigen::Vector3f(q12[i].u, q12[i].v, 1.f); m.predictedLevel = q12[i].level; m.desc = desc_mat(q12[i].desc);
        K1.mvpMapPoints[i] = &m;
    }
    for (int i = 0; i < K2v->n; ++i) if (has2[i]) {
        m2[i].reset(new MapPoint()); MapPoint& m = *m2[i];
        m.pos = Eigen::Vector3f(q21[i].u, q21[i].v, 1.f); m.predictedLevel = q21[i].level; m.desc = desc_mat(q21[i].desc);
        K2.mvpMapPoints[i] = &m; idx2[&m] = i;
    }
    ORBmatcher matcher(0.75f, true);
    std::vector<MapPointPtr> matches(K1v->n, nullptr);
    Sophus::Sim3f S12;
    KeyFramePtr p1 = &K1, p2 = &K2;
    const int n = matcher.SearchBySim3(p1, p2, matches, S12, th);
    for (int i = 0; i < K1v->n; ++i) { auto it = idx2.find(matches[i]); match12[i] = it == idx2.end() ? -1 : it->second; }
    return n;
}

// ORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)   src/ORBmatcher.cc:732-852
int ref_search_for_initialization(const FrameView* F1v, const FrameView* F2v, float* prev_matched, int window_size, float nn_ratio, int check_ori,
                                  int32_t* matches12)
{
    Frame F1; fill(F1, F1v);
    Frame F2; fill(F2, F2v);
    std::vector<cv::Point2f> prev(F1v->n);
    for (int i = 0; i < F1v->n; ++i) prev[i] = cv::Point2f(prev_matched[2 * i], prev_matched[2 * i + 1]);
    std::vector<int> m12;
    ORBmatcher matcher(nn_ratio, check_ori != 0);
    const int n = matcher.SearchForInitialization(F1, F2, prev, m12, window_size);
    for (int i = 0; i < F1v->n; ++i) { matches12[i] = m12[i]; prev_matched[2 * i] = prev[i].x; prev_matched[2 * i + 1] = prev[i].y; }
    return n;
}

// MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:389-461) on n observed descriptors (one keyframe per observation; the
// keyframes sit in one array, so the std::map keyed by KeyFrame* iterates them in index order).  Writes the chosen descriptor.
void ref_distinctive_descriptor(const uint8_t* desc, int n, uint8_t* chosen)
{
    std::unique_ptr<KeyFrame[]> kfs(new KeyFrame[n > 0 ? n : 1]);
    MapPoint mp;
    for (int i = 0; i < n; ++i) {
        kfs[i].mDescriptors = desc_mat(desc + (size_t)32 * i);
        mp.mObservations[&kfs[i]] = std::tuple<int, int>(0, -1);
    }
    mp.mDescriptor = cv::Mat(1, 32, CV_8U); std::memset(mp.mDescriptor.data, 0, 32);
    mp.ComputeDistinctiveDescriptors();
    std::memcpy(chosen, mp.mDescriptor.data, 32);
}

// Frame::ComputeStereoFromRGBD (src/Frame.cc:2251-2279): keys = 7 floats per keypoint (mvKeys == mvKeysUn: no distortion)
void ref_stereo_from_rgbd(const float* keys, int n, const float* depth, int w, int h, float bf, float* uright, float* kdepth)
{
    Frame F; F.N = n; F.mbf = bf;
    F.mvKeys.resize(n);
    for (int i = 0; i < n; ++i) F.mvKeys[i] = cv::KeyPoint(keys[7 * i], keys[7 * i + 1], keys[7 * i + 2]);
    F.mvKeysUn = F.mvKeys;
    cv::Mat im(h, w, 0, (void*)depth, (size_t)w * 4);
    F.ComputeStereoFromRGBD(im);
    for (int i = 0; i < n; ++i) { uright[i] = F.mvuRight[i]; kdepth[i] = F.mvDepth[i]; }
}

}  // extern "C"
