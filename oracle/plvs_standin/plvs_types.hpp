// TEST INFRASTRUCTURE ONLY -- stand-ins for the PLVS data model classes that the reference's src/ORBmatcher.cc reads
// (Frame, KeyFrame, MapPoint, GeometricCamera), so that ORBmatcher.cc can be compiled UNMODIFIED, where it lies, into
// oracle/_ref/libmatch_ref.so.  The real headers pull in the whole system (g2o, DBoW2 vocabulary, Boost serialization,
// Sophus -> Eigen, OpenCV); this file is force-included (-include) and pre-defines their include guards, so the
// `#include "MapPoint.h"` etc. inside the reference's ORBmatcher.h become empty.  Pointers.h and DBoW2's FeatureVector.h
// are the reference's own files.
//
// What is the reference's own code after this: every search in ORBmatcher.cc -- windows and level ranges, the claim checks
// on mvpMapPoints (Observations()>0), the right-coordinate gate, best/second-best bookkeeping and ratio test, the
// immediate claim writes, the rotation histogram with ComputeThreeMaxima, the FeatureVector merge-walk, epipole and
// distance gates of SearchForTriangulation, DescriptorDistance.  The frame grid is the reference's own text too: the definitions of
// Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (src/Frame.cc:716-805, 1305-1316, 1231-1303) and
// KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:1179-1229) are sliced out of the reference at build time (oracle/ref_build.py ->
// oracle/_ref/gen/, git-ignored) and compiled by oracle/ref_grid_slices.cpp against the classes below, which only DECLARE them.
// What is restated HERE (with the lines it follows):
//   Pinhole::epipolarConstrain (distance of kp2 to the epipolar line of kp1; the fundamental-matrix product is given)
//                                                                      src/CameraModels/Pinhole.cpp:125-147
//   the grid copy of the KeyFrame constructor (src/KeyFrame.cc:173-183), in the harness
// Poses and the camera are trivial on purpose: the parity harness passes world points that ARE the wanted projections
// (identity pose, project(p) = (p.x, p.y)), because the drop-in C ABI receives projected queries too (the caller-side
// shim projects with the reference's own Sophus/camera code, include/plvs_b200.h plvs_last_query).
#ifndef PLVS_B200_PLVS_TYPES_STANDIN
#define PLVS_B200_PLVS_TYPES_STANDIN
#define MAPPOINT_H
#define KEYFRAME_H
#define FRAME_H
#include <cmath>
#include <map>
#include <mutex>
#include <set>
#include <tuple>
#include <unordered_set>
#include <utility>
#include <vector>
#include <opencv2/opencv.hpp>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include "sophus/se3.hpp"
#include "sophus/sim3.hpp"
#include "Pointers.h"
#include "Thirdparty/DBoW2/DBoW2/FeatureVector.h"

#define FRAME_GRID_ROWS 48
#define FRAME_GRID_COLS 64
// line features: only so that the tail of Frame::AssignFeaturesToGrid (src/Frame.cc:748-805) compiles; mpLineExtractorLeft stays null
#define LINE_THETA_GRID_ROWS 36
#define LINE_D_GRID_COLS 160
namespace cv { namespace line_descriptor_c { struct KeyLine { float startPointX, startPointY, endPointX, endPointY; }; } }

namespace PLVS2 {

using std::vector; using std::pair; using std::set; using std::unordered_set; using std::tuple; using std::get;     // the reference's headers are written inside `using namespace std`-style code (Fuse's signature uses bare `vector`)

class ORBextractor;          // the reference's own class (include/ORBextractor.h), used through its public mvImagePyramid

class GeometricCamera {
public:
    float F12[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};      // row-major fundamental matrix, set by the harness
    virtual ~GeometricCamera() {}
    virtual Eigen::Vector2f project(const Eigen::Vector3f& p) { return Eigen::Vector2f(p(0), p(1)); }
    virtual float uncertainty2(const Eigen::Matrix<double, 2, 1>&) { return 1.0f; }
    // src/CameraModels/Pinhole.cpp:125-147 with F12 precomputed
    virtual bool epipolarConstrain(GeometricCamera*, const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const Eigen::Matrix3f&, const Eigen::Vector3f&,
                                   const float sigmaLevel, const float unc)
    {
        const float a = kp1.pt.x * F12[0] + kp1.pt.y * F12[3] + F12[6];
        const float b = kp1.pt.x * F12[1] + kp1.pt.y * F12[4] + F12[7];
        const float c = kp1.pt.x * F12[2] + kp1.pt.y * F12[5] + F12[8];
        const float num = a * kp2.pt.x + b * kp2.pt.y + c;
        const float den = a * a + b * b;
        if (den == 0) return false;
        const float dsqr = num * num / den;
        return dsqr < 3.84 * unc;
        (void)sigmaLevel;
    }
};

class MapPoint {
public:
    // tracking variables written by Frame::isInFrustum (include/MapPoint.h)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
    float mTrackViewCos = 1, mTrackViewCosR = 1;
    long unsigned int mnId = 0, mnFuseCandidateForKF = 0, mnLastFrameSeen = 0;
    // harness state
    Eigen::Vector3f pos = Eigen::Vector3f::Zero(), normal = Eigen::Vector3f(0, 0, 1);
    cv::Mat desc;
    int nObs = 1; bool bad = false;
    float minDist = 0, maxDist = 1e9f;
    // MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:389-461) is DEFINED by the reference's own text (oracle/ref_build.py slices it)
    std::map<KeyFramePtr, std::tuple<int, int>> mObservations;
    bool mbBad = false;
    std::mutex mMutexFeatures;
    cv::Mat mDescriptor;
    void ComputeDistinctiveDescriptors();
    int predictedLevel = 0;              // what PredictScale returns (the caller-side shim evaluates the real one)
    int addedIdx = -1;                   // Fuse: AddObservation(pKF, idx) was called with this idx
    MapPoint* fusedWith = nullptr;       // Fuse: the keyframe's map point this one was merged with (either Replace direction)

    Eigen::Vector3f GetWorldPos() { return pos; }
    Eigen::Vector3f GetNormal() { return normal; }
    cv::Mat GetDescriptor() { return desc; }
    int Observations() { return nObs; }
    bool isBad() { return bad; }
    float GetMinDistanceInvariance() { return minDist; }
    float GetMaxDistanceInvariance() { return maxDist; }
    int PredictScale(const float&, KeyFramePtr) { return predictedLevel; }
    int PredictScale(const float&, Frame*) { return predictedLevel; }
    bool IsInKeyFrame(KeyFramePtr) { return false; }
    std::tuple<int, int> GetIndexInKeyFrame(const KeyFramePtr&) { return std::tuple<int, int>(-1, -1); }
    void AddObservation(KeyFramePtr, size_t idx) { addedIdx = (int)idx; }
    void Replace(MapPointPtr other) { fusedWith = other; other->fusedWith = this; }
    std::map<KeyFramePtr, std::tuple<int, int>> GetObservations() { return {}; }
};

class FrameBase {          // what Frame and KeyFrame share for the matcher
public:
    int N = 0, Nleft = -1, NLeft = -1;
    std::vector<cv::KeyPoint> mvKeys, mvKeysUn, mvKeysRight;
    std::vector<float> mvuRight, mvDepth;
    cv::Mat mDescriptors;
    std::vector<float> mvScaleFactors, mvLevelSigma2, mvInvLevelSigma2;
    std::vector<int> mvLeftToRightMatch, mvRightToLeftMatch;
    GeometricCamera* mpCamera = nullptr; GeometricCamera* mpCamera2 = nullptr;
    float fx = 1, fy = 1, cx = 0, cy = 0, mbf = 0, mb = 0;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0;                     // Frame: static members, set once (src/Frame.cc:444-462)
    float mfGridElementWidthInv = 0, mfGridElementHeightInv = 0;
    DBoW2::FeatureVector mFeatVec;
    Sophus::SE3f mTcw, mTrl;

    bool IsInImage(const float& x, const float& y) const { return (x >= mnMinX && x < mnMaxX && y >= mnMinY && y < mnMaxY); }      // include/Frame.h / KeyFrame.cc
};

class Frame : public FrameBase {
public:
    std::vector<MapPointPtr> mvpMapPoints;
    std::vector<bool> mvbOutlier;
    // the frame grid: members as in include/Frame.h:345,383,522-524; the three functions are DEFINED by the reference's own text
    std::vector<std::size_t> mGrid[FRAME_GRID_COLS][FRAME_GRID_ROWS], mGridRight[FRAME_GRID_COLS][FRAME_GRID_ROWS];
    void* mpLineExtractorLeft = nullptr;
    int Nlines = 0, NlinesLeft = -1;
    std::vector<std::size_t> mLineGrid[LINE_D_GRID_COLS][LINE_THETA_GRID_ROWS], mLineGridRight[LINE_D_GRID_COLS][LINE_THETA_GRID_ROWS];
    std::vector<cv::line_descriptor_c::KeyLine> mvKeyLinesUn, mvKeyLinesRightUn;
    bool PosLineInGrid(const cv::line_descriptor_c::KeyLine&, int&, int&) { return false; }
    void AssignFeaturesToGrid();
    bool PosInGrid(const cv::KeyPoint& kp, int& posX, int& posY);
    std::vector<std::size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const int minLevel = -1, const int maxLevel = -1, const bool bRight = false) const;
    // members read / written by Frame::ComputeStereoMatches (src/Frame.cc:1780-1983); its body is compiled from the reference
    // (oracle/ref_build.py slices it out of Frame.cc at build time into oracle/_ref/gen/), see oracle/ref_stereo_harness.cpp
    ORBextractor* mpORBextractorLeft = nullptr; ORBextractor* mpORBextractorRight = nullptr;
    cv::Mat mDescriptorsRight;
    std::vector<float> mvInvScaleFactors;
    float mMedianDepth = 0.f;
    bool mbUseFovCentersKfGenCriterion = false;
    float ComputeSceneMedianDepth(int q = 2) { (void)q; return 0.f; }
    void ComputeStereoMatches();
    void ComputeStereoFromRGBD(const cv::Mat& imDepth);      // src/Frame.cc:2251-2279, defined by the reference's own text as well
    Sophus::SE3f GetPose() const { return mTcw; }
    Sophus::SE3f GetRelativePoseTrl() { return mTrl; }
    Sophus::SE3f GetRelativePoseTlr() { return mTrl.inverse(); }
};

class KeyFrame : public FrameBase {
public:
    std::vector<MapPointPtr> mvpMapPoints;
    long unsigned int mnId = 0;
    static constexpr float skFovCenterDistance = 1.5f;      // include/KeyFrame.h; only copied into mMedianDepth
    // include/KeyFrame.h:252-253,436,505; the search is DEFINED by the reference's own text (src/KeyFrame.cc:1179-1229)
    const int mnGridCols = FRAME_GRID_COLS, mnGridRows = FRAME_GRID_ROWS;
    std::vector<std::vector<std::vector<std::size_t>>> mGrid, mGridRight;
    std::vector<std::size_t> GetFeaturesInArea(const float& x, const float& y, const float& r, const bool bRight = false) const;
    bool isBad() { return false; }
    Sophus::SE3f GetPose() { return mTcw; }
    Sophus::SE3f GetPoseInverse() { return mTcw.inverse(); }
    Sophus::SE3f GetRightPose() { return mTrl * mTcw; }
    Sophus::SE3f GetRightPoseInverse() { return (mTrl * mTcw).inverse(); }
    Eigen::Vector3f GetCameraCenter() { return mTcw.inverse().translation(); }
    Eigen::Vector3f GetRightCameraCenter() { return (mTrl * mTcw).inverse().translation(); }
    MapPointPtr GetMapPoint(const size_t& idx) { return mvpMapPoints[idx]; }
    std::vector<MapPointPtr> GetMapPointMatches() { return mvpMapPoints; }
    std::set<MapPointPtr> GetMapPoints() { std::set<MapPointPtr> s; for (MapPointPtr p : mvpMapPoints) if (p) s.insert(p); return s; }
    std::unordered_set<MapPointPtr> GetMapPointsUnordered() { std::unordered_set<MapPointPtr> s; for (MapPointPtr p : mvpMapPoints) if (p) s.insert(p); return s; }
    void AddMapPoint(MapPointPtr p, const size_t& idx) { mvpMapPoints[idx] = p; }
};

}  // namespace PLVS2
#endif
