// TEST INFRASTRUCTURE ONLY -- stand-in classes for compiling the REFERENCE's own text of
//   Frame::isInFrustum(MapPointPtr&, float)                       src/Frame.cc:955-1030
//   MapPoint::PredictScale(const float&, Frame*)                  src/MapPoint.cc:598-613
//   MapPoint::GetWorldPos / GetNormal / Get{Min,Max}DistanceInvariance   src/MapPoint.cc:160-168, 569-579
//   Pinhole::project(const Eigen::Vector3f&) const                src/CameraModels/Pinhole.cpp:61-67
// which oracle/ref_build.py slices out at build time (oracle/_ref/gen/, deleted after linking) into oracle/_ref/libfrustum_ref.so.
// A separate library from libmatch_ref.so because there MapPoint::PredictScale is a harness hook (the C ABI of the searches receives
// the predicted level); here it is the reference's.  The classes only carry the members those functions touch and declare them.
// Note that `log`/`ceil` in PredictScale are the float overloads: `using namespace std` reaches MapPoint.cc through
// Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:36, so the slice is compiled under the same directive.
#ifndef PLVS_B200_FRUSTUM_TYPES_STANDIN
#define PLVS_B200_FRUSTUM_TYPES_STANDIN
#include <cmath>
#include <mutex>
#include <vector>
#include <Eigen/Core>
#include <Eigen/Geometry>

namespace PLVS2 {

class Frame;
class MapPoint;
class MapLine;
typedef MapPoint* MapPointPtr;
typedef MapLine* MapLinePtr;

class GeometricCamera {
public:
    virtual ~GeometricCamera() {}
    virtual Eigen::Vector2f project(const Eigen::Vector3f& v3D) const = 0;
};

class Pinhole : public GeometricCamera {
public:
    std::vector<float> mvParameters;           // fx, fy, cx, cy
    Eigen::Vector2f project(const Eigen::Vector3f& v3D) const;
};

class MapPoint {
public:
    // tracking variables (include/MapPoint.h)
    float mTrackProjX = 0, mTrackProjY = 0, mTrackDepth = 0, mTrackDepthR = 0, mTrackProjXR = 0, mTrackProjYR = 0;
    bool mbTrackInView = false, mbTrackInViewR = false;
    int mnTrackScaleLevel = 0, mnTrackScaleLevelR = 0;
    float mTrackViewCos = 0, mTrackViewCosR = 0;
    // position, normal, scale-invariance distances
    Eigen::Vector3f mWorldPos, mNormalVector;
    float mfMinDistance = 0, mfMaxDistance = 0;
    std::mutex mMutexPos;
    Eigen::Vector3f GetWorldPos();
    Eigen::Vector3f GetNormal();
    float GetMinDistanceInvariance();
    float GetMaxDistanceInvariance();
    int PredictScale(const float& currentDist, Frame* pF);
};

class Frame {
public:
    int Nleft = -1;
    Eigen::Matrix<float, 3, 3> mRcw; Eigen::Matrix<float, 3, 1> mtcw, mOw;
    GeometricCamera* mpCamera = nullptr; GeometricCamera* mpCamera2 = nullptr;
    float mnMinX = 0, mnMinY = 0, mnMaxX = 0, mnMaxY = 0, mbf = 0;
    float mfLogScaleFactor = 0; int mnScaleLevels = 0;
    bool isInFrustum(MapPointPtr& pMP, float viewingCosLimit);
    bool isInFrustumChecks(MapPointPtr, float, bool bRight = false) { (void)bRight; return false; }     // fisheye pair only (Nleft != -1)
};

}  // namespace PLVS2
#endif
