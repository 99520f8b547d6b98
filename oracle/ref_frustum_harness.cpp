// TEST INFRASTRUCTURE ONLY -- C driver around the REFERENCE's own Frame::isInFrustum (with MapPoint::PredictScale and Pinhole::project),
// compiled from text sliced out of the reference at build time (see oracle/plvs_standin/plvs_frustum_types.hpp, force-included).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

using namespace std;       // in effect in MapPoint.cc / Frame.cc as well (TemplatedVocabulary.h:36, Frame.cc:52)

namespace PLVS2 {
#include "gen/frustum_slices.inc"
}

using namespace PLVS2;

namespace {
struct MapPointC { float xw[3], normal[3], min_dist, max_dist; uint32_t flags; uint8_t desc[32]; };     // == plvs_map_point
struct FrustumC { float Rcw[9], tcw[3], Ow[3], fx, fy, cx, cy, bf, viewing_cos_limit, scale_factor; int32_t nlevels; float min_x, min_y, max_x, max_y; };
struct MpQuery { float proj_x, proj_y, proj_xr, track_depth, view_cos; int32_t level; uint32_t flags; uint8_t desc[32]; };
}

extern "C" {

// pts[i].min_dist / max_dist are mfMinDistance / mfMaxDistance (the getters apply 0.8 / 1.2); returns the number of points in view
int ref_in_frustum(const FrustumC* fr, const MapPointC* pts, int n, MpQuery* q, uint8_t* in_view)
{
    Pinhole cam; cam.mvParameters = {fr->fx, fr->fy, fr->cx, fr->cy};
    Frame F;
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) F.mRcw(i, j) = fr->Rcw[3 * i + j]; F.mtcw(i) = fr->tcw[i]; F.mOw(i) = fr->Ow[i]; }
    F.mpCamera = &cam; F.mbf = fr->bf;
    F.mnMinX = fr->min_x; F.mnMinY = fr->min_y; F.mnMaxX = fr->max_x; F.mnMaxY = fr->max_y;
    F.mfLogScaleFactor = log(fr->scale_factor);            // Frame.cc:239: `mfLogScaleFactor = log(mfScaleFactor)` on a float, under `using namespace std`
    F.mnScaleLevels = fr->nlevels;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        MapPoint mp;
        mp.mWorldPos = Eigen::Vector3f(pts[i].xw[0], pts[i].xw[1], pts[i].xw[2]);
        mp.mNormalVector = Eigen::Vector3f(pts[i].normal[0], pts[i].normal[1], pts[i].normal[2]);
        mp.mfMinDistance = pts[i].min_dist; mp.mfMaxDistance = pts[i].max_dist;
        MapPointPtr p = &mp;
        const bool in = F.isInFrustum(p, fr->viewing_cos_limit);
        in_view[i] = in ? 1 : 0; cnt += in;
        MpQuery& o = q[i];
        o.proj_x = mp.mTrackProjX; o.proj_y = mp.mTrackProjY; o.proj_xr = mp.mTrackProjXR; o.track_depth = mp.mTrackDepth; o.view_cos = mp.mTrackViewCos;
        o.level = mp.mnTrackScaleLevel; o.flags = pts[i].flags; std::memcpy(o.desc, pts[i].desc, 32);
    }
    return cnt;
}

}  // extern "C"
