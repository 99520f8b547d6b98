// TEST INFRASTRUCTURE ONLY -- gives the stand-in Frame / KeyFrame of oracle/plvs_standin/plvs_types.hpp the REFERENCE's own definitions of
//   Frame::AssignFeaturesToGrid   src/Frame.cc:716-805        Frame::PosInGrid           src/Frame.cc:1305-1316
//   Frame::GetFeaturesInArea      src/Frame.cc:1231-1303      KeyFrame::GetFeaturesInArea src/KeyFrame.cc:1179-1229
//   Frame::ComputeStereoFromRGBD  src/Frame.cc:2251-2279      MapPoint::ComputeDistinctiveDescriptors src/MapPoint.cc:389-461
// oracle/ref_build.py slices them out of the reference at build time into oracle/_ref/gen/ (git-ignored; nothing is stored in this
// repository) and this file includes the slices.  (plvs_types.hpp is force-included by the build.)
#include "ORBmatcher.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

using namespace std;      // src/Frame.cc:52, src/KeyFrame.cc do the same

namespace PLVS2 {
#include "gen/frame_grid_slices.inc"
#include "gen/keyframe_grid_slice.inc"
#include "gen/mappoint_slice.inc"
}
