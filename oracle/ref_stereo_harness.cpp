// TEST INFRASTRUCTURE ONLY -- C driver around the REFERENCE's own Frame::ComputeStereoMatches (src/Frame.cc:1780-1983).
//
// Frame.cc as a whole needs the entire system (g2o, IMU, line features, Tracking ...), so it cannot be compiled.  oracle/ref_build.py
// therefore extracts exactly the definition of Frame::ComputeStereoMatches from /root/reference/src/Frame.cc at BUILD time (from the line
// `void Frame::ComputeStereoMatches()` to its matching closing brace) into the git-ignored build directory oracle/_ref/gen/ -- nothing of it
// is stored in this repository -- and this file includes that slice, so the function body that runs is the reference's text, compiled
// against the stand-in Frame of oracle/plvs_standin/plvs_types.hpp.  The two extractors are the reference's own ORBextractor.cc: their
// pyramids (levels inside the 19-px BORDER_REFLECT_101 frame), keypoints and descriptors feed the function exactly as in
// Frame::Frame(stereo) (src/Frame.cc:314-330).
#include "ORBextractor.h"
#include "ORBmatcher.h"

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <vector>

using namespace std;      // src/Frame.cc:52 does the same; the slice relies on it (round, ceil, sort, pair, vector)

namespace PLVS2 {
#include "gen/frame_stereo_slice.inc"
}

using namespace PLVS2;

extern "C" {

// left / right: 8-bit rectified images.  Outputs: the left keypoints (7 floats each, like ref_orb_extract) so the caller can check that its
// own extraction is identical, mvuRight and mvDepth.  Returns N (left keypoints) or -needed.
int ref_compute_stereo_matches(const uint8_t* left, const uint8_t* right, int w, int h, int stride, int nfeatures, float scaleFactor, int nlevels,
                               int iniTh, int minTh, float mb, float mbf, float* keysL_out, int cap, float* uRight, float* depth, int* n_right)
{
    ORBextractor exL(nfeatures, scaleFactor, nlevels, iniTh, minTh), exR(nfeatures, scaleFactor, nlevels, iniTh, minTh);
    Frame F;
    std::vector<int> lap = {0, 0};
    cv::Mat imL(h, w, CV_8UC1, (void*)left, (size_t)stride), imR(h, w, CV_8UC1, (void*)right, (size_t)stride);
    exL(imL, cv::Mat(), F.mvKeys, F.mDescriptors, lap);
    exR(imR, cv::Mat(), F.mvKeysRight, F.mDescriptorsRight, lap);
    F.N = (int)F.mvKeys.size();
    if (n_right) *n_right = (int)F.mvKeysRight.size();
    if (F.N > cap) return -F.N;
    F.mpORBextractorLeft = &exL; F.mpORBextractorRight = &exR;
    F.mvScaleFactors = exL.GetScaleFactors(); F.mvInvScaleFactors = exL.GetInverseScaleFactors();
    F.mb = mb; F.mbf = mbf;
    F.ComputeStereoMatches();
    for (int i = 0; i < F.N; ++i) {
        float* o = keysL_out + 7 * i;
        o[0] = F.mvKeys[i].pt.x; o[1] = F.mvKeys[i].pt.y; o[2] = F.mvKeys[i].size; o[3] = F.mvKeys[i].angle; o[4] = F.mvKeys[i].response;
        o[5] = (float)F.mvKeys[i].octave; o[6] = (float)F.mvKeys[i].class_id;
        uRight[i] = F.mvuRight[i]; depth[i] = F.mvDepth[i];
    }
    return F.N;
}

}  // extern "C"
