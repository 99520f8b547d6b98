// TEST INFRASTRUCTURE ONLY -- C driver around the REFERENCE's own src/ORBextractor.cc.
//
// oracle/ref_build.py compiles /root/reference/src/ORBextractor.cc where it lies (nothing is copied) together with this
// file into oracle/_ref/liborb_ref.so, against the OpenCV stand-in of oracle/cv_standin/ (OpenCV's C++ headers are not
// installed; see that header for what is restated there -- cv::Mat bookkeeping and the four OpenCV primitives, which are
// forwarded to the C restatements that tests/test_oracle_orb.py pins bit-exactly to the real cv2).  This file only calls
// PLVS2::ORBextractor the way Frame::ExtractORB does (src/Frame.cc:806-813) and flattens the results.
#include <opencv2/opencv.hpp>
#include "ORBextractor.h"

#include <cstring>
#include <vector>

extern "C" {

void* ref_orb_create(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST)
{
    return new PLVS2::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST);
}
void ref_orb_destroy(void* h) { delete (PLVS2::ORBextractor*)h; }

// keypoints as 7 floats each (x, y, size, angle, response, octave, class_id) like orc_orb_extract; returns n (or -needed)
int ref_orb_extract(void* h, const uint8_t* gray, int w, int ht, int stride, int lap0, int lap1,
                    float* kps, uint8_t* desc, int cap, int* mono_index)
{
    PLVS2::ORBextractor* ex = (PLVS2::ORBextractor*)h;
    cv::Mat image(ht, w, CV_8UC1, (void*)gray, (size_t)stride);
    std::vector<cv::KeyPoint> keys;
    cv::Mat descriptors;
    std::vector<int> lap = {lap0, lap1};
    const int mono = (*ex)(image, cv::Mat(), keys, descriptors, lap);
    if (mono_index) *mono_index = mono;
    const int n = (int)keys.size();
    if (n > cap) return -n;
    for (int i = 0; i < n; ++i) {
        float* o = kps + 7 * i;
        o[0] = keys[i].pt.x; o[1] = keys[i].pt.y; o[2] = keys[i].size; o[3] = keys[i].angle; o[4] = keys[i].response;
        o[5] = (float)keys[i].octave; o[6] = (float)keys[i].class_id;
        std::memcpy(desc + 32 * (size_t)i, descriptors.ptr(i), 32);
    }
    return n;
}

// constructor tables through the public getters (include/ORBextractor.h:90-113)
void ref_orb_tables(void* h, float* scale, float* inv_scale, float* sigma2, float* inv_sigma2)
{
    PLVS2::ORBextractor* ex = (PLVS2::ORBextractor*)h;
    const int n = ex->GetLevels();
    const std::vector<float> a = ex->GetScaleFactors(), b = ex->GetInverseScaleFactors(), c = ex->GetScaleSigmaSquares(), d = ex->GetInverseScaleSigmaSquares();
    for (int i = 0; i < n; ++i) { scale[i] = a[i]; inv_scale[i] = b[i]; sigma2[i] = c[i]; inv_sigma2[i] = d[i]; }
}

// public members mvImagePyramid / mvImagePyramidFiltered (include/ORBextractor.h:125-127) after the last extraction
int ref_orb_level(void* h, int level, int filtered, uint8_t* out, int cap, int* w, int* ht)
{
    PLVS2::ORBextractor* ex = (PLVS2::ORBextractor*)h;
    const cv::Mat& m = filtered ? ex->mvImagePyramidFiltered[level] : ex->mvImagePyramid[level];
    *w = m.cols; *ht = m.rows;
    if (!out) return 0;
    if (cap < m.cols * m.rows) return -1;
    for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr(y), m.cols);
    return 0;
}

}  // extern "C"
