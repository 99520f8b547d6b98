// TEST INFRASTRUCTURE ONLY.  C entry point around the reference's own line-descriptor matcher, compiled unmodified by oracle/ref_build.py::build_linematch
// from /root/reference/Thirdparty/line_descriptor/src/binary_descriptor_matcher_custom.cpp (class declaration sliced out of
// include/line_descriptor/descriptor_custom.hpp at build time).  What the harness does is what LineMatcher::ComputeDescriptorMatches does around
// the library (/root/reference/src/LineMatcher.cc:2567-2615): knnMatch(query, train, matches, 2, queryMask, true) and the ratio test.
#include "line_descriptor_custom.hpp"
#include <vector>

extern "C" int ref_line_knn2(const unsigned char* q, int nq, const unsigned char* t, int nt, const unsigned char* mask /*nq or NULL*/, float nn_ratio,
                             int* query_idx, int* train_idx /*2 per row*/, float* dist /*2 per row*/, unsigned char* valid, int* n_valid)
{
    using namespace cv;
    Mat Q(nq, 32, CV_8UC1, (void*)q), T(nt, 32, CV_8UC1, (void*)t), M;
    if (mask) M = Mat(nq, 1, CV_8UC1, (void*)mask);
    Ptr<line_descriptor_c::BinaryDescriptorMatcher> bdm = line_descriptor_c::BinaryDescriptorMatcher::createBinaryDescriptorMatcher();
    std::vector<std::vector<DMatch> > lm;
    bdm->knnMatch(Q, T, lm, 2, M, true);
    int nv = 0;
    for (size_t i = 0; i < lm.size(); ++i) {
        query_idx[i] = lm[i].size() ? lm[i][0].queryIdx : -1;
        for (int k = 0; k < 2; ++k) {
            train_idx[2 * i + k] = (int)lm[i].size() > k ? lm[i][k].trainIdx : -1;
            dist[2 * i + k] = (int)lm[i].size() > k ? lm[i][k].distance : -1.f;
        }
        bool ok;
        if (lm[i].size() > 1) ok = lm[i][0].distance < nn_ratio * lm[i][1].distance;      // src/LineMatcher.cc:2598-2611
        else ok = true;
        valid[i] = ok; nv += ok;
    }
    *n_valid = nv;
    return (int)lm.size();
}
