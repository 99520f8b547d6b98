// TEST INFRASTRUCTURE ONLY -- C driver around the REFERENCE's own DBoW2 (Thirdparty/DBoW2/DBoW2: TemplatedVocabulary.h, FORB.cpp, BowVector.cpp,
// FeatureVector.cpp, ScoringObject.cpp), compiled where it lies by oracle/ref_build.py into oracle/_ref/libbow_ref.so against the OpenCV stand-in.
// ORBVocabulary is the typedef of include/ORBVocabulary.h; the vocabulary is read with the reference's own loadFromTextFile (the format of ORBvoc.txt),
// and Frame::ComputeBoW's call (src/Frame.cc:1498-1505: transform(vCurrentDesc, mBowVec, mFeatVec, 4)) is flattened into arrays.
#include "TemplatedVocabulary.h"
#include "FORB.h"

#include <cstdint>
#include <cstring>

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabularyBase;
struct ORBVocabulary : ORBVocabularyBase { using ORBVocabularyBase::transform; };      // the single-feature overload is protected

extern "C" {

void* ref_voc_load(const char* path)
{
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void ref_voc_destroy(void* h) { delete (ORBVocabulary*)h; }
int ref_voc_size(void* h) { return (int)((ORBVocabulary*)h)->size(); }

// per feature: the single-feature transform (word, weight, node at `levelsup`); then the whole-frame transform: BowVector (ascending word ids with
// their normalised values) and FeatureVector (ascending node ids, each with its feature indices).  Returns the BowVector size; *n_nodes the
// FeatureVector size; fv_offsets has n_nodes + 1 entries.
int ref_voc_transform(void* h, const uint8_t* desc, int n, int levelsup, uint32_t* word, double* weight, uint32_t* node,
                      uint32_t* bow_ids, double* bow_vals, int* n_nodes, uint32_t* fv_nodes, int32_t* fv_offsets, int32_t* fv_features)
{
    ORBVocabulary* voc = (ORBVocabulary*)h;
    std::vector<cv::Mat> feats(n);
    for (int i = 0; i < n; ++i) { feats[i] = cv::Mat(1, 32, CV_8U); std::memcpy(feats[i].data, desc + (size_t)i * 32, 32); }
    for (int i = 0; i < n; ++i) {
        DBoW2::WordId id; DBoW2::WordValue w; DBoW2::NodeId nid;
        voc->transform(feats[i], id, w, &nid, levelsup);
        word[i] = id; weight[i] = w; node[i] = nid;
    }
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    voc->transform(feats, bv, fv, levelsup);
    int k = 0;
    for (auto& e : bv) { bow_ids[k] = e.first; bow_vals[k] = e.second; ++k; }
    int m = 0, t = 0;
    fv_offsets[0] = 0;
    for (auto& e : fv) { fv_nodes[m] = e.first; for (unsigned f : e.second) fv_features[t++] = (int32_t)f; fv_offsets[++m] = t; }
    *n_nodes = m;
    return k;
}

}  // extern "C"
