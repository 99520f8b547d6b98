// TEST INFRASTRUCTURE ONLY -- CPU oracle for the bag-of-words transform (Frame::ComputeBoW, src/Frame.cc:1498-1505 ->
// ORBVocabulary::transform(features, BowVector&, FeatureVector&, levelsup = 4)).  DBoW2 is an in-tree third-party library of the reference
// (Thirdparty/DBoW2); restated here on flat arrays:
//   TemplatedVocabulary::loadFromTextFile         Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1351-1436  (the ORBvoc.txt format: "k L scoring weighting",
//                                                 then one line per node: parent, leaf flag, 32 descriptor bytes, weight; ids in file order from 1)
//   transform(feature, word, weight, nid, levelsup)   :1231-1271  (descend: first child with the smallest FORB::distance; node id at level L - levelsup)
//   transform(features, BowVector, FeatureVector, levelsup)   :1140-1207  (TF / TF_IDF: addWeight, IDF / BINARY: addIfNotExist; words of weight 0 dropped;
//                                                 L1 / L2 normalisation in ascending word order, BowVector.cpp:62-84; otherwise division by the size)
//   FORB::distance                                FORB.cpp:82-99  (popcount of the XOR)
// PINNED: tests/test_oracle_vs_reference_bow.py runs the reference's own DBoW2, compiled into oracle/_ref/libbow_ref.so, on the same vocabularies.
// A trailing empty line in the file makes the reference append a node read from failed extractions (uninitialised leaf flag / descriptor); this
// loader ignores empty lines, and the test vocabularies end without one.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct Voc {
    int k = 0, L = 0, scoring = 0, weighting = 0;
    std::vector<int> parent, word_id;
    std::vector<std::vector<int>> children;
    std::vector<uint8_t> desc;
    std::vector<double> weight;
    int n_words = 0;
};

int hamming32(const uint8_t* a, const uint8_t* b)
{
    int d = 0;
    for (int i = 0; i < 32; ++i) d += __builtin_popcount((unsigned)(a[i] ^ b[i]));
    return d;
}

void descend(const Voc& v, const uint8_t* f, int levelsup, unsigned* word, double* weight, unsigned* nid)
{
    const int nid_level = v.L - levelsup;
    if (nid_level <= 0) *nid = 0;
    int final_id = 0, level = 0;
    do {
        ++level;
        const std::vector<int>& ch = v.children[final_id];
        final_id = ch[0];
        double best = hamming32(f, &v.desc[(size_t)final_id * 32]);
        for (size_t c = 1; c < ch.size(); ++c) {
            const double d = hamming32(f, &v.desc[(size_t)ch[c] * 32]);
            if (d < best) { best = d; final_id = ch[c]; }
        }
        if (level == nid_level) *nid = (unsigned)final_id;
    } while (!v.children[final_id].empty());
    *word = (unsigned)v.word_id[final_id];
    *weight = v.weight[final_id];
}

}  // namespace

extern "C" {

void* orc_voc_load(const char* path)
{
    std::ifstream f(path);
    if (!f) return nullptr;
    Voc* v = new Voc();
    std::string line;
    std::getline(f, line);
    { std::stringstream ss(line); ss >> v->k >> v->L >> v->scoring >> v->weighting; }
    v->parent.assign(1, 0); v->word_id.assign(1, -1); v->children.resize(1); v->desc.assign(32, 0); v->weight.assign(1, 0.0);
    while (std::getline(f, line)) {
        if (line.empty()) continue;
        std::stringstream ss(line);
        int pid, leaf;
        ss >> pid >> leaf;
        const int nid = (int)v->parent.size();
        v->parent.push_back(pid); v->children.emplace_back(); v->children[pid].push_back(nid);
        for (int i = 0; i < 32; ++i) { int b; ss >> b; v->desc.push_back((uint8_t)b); }
        double w; ss >> w;
        v->weight.push_back(w);
        v->word_id.push_back(leaf > 0 ? v->n_words++ : -1);
    }
    return v;
}
void orc_voc_destroy(void* h) { delete (Voc*)h; }
int orc_voc_size(void* h) { return ((Voc*)h)->n_words; }

// flat copy for the product's plvs_voc_create: parent[n], leaf word id or -1 [n], desc[n*32], weight[n]; returns the node count (root included)
int orc_voc_export(void* h, int32_t* parent, int32_t* word_id, uint8_t* desc, double* weight, int* kLsw)
{
    Voc* v = (Voc*)h;
    const int n = (int)v->parent.size();
    if (kLsw) { kLsw[0] = v->k; kLsw[1] = v->L; kLsw[2] = v->scoring; kLsw[3] = v->weighting; }
    if (parent) for (int i = 0; i < n; ++i) { parent[i] = v->parent[i]; word_id[i] = v->word_id[i]; weight[i] = v->weight[i]; }
    if (desc) std::memcpy(desc, v->desc.data(), (size_t)n * 32);
    return n;
}

int orc_voc_transform(void* h, const uint8_t* desc, int n, int levelsup, uint32_t* word, double* weight, uint32_t* node,
                      uint32_t* bow_ids, double* bow_vals, int* n_nodes, uint32_t* fv_nodes, int32_t* fv_offsets, int32_t* fv_features)
{
    const Voc& v = *(Voc*)h;
    std::map<unsigned, double> bow;
    std::map<unsigned, std::vector<unsigned>> fv;
    const bool tf = v.weighting == 0 || v.weighting == 1;
    for (int i = 0; i < n; ++i) {
        unsigned id, nid = 0; double w;
        descend(v, desc + (size_t)i * 32, levelsup, &id, &w, &nid);
        word[i] = id; weight[i] = w; node[i] = nid;
        if (w > 0) {
            auto it = bow.lower_bound(id);
            if (it != bow.end() && it->first == id) { if (tf) it->second += w; }
            else bow.insert(it, std::make_pair(id, w));
            fv[nid].push_back((unsigned)i);
        }
    }
    // mustNormalize (ScoringObject.h:74-89): L1 for L1_NORM, CHI_SQUARE, KL, BHATTACHARYYA; L2 for L2_NORM; none for DOT_PRODUCT
    const bool must = v.scoring != 5;
    const bool l1 = v.scoring != 1;
    if (tf && !bow.empty() && !must) { const double nd = (double)bow.size(); for (auto& e : bow) e.second /= nd; }
    if (must) {
        double norm = 0.0;
        if (l1) { for (auto& e : bow) norm += std::fabs(e.second); }
        else { for (auto& e : bow) norm += e.second * e.second; norm = std::sqrt(norm); }
        if (norm > 0.0) for (auto& e : bow) e.second /= norm;
    }
    int k = 0;
    for (auto& e : bow) { bow_ids[k] = e.first; bow_vals[k] = e.second; ++k; }
    int m = 0, t = 0;
    fv_offsets[0] = 0;
    for (auto& e : fv) { fv_nodes[m] = e.first; for (unsigned fi : e.second) fv_features[t++] = (int32_t)fi; fv_offsets[++m] = t; }
    *n_nodes = m;
    return k;
}

}  // extern "C"
