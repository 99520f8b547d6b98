// Bag-of-words transform behind the C ABI: ORBVocabulary::loadFromTextFile + transform as Frame::ComputeBoW calls it
// (src/Frame.cc:1498-1505; Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1140-1271, 1351-1436).  The tree descents (N x L x k Hamming distances) and
// the FeatureVector grouping run on the device; the BowVector -- a few thousand double additions whose order is part of the result -- is
// accumulated here on the host in the reference's order (BowVector.cpp:34-84).
#include <algorithm>
#include <cmath>
#include <fstream>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>
#include "common.cuh"

using namespace plvs;

namespace {
#include "bow_kernels.cuh"
}

struct plvs_voc {
    int device = 0;
    cudaStream_t stream = nullptr;
    int k = 0, L = 0, scoring = 0, weighting = 0, n_nodes = 0, n_words = 0;
    DevBuf<int> d_child_off, d_child_id, d_word_id;
    DevBuf<uint8_t> d_child_desc;
    DevBuf<double> d_weight;
    // per-call workspace
    DevBuf<uint8_t> d_desc;
    DevBuf<uint32_t> d_word, d_node, d_sorted_node, d_fv_nodes;
    DevBuf<double> d_w;
    DevBuf<int32_t> d_sorted_feat, d_fv_off;
    DevBuf<int> d_cnt;
    PinBuf<int> p_cnt;
    std::mutex mu;
};

extern "C" {

int plvs_voc_create(int device, int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent, const int32_t* word_id, const uint8_t* desc,
                    const double* weight, plvs_voc** out)
{
    if (!out || n_nodes < 1 || !parent || !word_id || !desc || !weight || L < 1 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3) {
        set_error("bad vocabulary arguments"); return PLVS_EINVAL;
    }
    for (int i = 1; i < n_nodes; ++i) if (parent[i] < 0 || parent[i] >= i) { set_error("node %d: a parent must precede its children", i); return PLVS_EINVAL; }
    // children in id order == the order loadFromTextFile pushes them (:1405)
    std::vector<int> off(n_nodes + 1, 0), id(std::max(n_nodes - 1, 1));
    for (int i = 1; i < n_nodes; ++i) ++off[parent[i] + 1];
    for (int i = 0; i < n_nodes; ++i) off[i + 1] += off[i];
    std::vector<int> cur(off.begin(), off.end() - 1);
    std::vector<uint8_t> cdesc((size_t)std::max(n_nodes - 1, 1) * 32);
    for (int i = 1; i < n_nodes; ++i) { const int s = cur[parent[i]]++; id[s] = i; std::memcpy(&cdesc[(size_t)s * 32], desc + (size_t)i * 32, 32); }
    if (off[1] == 0) { set_error("the root has no children"); return PLVS_EINVAL; }
    int words = 0;
    for (int i = 0; i < n_nodes; ++i) {
        const bool leaf = off[i + 1] == off[i];
        if (leaf && word_id[i] < 0) { set_error("node %d has no children and no word id", i); return PLVS_EINVAL; }
        if (word_id[i] >= 0) words = std::max(words, word_id[i] + 1);
    }
    plvs_voc* h = new plvs_voc();
    h->device = device; h->k = k; h->L = L; h->scoring = scoring; h->weighting = weighting; h->n_nodes = n_nodes; h->n_words = words;
    auto fail = [&](int rc) { delete h; return rc; };
    if (cudaSetDevice(device) != cudaSuccess) { set_error("cudaSetDevice(%d) failed", device); return fail(PLVS_ENODEV); }
    if (create_handle_stream(&h->stream, 1) != cudaSuccess) { set_error("stream creation failed"); return fail(PLVS_ENODEV); }
    int rc;
    if ((rc = h->d_child_off.alloc(n_nodes + 1)) || (rc = h->d_child_id.alloc(id.size())) || (rc = h->d_child_desc.alloc(cdesc.size())) ||
        (rc = h->d_word_id.alloc(n_nodes)) || (rc = h->d_weight.alloc(n_nodes)) || (rc = h->d_cnt.alloc(4)) || (rc = h->p_cnt.alloc(4))) return fail(rc);
    if (cudaMemcpy(h->d_child_off.p, off.data(), (size_t)(n_nodes + 1) * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_child_id.p, id.data(), id.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_child_desc.p, cdesc.data(), cdesc.size(), cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_word_id.p, word_id, (size_t)n_nodes * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(h->d_weight.p, weight, (size_t)n_nodes * 8, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("vocabulary upload failed"); return fail(PLVS_ENODEV); }
    *out = h;
    return PLVS_OK;
}

int plvs_voc_load_text(const char* path, int device, plvs_voc** out)
{
    if (!path || !out) { set_error("null argument"); return PLVS_EINVAL; }
    std::ifstream f(path);
    if (!f) { set_error("cannot open %s", path); return PLVS_EINVAL; }
    std::string line;
    std::getline(f, line);
    int k = -1, L = -1, n1 = -1, n2 = -1;
    { std::stringstream ss(line); ss >> k >> L >> n1 >> n2; }
    if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) { set_error("%s: not a vocabulary text file", path); return PLVS_EINVAL; }   // :1372-1376
    std::vector<int32_t> parent(1, 0), word(1, -1);
    std::vector<uint8_t> desc(32, 0);
    std::vector<double> weight(1, 0.0);
    int words = 0;
    while (std::getline(f, line)) {
        if (line.empty()) continue;                 // the reference turns a trailing empty line into a node of unread values; ignored here
        std::stringstream ss(line);
        int pid = -1, leaf = 0;
        ss >> pid >> leaf;
        if (!ss || pid < 0 || pid >= (int)parent.size()) { set_error("%s: bad node line %zu", path, parent.size()); return PLVS_EINVAL; }
        parent.push_back(pid);
        for (int i = 0; i < 32; ++i) { int b = 0; ss >> b; desc.push_back((uint8_t)b); }
        double w = 0.0; ss >> w;
        weight.push_back(w);
        word.push_back(leaf > 0 ? words++ : -1);
    }
    return plvs_voc_create(device, k, L, n1, n2, (int)parent.size(), parent.data(), word.data(), desc.data(), weight.data(), out);
}

void plvs_voc_destroy(plvs_voc* h)
{
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    delete h;
}

int plvs_voc_size(const plvs_voc* h) { return h ? h->n_words : 0; }

// BowVector (TemplatedVocabulary.h:1158-1206, BowVector.cpp:34-84) from the per-feature words and weights: per word the weights in feature order
// (TF / TF_IDF) or the first one (IDF / BINARY), stopped words (weight 0) dropped, then the norm in ascending word order.  Host arithmetic only.
int plvs_bow_vector(int scoring, int weighting, const uint32_t* word, const double* weight, int n, uint32_t* bow_ids, double* bow_vals, int* n_bow)
{
    if (n < 0 || (n && (!word || !weight)) || !bow_ids || !bow_vals || !n_bow || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3) {
        set_error("bad argument"); return PLVS_EINVAL;
    }
    std::vector<std::pair<uint32_t, int>> items;
    items.reserve(n);
    for (int i = 0; i < n; ++i) if (weight[i] > 0) items.emplace_back(word[i], i);
    std::sort(items.begin(), items.end());
    const bool tf = weighting == 0 || weighting == 1;
    int m = 0;
    for (size_t a = 0; a < items.size();) {
        size_t b;
        double v = weight[items[a].second];
        for (b = a + 1; b < items.size() && items[b].first == items[a].first; ++b) if (tf) v += weight[items[b].second];
        bow_ids[m] = items[a].first; bow_vals[m] = v; ++m;
        a = b;
    }
    const bool must = scoring != 5, l1 = scoring != 1;          // mustNormalize, ScoringObject.h:74-89
    if (tf && m > 0 && !must) { const double nd = (double)m; for (int i = 0; i < m; ++i) bow_vals[i] /= nd; }
    if (must) {
        double norm = 0.0;
        if (l1) { for (int i = 0; i < m; ++i) norm += std::fabs(bow_vals[i]); }
        else { for (int i = 0; i < m; ++i) norm += bow_vals[i] * bow_vals[i]; norm = std::sqrt(norm); }
        if (norm > 0.0) for (int i = 0; i < m; ++i) bow_vals[i] /= norm;
    }
    *n_bow = m;
    return PLVS_OK;
}

int plvs_voc_transform(plvs_voc* h, const uint8_t* desc, int n, int desc_on_device, int levelsup, uint32_t* word, double* weight, uint32_t* node,
                       uint32_t* bow_ids, double* bow_vals, int* n_bow, uint32_t* fv_nodes, int32_t* fv_offsets, int32_t* fv_features, int* n_fv_nodes,
                       plvs_featvec* fv_device)
{
    if (!h || n < 0 || (n && !desc) || n > 65535) { set_error("bad argument"); return PLVS_EINVAL; }
    std::lock_guard<std::mutex> lock(h->mu);
    PLVS_CUDA(cudaSetDevice(h->device));
    if (n_bow) *n_bow = 0;
    if (n_fv_nodes) *n_fv_nodes = 0;
    if (fv_offsets) fv_offsets[0] = 0;
    if (fv_device) *fv_device = plvs_featvec{0, nullptr, nullptr, nullptr};
    if (n == 0) return PLVS_OK;
    cudaStream_t st = h->stream;
    int rc;
    if ((rc = h->d_word.alloc(n)) || (rc = h->d_w.alloc(n)) || (rc = h->d_node.alloc(n)) || (rc = h->d_sorted_feat.alloc(n)) || (rc = h->d_sorted_node.alloc(n)) ||
        (rc = h->d_fv_nodes.alloc(n)) || (rc = h->d_fv_off.alloc(n + 1))) return rc;
    const uint8_t* dd = desc;
    if (!desc_on_device) {
        if ((rc = h->d_desc.alloc((size_t)n * 32))) return rc;
        PLVS_CUDA(cudaMemcpyAsync(h->d_desc.p, desc, (size_t)n * 32, cudaMemcpyHostToDevice, st));
        dd = h->d_desc.p;
    }
    const VocDev V{h->d_child_off.p, h->d_child_id.p, h->d_child_desc.p, h->d_word_id.p, h->d_weight.p, h->L};
    PLVS_CUDA(cudaMemsetAsync(h->d_cnt.p, 0, 16, st));
    k_bow_descend<<<div_up(n, 8), 256, 0, st>>>(V, dd, n, levelsup, h->d_word.p, h->d_w.p, h->d_node.p);
    k_bow_rank<<<div_up(n, 256), 256, 0, st>>>(h->d_node.p, h->d_w.p, n, h->d_sorted_feat.p, h->d_sorted_node.p, h->d_cnt.p);
    k_bow_offsets<<<1, 1024, 0, st>>>(h->d_sorted_node.p, h->d_cnt.p, h->d_fv_nodes.p, h->d_fv_off.p, h->d_cnt.p + 1);
    PLVS_CUDA(cudaGetLastError());
    std::vector<uint32_t> hword(n); std::vector<double> hw(n);
    PLVS_CUDA(cudaMemcpyAsync(hword.data(), h->d_word.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    PLVS_CUDA(cudaMemcpyAsync(hw.data(), h->d_w.p, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    if (node) PLVS_CUDA(cudaMemcpyAsync(node, h->d_node.p, (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    PLVS_CUDA(cudaMemcpyAsync(h->p_cnt.h, h->d_cnt.p, 8, cudaMemcpyDeviceToHost, st));
    PLVS_CUDA(cudaStreamSynchronize(st));
    const int kept = h->p_cnt.h[0], nn = h->p_cnt.h[1];
    if (word) std::memcpy(word, hword.data(), (size_t)n * 4);
    if (weight) std::memcpy(weight, hw.data(), (size_t)n * 8);
    if (n_fv_nodes) *n_fv_nodes = nn;
    if (fv_nodes && nn) PLVS_CUDA(cudaMemcpyAsync(fv_nodes, h->d_fv_nodes.p, (size_t)nn * 4, cudaMemcpyDeviceToHost, st));
    if (fv_offsets) PLVS_CUDA(cudaMemcpyAsync(fv_offsets, h->d_fv_off.p, (size_t)(nn + 1) * 4, cudaMemcpyDeviceToHost, st));
    if (fv_features && kept) PLVS_CUDA(cudaMemcpyAsync(fv_features, h->d_sorted_feat.p, (size_t)kept * 4, cudaMemcpyDeviceToHost, st));
    if (fv_device) *fv_device = plvs_featvec{nn, h->d_fv_nodes.p, h->d_fv_off.p, h->d_sorted_feat.p};      // valid until the next transform on this handle
    if (bow_ids || bow_vals || n_bow) {
        int nb = 0;
        std::vector<uint32_t> ids(n); std::vector<double> vals(n);
        plvs_bow_vector(h->scoring, h->weighting, hword.data(), hw.data(), n, ids.data(), vals.data(), &nb);
        if (n_bow) *n_bow = nb;
        if (bow_ids) std::memcpy(bow_ids, ids.data(), (size_t)nb * 4);
        if (bow_vals) std::memcpy(bow_vals, vals.data(), (size_t)nb * 8);
    }
    PLVS_CUDA(cudaStreamSynchronize(st));
    return PLVS_OK;
}

}  // extern "C"
