// TEST INFRASTRUCTURE ONLY -- CPU oracle for the matching rows of SURVEY.md §8a (a11-a16).
//
// Restates, on flat arrays, PLVS2::ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2198-2225),
// Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (src/Frame.cc:716-746,1305-1316,
// 1231-1303), ORBmatcher::SearchByProjection(Frame&, vector<MapPointPtr>&,...) (src/ORBmatcher.cc:71-157,
// RGB-D / Nleft==-1 branch), SearchByProjection(Cur, Last,...) (:1774-1993), SearchForTriangulation
// (:999-1242, single-camera branch), ComputeThreeMaxima (:2123-2164) and the line test of
// Pinhole::epipolarConstrain (src/CameraModels/Pinhole.cpp:133-147).  Sequential, claim-mutating,
// exactly in the reference's iteration order.  The structs below mirror include/plvs_b200.h so the
// tests hand the same buffers to the oracle and to the CUDA path.
//
// Parity status: the three searches and DescriptorDistance are PINNED -- tests/test_oracle_vs_reference_match.py checks them
// bit-exactly against the reference's own src/ORBmatcher.cc, compiled into oracle/_ref/libmatch_ref.so by oracle/ref_build.py
// (data-model stand-ins in oracle/plvs_standin), and tests/golden/match_ref.npz holds outputs recorded from it.
// ComputeStereoMatches (below, src/Frame.cc:1780-1983) is PINNED as well: Frame.cc as a whole cannot be compiled, so
// oracle/ref_build.py slices exactly that function definition out of it at build time (into the git-ignored oracle/_ref/gen/) and
// compiles it with the reference's own ORBextractor.cc; tests/test_oracle_vs_reference_stereo.py compares bit-exactly.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

const int TH_HIGH = 100, TH_LOW = 50, HISTO_LENGTH = 12;
const int GRID_COLS = 64, GRID_ROWS = 48;

struct KeyPoint { float x, y, size, angle, response; int32_t octave, class_id; };

struct FrameView {
    int32_t n;
    const KeyPoint* keys;
    const uint8_t* desc;
    const float* uright;
    float min_x, min_y, max_x, max_y;
    float grid_inv_w, grid_inv_h;
    float scale_factors[16];
    float level_sigma2[16];
    int32_t nlevels;
    float bf;
    int32_t on_device;
    uint64_t cache_key;
};

struct MpQuery { float proj_x, proj_y, proj_xr, track_depth, view_cos; int32_t level; uint32_t flags; uint8_t desc[32]; };
struct LastQuery { float u, v, invz; int32_t last_octave; float angle; uint32_t flags; uint8_t desc[32]; };
struct FeatVec { int32_t n_nodes; const uint32_t* node_ids; const int32_t* offsets; const int32_t* features; };

int hamming(const uint8_t* a, const uint8_t* b)
{
    uint64_t pa[4], pb[4];
    std::memcpy(pa, a, 32); std::memcpy(pb, b, 32);
    int d = 0;
    for (int i = 0; i < 4; ++i) d += __builtin_popcountll(pa[i] ^ pb[i]);
    return d;
}

struct Grid {
    std::vector<int> cell[GRID_COLS][GRID_ROWS];
    const FrameView* f;
    explicit Grid(const FrameView* fv) : f(fv)
    {
        for (int i = 0; i < f->n; ++i) {
            const KeyPoint& kp = f->keys[i];
            int px = (int)std::round((kp.x - f->min_x) * f->grid_inv_w);
            int py = (int)std::round((kp.y - f->min_y) * f->grid_inv_h);
            if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
            cell[px][py].push_back(i);
        }
    }
    void in_area(float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const
    {
        out.clear();
        const int minCX = std::max(0, (int)std::floor((x - f->min_x - r) * f->grid_inv_w));
        if (minCX >= GRID_COLS) return;
        const int maxCX = std::min(GRID_COLS - 1, (int)std::ceil((x - f->min_x + r) * f->grid_inv_w));
        if (maxCX < 0) return;
        const int minCY = std::max(0, (int)std::floor((y - f->min_y - r) * f->grid_inv_h));
        if (minCY >= GRID_ROWS) return;
        const int maxCY = std::min(GRID_ROWS - 1, (int)std::ceil((y - f->min_y + r) * f->grid_inv_h));
        if (maxCY < 0) return;
        const bool check = (minLevel > 0) || (maxLevel >= 0);
        for (int ix = minCX; ix <= maxCX; ++ix)
            for (int iy = minCY; iy <= maxCY; ++iy)
                for (int id : cell[ix][iy]) {
                    const KeyPoint& kp = f->keys[id];
                    if (check) {
                        if (kp.octave < minLevel) continue;
                        if (kp.octave > maxLevel) continue;      // NB: applied even when maxLevel == -1 (src/Frame.cc:1283-1286)
                    }
                    const float dx = kp.x - x, dy = kp.y - y;
                    if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(id);
                }
    }
};

void three_maxima(const std::vector<int>* histo, int L, int& i1, int& i2, int& i3)
{
    int m1 = 0, m2 = 0, m3 = 0;
    for (int i = 0; i < L; ++i) {
        const int s = (int)histo[i].size();
        if (s > m1) { m3 = m2; m2 = m1; m1 = s; i3 = i2; i2 = i1; i1 = i; }
        else if (s > m2) { m3 = m2; m2 = s; i3 = i2; i2 = i; }
        else if (s > m3) { m3 = s; i3 = i; }
    }
    if (m2 < 0.1f * (float)m1) { i2 = -1; i3 = -1; }
    else if (m3 < 0.1f * (float)m1) { i3 = -1; }
}

int rot_bin(float a1, float a2)
{
    const float factor = HISTO_LENGTH / 360.0f;
    float rot = a1 - a2;
    if (rot < 0.0) rot += 360.0f;
    int bin = (int)std::round(rot * factor);
    if (bin == HISTO_LENGTH) bin = 0;
    return bin;
}

}  // namespace

extern "C" {

int orc_hamming256(const uint8_t* a, const uint8_t* b) { return hamming(a, b); }

// holder[i]: -1 free, -2 pre-claimed by a map point with Observations()>0, >=0 index of the query
// that wrote it during this call.  assign[i] = holder[i] >= 0 ? holder[i] : -1.
int orc_search_by_projection_map(const FrameView* F, const MpQuery* q, int nq, float th, float nn_ratio,
                                 int far_points, float th_far, const uint8_t* claimed_in, int32_t* assign)
{
    Grid grid(F);
    std::vector<int> holder(F->n, -1);
    if (claimed_in) for (int i = 0; i < F->n; ++i) if (claimed_in[i]) holder[i] = -2;
    int nmatches = 0;
    const bool bFactor = th != 1.0;
    std::vector<int> cand;
    for (int iq = 0; iq < nq; ++iq) {
        const MpQuery& m = q[iq];
        if (far_points && m.track_depth > th_far) continue;
        const int lvl = m.level;
        float r = m.view_cos > 0.998 ? 2.5f : 4.0f;
        if (bFactor) r *= th;
        grid.in_area(m.proj_x, m.proj_y, r * F->scale_factors[lvl], lvl - 1, lvl, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : cand) {
            if (holder[idx] == -2) continue;
            if (holder[idx] >= 0 && (q[holder[idx]].flags & 1u)) continue;      // Observations()>0
            if (F->uright && F->uright[idx] > 0) {
                const float er = std::fabs(m.proj_xr - F->uright[idx]);
                if (er > r * F->scale_factors[lvl]) continue;
            }
            const int dist = hamming(m.desc, F->desc + (size_t)idx * 32);
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel;
                bestLevel = F->keys[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = F->keys[idx].octave; bestDist2 = dist;
            }
        }
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nn_ratio * bestDist2) continue;
            if (bestLevel != bestLevel2 || bestDist <= nn_ratio * bestDist2) { holder[bestIdx] = iq; ++nmatches; }
        }
    }
    for (int i = 0; i < F->n; ++i) assign[i] = holder[i] >= 0 ? holder[i] : -1;
    return nmatches;
}

int orc_search_by_projection_last(const FrameView* C, const LastQuery* q, int nq, float th, int forward, int backward,
                                  int check_ori, const uint8_t* claimed_in, int32_t* assign)
{
    Grid grid(C);
    std::vector<int> holder(C->n, -1);
    if (claimed_in) for (int i = 0; i < C->n; ++i) if (claimed_in[i]) holder[i] = -2;
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0;
    std::vector<int> cand;
    for (int iq = 0; iq < nq; ++iq) {
        const LastQuery& m = q[iq];
        if (m.invz < 0) continue;
        if (m.u < C->min_x || m.u > C->max_x) continue;
        if (m.v < C->min_y || m.v > C->max_y) continue;
        const int oct = m.last_octave;
        const float radius = th * C->scale_factors[oct];
        if (forward) grid.in_area(m.u, m.v, radius, oct, -1, cand);
        else if (backward) grid.in_area(m.u, m.v, radius, 0, oct, cand);
        else grid.in_area(m.u, m.v, radius, oct - 1, oct + 1, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if (holder[i2] == -2) continue;
            if (holder[i2] >= 0 && (q[holder[i2]].flags & 1u)) continue;
            if (C->uright && C->uright[i2] > 0) {
                const float ur = m.u - C->bf * m.invz;
                const float er = std::fabs(ur - C->uright[i2]);
                if (er > radius) continue;
            }
            const int dist = hamming(m.desc, C->desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= TH_HIGH) {
            holder[bestIdx2] = iq;
            ++nmatches;
            if (check_ori) rotHist[rot_bin(m.angle, C->keys[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (check_ori) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != i1 && i != i2 && i != i3)
                for (int idx : rotHist[i]) { holder[idx] = -1; --nmatches; }
    }
    for (int i = 0; i < C->n; ++i) assign[i] = holder[i] >= 0 ? holder[i] : -1;
    return nmatches;
}

int orc_search_for_triangulation(const FrameView* K1, const FrameView* K2, const FeatVec* fv1, const FeatVec* fv2,
                                 const uint8_t* has_mp1, const uint8_t* has_mp2, const float* F12, const float* ep,
                                 int only_stereo, int coarse, int check_ori, int32_t* match12)
{
    int nmatches = 0;
    for (int i = 0; i < K1->n; ++i) match12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    int a = 0, b = 0;
    while (a < fv1->n_nodes && b < fv2->n_nodes) {
        if (fv1->node_ids[a] == fv2->node_ids[b]) {
            for (int i1 = fv1->offsets[a]; i1 < fv1->offsets[a + 1]; ++i1) {
                const int idx1 = fv1->features[i1];
                if (has_mp1[idx1]) continue;
                const bool stereo1 = K1->uright && K1->uright[idx1] >= 0;
                if (only_stereo && !stereo1) continue;
                const KeyPoint& kp1 = K1->keys[idx1];
                int bestDist = TH_LOW, bestIdx2 = -1;
                for (int i2 = fv2->offsets[b]; i2 < fv2->offsets[b + 1]; ++i2) {
                    const int idx2 = fv2->features[i2];
                    if (has_mp2[idx2]) continue;
                    const bool stereo2 = K2->uright && K2->uright[idx2] >= 0;
                    if (only_stereo && !stereo2) continue;
                    const int dist = hamming(K1->desc + (size_t)idx1 * 32, K2->desc + (size_t)idx2 * 32);
                    if (dist > TH_LOW || dist > bestDist) continue;
                    const KeyPoint& kp2 = K2->keys[idx2];
                    if (!stereo1 && !stereo2) {
                        const float dx = ep[0] - kp2.x, dy = ep[1] - kp2.y;
                        if (dx * dx + dy * dy < 100 * K2->scale_factors[kp2.octave]) continue;
                    }
                    bool ok = coarse != 0;
                    if (!ok) {   // Pinhole::epipolarConstrain line test with the caller's F12
                        const float la = kp1.x * F12[0] + kp1.y * F12[3] + F12[6];
                        const float lb = kp1.x * F12[1] + kp1.y * F12[4] + F12[7];
                        const float lc = kp1.x * F12[2] + kp1.y * F12[5] + F12[8];
                        const float num = la * kp2.x + lb * kp2.y + lc;
                        const float den = la * la + lb * lb;
                        if (den != 0) {
                            const float dsqr = num * num / den;
                            ok = dsqr < 3.84 * K2->level_sigma2[kp2.octave];
                        }
                    }
                    if (ok) { bestIdx2 = idx2; bestDist = dist; }
                }
                if (bestIdx2 >= 0) {
                    match12[idx1] = bestIdx2;
                    ++nmatches;
                    if (check_ori) rotHist[rot_bin(kp1.angle, K2->keys[bestIdx2].angle)].push_back(idx1);
                }
            }
            ++a; ++b;
        } else if (fv1->node_ids[a] < fv2->node_ids[b]) {
            while (a < fv1->n_nodes && fv1->node_ids[a] < fv2->node_ids[b]) ++a;     // lower_bound
        } else {
            while (b < fv2->n_nodes && fv2->node_ids[b] < fv1->node_ids[a]) ++b;
        }
    }
    if (check_ori) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int idx : rotHist[i]) { match12[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// Frame::isInFrustum(MapPointPtr&, viewingCosLimit) (src/Frame.cc:955-1017, Nleft == -1) with MapPoint::PredictScale(dist, Frame*)
// (src/MapPoint.cc:598-613) and Pinhole::project (src/CameraModels/Pinhole.cpp:61-67): the step that turns the local map into the
// queries of SearchByProjection(F, vpMapPoints) (Tracking::SearchLocalPoints).  Eigen's 3-term order e0 + (e1 + e2); `log` / `ceil` are the
// float overloads (`using namespace std` reaches MapPoint.cc).  min_dist / max_dist are mfMinDistance / mfMaxDistance.
struct MapPointIn { float xw[3], normal[3], min_dist, max_dist; uint32_t flags; uint8_t desc[32]; };
struct FrustumIn { float Rcw[9], tcw[3], Ow[3], fx, fy, cx, cy, bf, viewing_cos_limit, scale_factor; int32_t nlevels; float min_x, min_y, max_x, max_y; };

static inline int predict_scale(float max_distance, float current_dist, float log_scale_factor, int nlevels)
{
    const float ratio = max_distance / current_dist;
    int nScale = (int)std::ceil(std::log(ratio) / log_scale_factor);        // float log, float division, float ceil
    if (nScale < 0) nScale = 0; else if (nScale >= nlevels) nScale = nlevels - 1;
    return nScale;
}

int orc_in_frustum(const FrustumIn* fr, const MapPointIn* pts, int n, MpQuery* q, uint8_t* in_view)
{
    const float L = std::log(fr->scale_factor);
    const float* R = fr->Rcw;
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        const MapPointIn& p = pts[i];
        MpQuery& o = q[i];
        std::memset(&o, 0, sizeof(o));
        o.proj_x = -1; o.proj_y = -1; o.flags = p.flags; std::memcpy(o.desc, p.desc, 32);
        in_view[i] = 0;
        const float X = p.xw[0], Y = p.xw[1], Z = p.xw[2];
        const float pcx = (R[0] * X + (R[1] * Y + R[2] * Z)) + fr->tcw[0];
        const float pcy = (R[3] * X + (R[4] * Y + R[5] * Z)) + fr->tcw[1];
        const float pcz = (R[6] * X + (R[7] * Y + R[8] * Z)) + fr->tcw[2];
        const float pc_dist = std::sqrt(pcx * pcx + (pcy * pcy + pcz * pcz));
        const float invz = 1.0f / pcz;
        if (pcz < 0.0f) continue;
        const float u = fr->fx * pcx / pcz + fr->cx, v = fr->fy * pcy / pcz + fr->cy;
        if (u < fr->min_x || u > fr->max_x) continue;
        if (v < fr->min_y || v > fr->max_y) continue;
        o.proj_x = u; o.proj_y = v;
        const float maxDistance = 1.2f * p.max_dist, minDistance = 0.8f * p.min_dist;
        const float pox = X - fr->Ow[0], poy = Y - fr->Ow[1], poz = Z - fr->Ow[2];
        const float dist = std::sqrt(pox * pox + (poy * poy + poz * poz));
        if (dist < minDistance || dist > maxDistance) continue;
        const float viewCos = (pox * p.normal[0] + (poy * p.normal[1] + poz * p.normal[2])) / dist;
        if (viewCos < fr->viewing_cos_limit) continue;
        o.level = predict_scale(p.max_dist, dist, L, fr->nlevels);
        o.proj_xr = u - fr->bf * invz;
        o.track_depth = pc_dist;
        o.view_cos = viewCos;
        in_view[i] = 1; ++cnt;
    }
    return cnt;
}

// The device evaluates PredictScale without a logarithm: with g(ratio) = logf(ratio) / L non-decreasing, level = #{k in [0, nlevels-2] : ratio >= T_k}
// where T_k = the smallest float with ceil(g) > k, found on the HOST with the host's libm (bisection over the ordered positive floats).
void orc_scale_thresholds(float scale_factor, int nlevels, float* T /*nlevels-1*/)
{
    const float L = std::log(scale_factor);
    for (int k = 0; k + 1 < nlevels; ++k) {
        uint32_t lo = 0x00800000u, hi = 0x7f7fffffu;            // positive normal floats; predicate false at lo, true at hi
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            float r; std::memcpy(&r, &mid, 4);
            if (std::ceil(std::log(r) / L) > (float)k) hi = mid; else lo = mid;
        }
        std::memcpy(&T[k], &hi, 4);
    }
}

// exhaustive check of that equivalence for every float in [lo, hi]: returns the number of ratios where the threshold count differs from
// the direct formula (0 unless logf is non-monotonic somewhere in the range)
long long orc_scale_threshold_mismatches(float scale_factor, int nlevels, float lo, float hi)
{
    std::vector<float> T(nlevels > 1 ? nlevels - 1 : 1);
    orc_scale_thresholds(scale_factor, nlevels, T.data());
    const float L = std::log(scale_factor);
    uint32_t a, b; std::memcpy(&a, &lo, 4); std::memcpy(&b, &hi, 4);
    long long bad = 0;
    for (uint32_t u = a; u <= b; ++u) {
        float r; std::memcpy(&r, &u, 4);
        int direct = (int)std::ceil(std::log(r) / L);
        if (direct < 0) direct = 0; else if (direct >= nlevels) direct = nlevels - 1;
        int cnt = 0;
        for (int k = 0; k + 1 < nlevels; ++k) cnt += r >= T[k];
        bad += cnt != direct;
    }
    return bad;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, KeyFramePtr& pKF, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1996-2122):
// queries = the keyframe's map points that passed the caller-side gates, pre-projected (LastQuery records, last_octave = predicted
// level).  Level window [l-1, l+1], no right-coordinate gate, ANY non-null mvpMapPoints entry blocks, accept if best <= ORBdist.
int orc_search_by_projection_reloc(const FrameView* C, const LastQuery* q, int nq, float th, int orb_dist, int check_ori,
                                   const uint8_t* claimed_in, int32_t* assign)
{
    Grid grid(C);
    std::vector<int> holder(C->n, -1);
    if (claimed_in) for (int i = 0; i < C->n; ++i) if (claimed_in[i]) holder[i] = -2;
    std::vector<int> rotHist[HISTO_LENGTH];
    int nmatches = 0;
    std::vector<int> cand;
    for (int iq = 0; iq < nq; ++iq) {
        const LastQuery& m = q[iq];
        if (m.u < C->min_x || m.u > C->max_x) continue;
        if (m.v < C->min_y || m.v > C->max_y) continue;
        const int lvl = m.last_octave;
        const float radius = th * C->scale_factors[lvl];
        grid.in_area(m.u, m.v, radius, lvl - 1, lvl + 1, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if (holder[i2] != -1) continue;
            const int dist = hamming(m.desc, C->desc + (size_t)i2 * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx2 = i2; }
        }
        if (bestDist <= orb_dist) {
            holder[bestIdx2] = iq;
            ++nmatches;
            if (check_ori) rotHist[rot_bin(m.angle, C->keys[bestIdx2].angle)].push_back(bestIdx2);
        }
    }
    if (check_ori) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i)
            if (i != i1 && i != i2 && i != i3)
                for (int idx : rotHist[i]) { holder[idx] = -1; --nmatches; }
    }
    for (int i = 0; i < C->n; ++i) assign[i] = holder[i] >= 0 ? holder[i] : -1;
    return nmatches;
}

// ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th, ratioHamming) (src/ORBmatcher.cc:509-615; :617-730 is the same
// search): window of KeyFrame::GetFeaturesInArea (no level filter), level gate [l-1, l], any non-null vpMatched entry blocks,
// accept if best <= TH_LOW * ratioHamming.
int orc_search_by_projection_sim3(const FrameView* K, const LastQuery* q, int nq, float th, float ratio_hamming, const uint8_t* matched_in, int32_t* assign)
{
    Grid grid(K);
    std::vector<int> holder(K->n, -1);
    if (matched_in) for (int i = 0; i < K->n; ++i) if (matched_in[i]) holder[i] = -2;
    int nmatches = 0;
    std::vector<int> cand;
    for (int iq = 0; iq < nq; ++iq) {
        const LastQuery& m = q[iq];
        const int lvl = m.last_octave;
        const float radius = th * K->scale_factors[lvl];
        grid.in_area(m.u, m.v, radius, -1, -1, cand);
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            if (holder[idx] != -1) continue;
            const int kpLevel = K->keys[idx].octave;
            if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
            const int dist = hamming(m.desc, K->desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        if (bestDist <= TH_LOW * ratio_hamming) { holder[bestIdx] = iq; ++nmatches; }
    }
    for (int i = 0; i < K->n; ++i) assign[i] = holder[i] >= 0 ? holder[i] : -1;
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFramePtr& pKF, Frame& F, vector<MapPointPtr>& vpMapPointMatches) (src/ORBmatcher.cc:300-506),
// Nleft == -1 branch: merge-walk of the two FeatureVectors; inside a shared node the keyframe features that carry a (good)
// map point are visited in order, each takes the best frame feature of the node that is still unmatched (best/second with
// the multiset semantics of the if / else-if pair, strict `<` so the first index wins ties), claims it if
// best <= TH_LOW and best < ratio * second; rotation histogram over the claimed frame features at the end.
// has_mp[i] != 0 <=> vpMapPointsKF[i] && !isBad().  match_f[iF] = keyframe feature whose map point went to frame feature iF.
int orc_search_by_bow(const FrameView* K, const FrameView* F, const FeatVec* fvK, const FeatVec* fvF, const uint8_t* has_mp,
                      float nn_ratio, int check_ori, int32_t* match_f)
{
    int nmatches = 0;
    for (int i = 0; i < F->n; ++i) match_f[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    int a = 0, b = 0;
    while (a < fvK->n_nodes && b < fvF->n_nodes) {
        if (fvK->node_ids[a] == fvF->node_ids[b]) {
            for (int iK = fvK->offsets[a]; iK < fvK->offsets[a + 1]; ++iK) {
                const int idxK = fvK->features[iK];
                if (!has_mp[idxK]) continue;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int iF = fvF->offsets[b]; iF < fvF->offsets[b + 1]; ++iF) {
                    const int idxF = fvF->features[iF];
                    if (match_f[idxF] >= 0) continue;
                    const int dist = hamming(K->desc + (size_t)idxK * 32, F->desc + (size_t)idxF * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = idxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW) {
                    if ((float)bestDist1 < nn_ratio * (float)bestDist2) {
                        match_f[bestIdxF] = idxK;
                        if (check_ori) rotHist[rot_bin(K->keys[idxK].angle, F->keys[bestIdxF].angle)].push_back(bestIdxF);
                        ++nmatches;
                    }
                }
            }
            ++a; ++b;
        } else if (fvK->node_ids[a] < fvF->node_ids[b]) {
            while (a < fvK->n_nodes && fvK->node_ids[a] < fvF->node_ids[b]) ++a;
        } else {
            while (b < fvF->n_nodes && fvF->node_ids[b] < fvK->node_ids[a]) ++b;
        }
    }
    if (check_ori) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int idx : rotHist[i]) { match_f[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// ORBmatcher::SearchByBoW(KeyFramePtr& pKF1, KeyFramePtr& pKF2, vector<MapPointPtr>& vpMatches12) (src/ORBmatcher.cc:853-997):
// like the frame overload, but set-2 features must carry a good map point too, the claim lives in vbMatched2, the acceptance is
// `bestDist1 < TH_LOW` (strict) and the result is indexed by the KF1 feature.
int orc_search_by_bow_kf(const FrameView* K1, const FrameView* K2, const FeatVec* fv1, const FeatVec* fv2, const uint8_t* has1, const uint8_t* has2,
                         float nn_ratio, int check_ori, int32_t* match12)
{
    int nmatches = 0;
    for (int i = 0; i < K1->n; ++i) match12[i] = -1;
    std::vector<char> matched2(K2->n, 0);
    std::vector<int> rotHist[HISTO_LENGTH];
    int a = 0, b = 0;
    while (a < fv1->n_nodes && b < fv2->n_nodes) {
        if (fv1->node_ids[a] == fv2->node_ids[b]) {
            for (int i1 = fv1->offsets[a]; i1 < fv1->offsets[a + 1]; ++i1) {
                const int idx1 = fv1->features[i1];
                if (!has1[idx1]) continue;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int i2 = fv2->offsets[b]; i2 < fv2->offsets[b + 1]; ++i2) {
                    const int idx2 = fv2->features[i2];
                    if (matched2[idx2] || !has2[idx2]) continue;
                    const int dist = hamming(K1->desc + (size_t)idx1 * 32, K2->desc + (size_t)idx2 * 32);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW) {
                    if ((float)bestDist1 < nn_ratio * (float)bestDist2) {
                        match12[idx1] = bestIdx2; matched2[bestIdx2] = 1;
                        if (check_ori) rotHist[rot_bin(K1->keys[idx1].angle, K2->keys[bestIdx2].angle)].push_back(idx1);
                        ++nmatches;
                    }
                }
            }
            ++a; ++b;
        } else if (fv1->node_ids[a] < fv2->node_ids[b]) {
            while (a < fv1->n_nodes && fv1->node_ids[a] < fv2->node_ids[b]) ++a;
        } else {
            while (b < fv2->n_nodes && fv2->node_ids[b] < fv1->node_ids[a]) ++b;
        }
    }
    if (check_ori) {
        int i1 = -1, i2 = -1, i3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, i1, i2, i3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == i1 || i == i2 || i == i3) continue;
            for (int idx : rotHist[i]) { match12[idx] = -1; --nmatches; }
        }
    }
    return nmatches;
}

// ORBmatcher::Fuse(pKF, vpMapPoints, th, bRight=false) (src/ORBmatcher.cc:1244-1435), the search part (:1340-1406): window from
// KeyFrame::GetFeaturesInArea (src/KeyFrame.cc:1179-1229, no level filter), level gate [l-1, l], chi-square gate on the
// reprojection error (7.8 with a right coordinate, 5.99 without), best distance, first wins ties.  The gates before the search
// and the Replace / AddObservation bookkeeping after it are the caller's (pointer graph); the search does not read map points,
// so the queries are independent of each other.
struct FuseQuery { float u, v, ur; int32_t level; uint8_t desc[32]; };

static int fuse_search(const FrameView* K, const float* inv_level_sigma2, const FuseQuery* q, int nq, float th, int32_t* best_idx, int32_t* best_dist, bool chi2);
int orc_fuse(const FrameView* K, const float* inv_level_sigma2, const FuseQuery* q, int nq, float th, int32_t* best_idx, int32_t* best_dist)
{
    return fuse_search(K, inv_level_sigma2, q, nq, th, best_idx, best_dist, true);
}
// Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:1437-1553): the same search without the chi-square gate
int orc_fuse_sim3(const FrameView* K, const FuseQuery* q, int nq, float th, int32_t* best_idx, int32_t* best_dist)
{
    return fuse_search(K, nullptr, q, nq, th, best_idx, best_dist, false);
}
static int fuse_search(const FrameView* K, const float* inv_level_sigma2, const FuseQuery* q, int nq, float th, int32_t* best_idx, int32_t* best_dist, bool chi2)
{
    Grid grid(K);
    std::vector<int> cand;
    int nfused = 0;
    for (int iq = 0; iq < nq; ++iq) {
        const FuseQuery& m = q[iq];
        const int lvl = m.level;
        const float radius = th * K->scale_factors[lvl];
        grid.in_area(m.u, m.v, radius, -1, -1, cand);            // minLevel <= 0 and maxLevel < 0: no level filter
        int bestDist = 256, bestIdx = -1;
        for (int idx : cand) {
            const KeyPoint& kp = K->keys[idx];
            const int kpLevel = kp.octave;
            if (kpLevel < lvl - 1 || kpLevel > lvl) continue;
            if (!chi2) {
            } else if (K->uright && K->uright[idx] >= 0) {
                const float ex = m.u - kp.x, ey = m.v - kp.y, er = m.ur - K->uright[idx];
                const float e2 = ex * ex + ey * ey + er * er;
                if (e2 * inv_level_sigma2[kpLevel] > 7.8) continue;
            } else {
                const float ex = m.u - kp.x, ey = m.v - kp.y;
                const float e2 = ex * ex + ey * ey;
                if (e2 * inv_level_sigma2[kpLevel] > 5.99) continue;
            }
            const int dist = hamming(m.desc, K->desc + (size_t)idx * 32);
            if (dist < bestDist) { bestDist = dist; bestIdx = idx; }
        }
        best_idx[iq] = bestIdx; best_dist[iq] = bestDist;
        if (bestDist <= TH_LOW) ++nfused;
    }
    return nfused;
}

// ---------------------------------------------------------------------------------------------
// ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:732-852): monocular start-up.  Level-0 keypoints of F1 search a square
// window of F2 around their previously matched position; a candidate is skipped when an earlier accepted match on it was at least
// as good (vMatchedDistance), the winner may steal a feature from an earlier query (vnMatches21), the rotation histogram keeps every
// accepted i1 (also the ones stolen later).  prev_matched (cv::Point2f per F1 keypoint) is updated for the final matches.
// PINNED: tests/test_oracle_vs_reference_match.py runs the reference's own function on the same inputs.
// ---------------------------------------------------------------------------------------------
int orc_search_for_initialization(const FrameView* F1, const FrameView* F2, float* prev_matched, int window_size, float nn_ratio, int check_ori,
                                  int32_t* matches12)
{
    int nmatches = 0;
    for (int i = 0; i < F1->n; ++i) matches12[i] = -1;
    std::vector<int> rotHist[HISTO_LENGTH];
    std::vector<int> matched_distance(F2->n, INT32_MAX), matches21(F2->n, -1);
    Grid grid(F2);
    std::vector<int> cand;
    for (int i1 = 0; i1 < F1->n; ++i1) {
        const KeyPoint& kp1 = F1->keys[i1];
        const int level1 = kp1.octave;
        if (level1 > 0) continue;
        grid.in_area(prev_matched[2 * i1], prev_matched[2 * i1 + 1], (float)window_size, level1, level1, cand);
        if (cand.empty()) continue;
        const uint8_t* d1 = F1->desc + (size_t)i1 * 32;
        int bestDist = INT32_MAX, bestDist2 = INT32_MAX, bestIdx2 = -1;
        for (int i2 : cand) {
            const int dist = hamming(d1, F2->desc + (size_t)i2 * 32);
            if (matched_distance[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= TH_LOW && bestDist < (float)bestDist2 * nn_ratio) {
            if (matches21[bestIdx2] >= 0) { matches12[matches21[bestIdx2]] = -1; nmatches--; }
            matches12[i1] = bestIdx2; matches21[bestIdx2] = i1; matched_distance[bestIdx2] = bestDist;
            nmatches++;
            if (check_ori) rotHist[rot_bin(F1->keys[i1].angle, F2->keys[bestIdx2].angle)].push_back(i1);
        }
    }
    if (check_ori) {
        int ind1 = -1, ind2 = -1, ind3 = -1;
        three_maxima(rotHist, HISTO_LENGTH, ind1, ind2, ind3);
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int idx1 : rotHist[i]) if (matches12[idx1] >= 0) { matches12[idx1] = -1; nmatches--; }
        }
    }
    for (int i1 = 0; i1 < F1->n; ++i1)
        if (matches12[i1] >= 0) { prev_matched[2 * i1] = F2->keys[matches12[i1]].x; prev_matched[2 * i1 + 1] = F2->keys[matches12[i1]].y; }
    return nmatches;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// a17: Frame::ComputeStereoMatches (src/Frame.cc:1780-1983) -- rectified stereo row matching with
// 11x11 SAD refinement on the UNBLURRED pyramid level of the left keypoint, parabola sub-pixel fit,
// median-based outlier cut.  The reference's pyramid levels sit inside a 19-px BORDER_REFLECT_101
// frame (src/ORBextractor.cc:1496-1501); the SAD windows may reach up to 10 columns into it on the
// left, so out-of-level columns are read through the same reflection.
// ---------------------------------------------------------------------------------------------
extern "C" {

struct OrcLevel { const uint8_t* data; int w, h, stride; };

static inline int refl101(int p, int n) { if (p < 0) p = -p; if (p >= n) p = 2 * n - 2 - p; return p; }

int orc_compute_stereo_matches(int N, const KeyPoint* keysL, const uint8_t* descL, int Nr, const KeyPoint* keysR, const uint8_t* descR,
                               const OrcLevel* pyrL, const OrcLevel* pyrR, const float* scale, const float* inv_scale,
                               float mb, float mbf, float* uRight, float* depthOut)
{
    for (int i = 0; i < N; ++i) { uRight[i] = -1.0f; depthOut[i] = -1.0f; }
    const int thOrbDist = (TH_HIGH + TH_LOW) / 2;
    const int nRows = pyrL[0].h;
    std::vector<std::vector<int>> rows(nRows);
    for (int iR = 0; iR < Nr; ++iR) {
        const float kpY = keysR[iR].y;
        const float r = 2.0f * scale[keysR[iR].octave];
        const int maxr = (int)std::ceil(kpY + r), minr = (int)std::floor(kpY - r);
        for (int yi = minr; yi <= maxr; ++yi) if (yi >= 0 && yi < nRows) rows[yi].push_back(iR);   // the reference indexes unchecked; keypoints are >=19 px inside
    }
    const float minZ = mb, minD = 0, maxD = mbf / minZ;
    std::vector<std::pair<int, int>> vDistIdx;
    for (int iL = 0; iL < N; ++iL) {
        const KeyPoint& kpL = keysL[iL];
        const int levelL = kpL.octave;
        const float vL = kpL.y, uL = kpL.x;
        const std::vector<int>& cands = rows[(int)vL];
        if (cands.empty()) continue;
        const float minU = uL - maxD, maxU = uL - minD;
        if (maxU < 0) continue;
        int bestDist = TH_HIGH; int bestIdxR = 0;
        for (int iR : cands) {
            const KeyPoint& kpR = keysR[iR];
            if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
            const float uR = kpR.x;
            if (uR >= minU && uR <= maxU) {
                const int dist = hamming(descL + (size_t)iL * 32, descR + (size_t)iR * 32);
                if (dist < bestDist) { bestDist = dist; bestIdxR = iR; }
            }
        }
        if (bestDist < thOrbDist) {
            const float uR0 = keysR[bestIdxR].x;
            const float sf = inv_scale[kpL.octave];
            const float scaleduL = std::round(kpL.x * sf), scaledvL = std::round(kpL.y * sf), scaleduR0 = std::round(uR0 * sf);
            const int w = 5, L = 5;
            const OrcLevel& IL = pyrL[kpL.octave]; const OrcLevel& IR = pyrR[kpL.octave];
            int bestD = 2147483647, bestincR = 0;
            float vDists[11];
            const float iniu = scaleduR0 + L - w, endu = scaleduR0 + L + w + 1;
            if (iniu < 0 || endu >= IR.w) continue;
            for (int incR = -L; incR <= L; ++incR) {
                double acc = 0;
                for (int dy = -w; dy <= w; ++dy)
                    for (int dx = -w; dx <= w; ++dx) {
                        const int yl = refl101((int)scaledvL + dy, IL.h), xl = refl101((int)scaleduL + dx, IL.w);
                        const int yr = refl101((int)scaledvL + dy, IR.h), xr = refl101((int)scaleduR0 + incR + dx, IR.w);
                        acc += std::abs((int)IL.data[(size_t)yl * IL.stride + xl] - (int)IR.data[(size_t)yr * IR.stride + xr]);
                    }
                const float dist = (float)acc;            // cv::norm returns double, stored into float
                if (dist < bestD) { bestD = (int)dist; bestincR = incR; }
                vDists[L + incR] = dist;
            }
            if (bestincR == -L || bestincR == L) continue;
            const float dist1 = vDists[L + bestincR - 1], dist2 = vDists[L + bestincR], dist3 = vDists[L + bestincR + 1];
            const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
            if (deltaR < -1 || deltaR > 1) continue;
            float bestuR = scale[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
            float disparity = (uL - bestuR);
            if (disparity >= minD && disparity < maxD) {
                if (disparity <= 0) { disparity = 0.01; bestuR = uL - 0.01; }
                depthOut[iL] = mbf / disparity;
                uRight[iL] = bestuR;
                vDistIdx.push_back(std::pair<int, int>(bestD, iL));
            }
        }
    }
    if (vDistIdx.empty()) return 0;          // (the reference would index an empty vector here)
    std::sort(vDistIdx.begin(), vDistIdx.end());
    const float median = (float)vDistIdx[vDistIdx.size() / 2].first;
    const float thDist = 1.5f * 1.4f * median;
    int kept = (int)vDistIdx.size();
    for (int i = (int)vDistIdx.size() - 1; i >= 0; --i) {
        if (vDistIdx[i].first < thDist) break;
        uRight[vDistIdx[i].second] = -1; depthOut[vDistIdx[i].second] = -1; --kept;
    }
    return kept;
}

}  // extern "C"
