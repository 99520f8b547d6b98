// TEST INFRASTRUCTURE ONLY -- CPU oracle for the ORB-extraction rows of SURVEY.md §8a.
//
// This file restates, on the CPU, the algorithm of the reference's
// PLVS2::ORBextractor (reference: src/ORBextractor.cc) plus the OpenCV 4.x
// primitives it calls (cv::resize INTER_LINEAR, cv::GaussianBlur 7x7 s=2,
// cv::FAST 9/16 + NMS, cv::fastAtan2).  It is the *checker* for the CUDA path:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load it.  The product (plvs_b200/) never links or calls it.
//
// Parity status: PINNED.  The reference ships no golden vectors for this path
// (SURVEY.md §4/§8c), so it is pinned to the reference itself: the OpenCV
// primitives restated here are checked bit-exactly against the cv2 4.13 build in
// this image (tests/test_oracle_orb.py: resize, blur, per-cell FAST incl. order,
// fastAtan2), and the whole extractor against the reference's own
// src/ORBextractor.cc compiled into oracle/_ref/liborb_ref.so by
// oracle/ref_build.py (tests/test_oracle_vs_reference_orb.py).
//
// Build: g++ -O2 -ffp-contract=off (no -march=native: the reference's Ubuntu
// 24.04 configuration, config.sh:19-21, so no FMA contraction).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

namespace {

inline int round_half_even(float v) { return (int)lrintf(v); }   // cvRound(float), SSE cvtss2si
inline int round_half_even(double v) { return (int)lrint(v); }   // cvRound(double)

const int kPatch = 31, kHalfPatch = 15, kEdge = 19;

static const int kPattern[1024] = {
#include "../plvs_b200/csrc/orb_pattern.inc"
};

}  // namespace

extern "C" {

// ---------------------------------------------------------------------------
// a1: constructor tables.  ref: src/ORBextractor.cc:446-523
// ---------------------------------------------------------------------------
void orc_orb_tables(int nfeatures, float scaleFactor, int nlevels,
                    float* scale, float* inv_scale, float* sigma2, float* inv_sigma2,
                    int* feat_per_level, int* umax /*16*/)
{
    scale[0] = 1.0f; sigma2[0] = 1.0f;
    for (int i = 1; i < nlevels; ++i) {
        scale[i] = scale[i - 1] * scaleFactor;           // float recurrence (:458-463)
        sigma2[i] = scale[i] * scale[i];
    }
    for (int i = 0; i < nlevels; ++i) {
        inv_scale[i] = 1.0f / scale[i];
        inv_sigma2[i] = 1.0f / sigma2[i];
    }
    float factor = 1.0f / scaleFactor;
    float want = nfeatures * (1 - factor) / (1 - (float)pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; ++l) {
        feat_per_level[l] = round_half_even(want);
        sum += feat_per_level[l];
        want *= factor;
    }
    feat_per_level[nlevels - 1] = std::max(nfeatures - sum, 0);

    // disc half-widths (:499-516)
    int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax[v] = round_half_even(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

// pyramid level size.  ref: src/ORBextractor.cc:1485-1486
void orc_level_size(int w0, int h0, float inv_scale, int* w, int* h)
{
    *w = round_half_even((float)w0 * inv_scale);
    *h = round_half_even((float)h0 * inv_scale);
}

// ---------------------------------------------------------------------------
// a2: cv::resize(INTER_LINEAR) on 8-bit single channel (OpenCV fixed point:
// 11-bit coefficients, SURVEY.md §8c' item 1).  call site: src/ORBextractor.cc:1494
// ---------------------------------------------------------------------------
static void linear_taps(int S, int D, std::vector<int>& idx0, std::vector<int>& idx1,
                        std::vector<short>& c0, std::vector<short>& c1)
{
    idx0.resize(D); idx1.resize(D); c0.resize(D); c1.resize(D);
    double inv = (double)D / S;
    double sc = 1. / inv;
    for (int d = 0; d < D; ++d) {
        float f = (float)((d + 0.5) * sc - 0.5);
        int s = (int)std::floor(f);
        f -= s;
        if (s < 0) { s = 0; f = 0.f; }
        if (s >= S - 1) { s = S - 1; f = 0.f; }
        idx0[d] = s;
        idx1[d] = std::min(s + 1, S - 1);
        int a0 = round_half_even((1.f - f) * 2048.f), a1 = round_half_even(f * 2048.f);
        c0[d] = (short)std::min(std::max(a0, -32768), 32767);
        c1[d] = (short)std::min(std::max(a1, -32768), 32767);
    }
}

void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride,
                          uint8_t* dst, int dw, int dh, int dstride)
{
    std::vector<int> x0, x1, y0, y1;
    std::vector<short> ax0, ax1, by0, by1;
    linear_taps(sw, dw, x0, x1, ax0, ax1);
    linear_taps(sh, dh, y0, y1, by0, by1);
    std::vector<int> r0(dw), r1(dw);
    for (int dy = 0; dy < dh; ++dy) {
        const uint8_t* s0 = src + (size_t)y0[dy] * sstride;
        const uint8_t* s1 = src + (size_t)y1[dy] * sstride;
        for (int dx = 0; dx < dw; ++dx) {
            r0[dx] = s0[x0[dx]] * ax0[dx] + s0[x1[dx]] * ax1[dx];
            r1[dx] = s1[x0[dx]] * ax0[dx] + s1[x1[dx]] * ax1[dx];
        }
        int b0 = by0[dy], b1 = by1[dy];
        uint8_t* d = dst + (size_t)dy * dstride;
        for (int dx = 0; dx < dw; ++dx)
            d[dx] = (uint8_t)((((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2);
    }
}

// ---------------------------------------------------------------------------
// a7: cv::GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) on an isolated 8-bit
// image (OpenCV 8-bit fixed-point path, SURVEY.md §8c' item 2).
// call site: src/ORBextractor.cc:1343-1344 (blur of a clone => border reflects the image itself)
// ---------------------------------------------------------------------------
static inline int reflect101(int p, int n)
{
    if (n == 1) return 0;
    while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; }
    return p;
}

void orc_gauss7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride)
{
    static const int tap[7] = {18, 34, 48, 56, 48, 34, 18};
    std::vector<uint16_t> rows((size_t)w * h);
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * sstride;
        for (int x = 0; x < w; ++x) {
            unsigned acc = 0;
            for (int k = -3; k <= 3; ++k) acc += tap[k + 3] * s[reflect101(x + k, w)];
            rows[(size_t)y * w + x] = (uint16_t)acc;                     // 8.8 fixed point
        }
    }
    for (int y = 0; y < h; ++y) {
        uint8_t* d = dst + (size_t)y * dstride;
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int k = -3; k <= 3; ++k) acc += (uint32_t)tap[k + 3] * rows[(size_t)reflect101(y + k, h) * w + x];
            d[x] = (uint8_t)((acc + 32768u) >> 16);                        // 16.16 -> u8, single rounding
        }
    }
}

// ---------------------------------------------------------------------------
// a3: FAST-9/16.  Corner score of OpenCV = max threshold for which the pixel
// is still a corner; "corner at t" <=> score >= t (SURVEY.md §8c' item 3).
// ---------------------------------------------------------------------------
static const int kCircle[16][2] = {{0, 3}, {1, 3}, {2, 2}, {3, 1}, {3, 0}, {3, -1}, {2, -2}, {1, -3},
                                   {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

static inline int fast_score_px(const uint8_t* p, int stride)
{
    int d[25];
    int v = p[0];
    for (int k = 0; k < 16; ++k) d[k] = v - p[kCircle[k][1] * stride + kCircle[k][0]];
    for (int k = 16; k < 25; ++k) d[k] = d[k - 16];
    int best = -1000;
    for (int k = 0; k < 16; ++k) {           // all 16 arcs of 9 contiguous ring pixels
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { mn = std::min(mn, d[k + j]); mx = std::max(mx, d[k + j]); }
        best = std::max(best, std::max(mn, -mx));   // brighter-centre arc / darker-centre arc
    }
    return best - 1;                          // < 0 when no arc has one sign
}

// score map over the whole image: out[y*ostride+x] = score clipped to [0,255]
// (0 also for the 3-px frame and for scores < min_th).
void orc_fast_score_map(const uint8_t* img, int w, int h, int stride, int min_th, uint8_t* out, int ostride)
{
    for (int y = 0; y < h; ++y) std::memset(out + (size_t)y * ostride, 0, w);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = fast_score_px(img + (size_t)y * stride + x, stride);
            out[(size_t)y * ostride + x] = (uint8_t)(s >= min_th ? std::min(s, 255) : 0);
        }
}

// cv::FAST(img(rect), kps, th, nonmax=true) for rect = [x0,x1) x [y0,y1): returns
// keypoints in OpenCV's emission order (raster), coordinates relative to the rect.
int orc_fast_rect(const uint8_t* img, int stride, int x0, int y0, int x1, int y1, int th,
                  int* xs, int* ys, int* resp, int cap)
{
    int w = x1 - x0, h = y1 - y0, n = 0;
    if (w < 7 || h < 7) return 0;
    std::vector<int> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = fast_score_px(img + (size_t)(y0 + y) * stride + x0 + x, stride);
            sc[(size_t)y * w + x] = s >= th ? s : 0;
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = sc[(size_t)y * w + x];
            if (!s) continue;
            const int* r = &sc[(size_t)y * w + x];
            if (s > r[-1] && s > r[1] && s > r[-w - 1] && s > r[-w] && s > r[-w + 1] &&
                s > r[w - 1] && s > r[w] && s > r[w + 1]) {
                if (n < cap) { xs[n] = x; ys[n] = y; resp[n] = s; }
                ++n;
            }
        }
    return n;
}

// ---------------------------------------------------------------------------
// a3 (cells): per-cell FAST with threshold fallback.  ref: src/ORBextractor.cc:867-998
// Output: candidates (x,y relative to the ROI origin (16,16), response) in the
// reference's order.  `score` may be NULL (then FAST is evaluated per cell);
// when non-NULL it is the full-image score map from orc_fast_score_map(min_th)
// (masking formulation, identical result).
// ---------------------------------------------------------------------------
int orc_fast_cells(const uint8_t* img, int w, int h, int stride, int ini_th, int min_th,
                   int* xs, int* ys, int* resp, int cap)
{
    const float W = 35;
    const int minBX = kEdge - 3, minBY = minBX;
    const int maxBX = w - kEdge + 3, maxBY = h - kEdge + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    if (width <= 0 || height <= 0) return 0;
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols == 0 || nRows == 0) return 0;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    int n = 0;
    std::vector<int> cx(4096), cy(4096), cr(4096);
    for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; ++j) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            int m = orc_fast_rect(img, stride, (int)iniX, (int)iniY, (int)maxX, (int)maxY, ini_th,
                                  cx.data(), cy.data(), cr.data(), 4096);
            if (m == 0)
                m = orc_fast_rect(img, stride, (int)iniX, (int)iniY, (int)maxX, (int)maxY, min_th,
                                  cx.data(), cy.data(), cr.data(), 4096);
            for (int k = 0; k < m; ++k) {
                if (n < cap) { xs[n] = cx[k] + j * wCell; ys[n] = cy[k] + i * hCell; resp[n] = cr[k]; }
                ++n;
            }
        }
    }
    return n;
}

// ---------------------------------------------------------------------------
// a4: DistributeOctTree.  ref: src/ORBextractor.cc:536-865
// Faithful list-based restatement (front insertion, erase, libstdc++ std::sort on
// (size, UL.x) with the reference's non-stable comparator, first-max per node).
// Candidates are (x,y,response) with integer-valued float coordinates relative to
// (minX,minY).  Returns the number of selected candidates; sel[] holds their
// indices into the input, in final list order.
// ---------------------------------------------------------------------------
namespace {
struct Cell {
    int ulx, uly, urx, bry;            // the only corner fields the reference's maths reads
    std::vector<int> pts;              // candidate indices, input order preserved
    bool leaf = false;
    std::list<Cell>::iterator self;
};

void split4(const Cell& c, const float* px, const float* py, Cell out[4])
{
    const int halfX = (int)std::ceil((float)(c.urx - c.ulx) / 2);
    const int halfY = (int)std::ceil((float)(c.bry - c.uly) / 2);
    const int midx = c.ulx + halfX, midy = c.uly + halfY;
    out[0].ulx = c.ulx; out[0].uly = c.uly; out[0].urx = midx;  out[0].bry = midy;
    out[1].ulx = midx;  out[1].uly = c.uly; out[1].urx = c.urx; out[1].bry = midy;
    out[2].ulx = c.ulx; out[2].uly = midy;  out[2].urx = midx;  out[2].bry = c.bry;
    out[3].ulx = midx;  out[3].uly = midy;  out[3].urx = c.urx; out[3].bry = c.bry;
    for (int q = 0; q < 4; ++q) { out[q].pts.clear(); out[q].leaf = false; }
    for (int id : c.pts) {
        if (px[id] < (float)midx) out[py[id] < (float)midy ? 0 : 2].pts.push_back(id);
        else                      out[py[id] < (float)midy ? 1 : 3].pts.push_back(id);
    }
    for (int q = 0; q < 4; ++q) if (out[q].pts.size() == 1) out[q].leaf = true;
}

bool node_less(const std::pair<int, Cell*>& a, const std::pair<int, Cell*>& b)
{
    if (a.first < b.first) return true;
    if (a.first > b.first) return false;
    return a.second->ulx < b.second->ulx;
}
}  // namespace

int orc_distribute_octree(int n, const float* px, const float* py, const float* resp,
                          int minX, int maxX, int minY, int maxY, int N, int* sel)
{
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni == 0) return 0;
    const float hX = (float)(maxX - minX) / nIni;

    std::list<Cell> cells;
    std::vector<Cell*> roots(nIni);
    for (int i = 0; i < nIni; ++i) {
        cells.emplace_back();
        Cell& c = cells.back();
        c.ulx = (int)(hX * (float)i);       c.uly = 0;
        c.urx = (int)(hX * (float)(i + 1)); c.bry = maxY - minY;
        roots[i] = &c;
    }
    for (int i = 0; i < n; ++i) roots[(size_t)(px[i] / hX)]->pts.push_back(i);

    for (auto it = cells.begin(); it != cells.end();) {
        if (it->pts.size() == 1) { it->leaf = true; ++it; }
        else if (it->pts.empty()) it = cells.erase(it);
        else ++it;
    }

    std::vector<std::pair<int, Cell*>> expandable;
    bool done = false;
    Cell kids[4];
    auto push_kids = [&](bool count, int& nToExpand) {
        for (int q = 0; q < 4; ++q) {
            if (kids[q].pts.empty()) continue;
            cells.push_front(kids[q]);
            if (kids[q].pts.size() > 1) {
                if (count) ++nToExpand;
                expandable.emplace_back((int)kids[q].pts.size(), &cells.front());
                cells.front().self = cells.begin();
            }
        }
    };
    while (!done) {
        int prev = (int)cells.size();
        int nToExpand = 0;
        expandable.clear();
        for (auto it = cells.begin(); it != cells.end();) {
            if (it->leaf) { ++it; continue; }
            split4(*it, px, py, kids);
            push_kids(true, nToExpand);
            it = cells.erase(it);
        }
        if ((int)cells.size() >= N || (int)cells.size() == prev) {
            done = true;
        } else if ((int)cells.size() + nToExpand * 3 > N) {
            while (!done) {
                prev = (int)cells.size();
                std::vector<std::pair<int, Cell*>> order = expandable;
                expandable.clear();
                std::sort(order.begin(), order.end(), node_less);
                for (int j = (int)order.size() - 1; j >= 0; --j) {
                    split4(*order[j].second, px, py, kids);
                    int dummy = 0;
                    push_kids(false, dummy);
                    cells.erase(order[j].second->self);
                    if ((int)cells.size() >= N) break;
                }
                if ((int)cells.size() >= N || (int)cells.size() == prev) done = true;
            }
        }
    }
    int m = 0;
    for (const Cell& c : cells) {
        int best = c.pts[0];
        float br = resp[best];
        for (size_t k = 1; k < c.pts.size(); ++k)
            if (resp[c.pts[k]] > br) { best = c.pts[k]; br = resp[best]; }
        sel[m++] = best;
    }
    return m;
}

// ---------------------------------------------------------------------------
// a6: scalar cv::fastAtan2 (degrees) -- SURVEY.md §8c' item 4 -- and IC_Angle.
// ref: src/ORBextractor.cc:110-137
// ---------------------------------------------------------------------------
float orc_fast_atan2(float y, float x)
{
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float eps = (float)2.2204460492503131e-16;
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + eps); c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + eps); c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

float orc_ic_angle(const uint8_t* img, int stride, float ptx, float pty, const int* umax)
{
    int m01 = 0, m10 = 0;
    const uint8_t* c = img + (size_t)round_half_even(pty) * stride + round_half_even(ptx);
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vs = 0, d = umax[v];
        for (int u = -d; u <= d; ++u) {
            int lo = c[u + v * stride], hi = c[u - v * stride];
            vs += lo - hi;
            m10 += u * (lo + hi);
        }
        m01 += v * vs;
    }
    return orc_fast_atan2((float)m01, (float)m10);
}

// ---------------------------------------------------------------------------
// a8: steered rBRIEF-256.  ref: src/ORBextractor.cc:139-180.  cos/sin are glibc
// cosf/sinf (float overloads through `using namespace std`).
// ---------------------------------------------------------------------------
void orc_orb_descriptor(const uint8_t* img, int stride, float ptx, float pty, float angle_deg, uint8_t* desc)
{
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = angle_deg * factorPI;
    const float a = cosf(ang), b = sinf(ang);
    const uint8_t* c = img + (size_t)round_half_even(pty) * stride + round_half_even(ptx);
    const int* p = kPattern;
    for (int i = 0; i < 32; ++i) {
        int val = 0;
        for (int k = 0; k < 8; ++k, p += 4) {
            float x0 = (float)p[0], y0 = (float)p[1], x1 = (float)p[2], y1 = (float)p[3];
            int t0 = c[round_half_even(x0 * b + y0 * a) * stride + round_half_even(x0 * a - y0 * b)];
            int t1 = c[round_half_even(x1 * b + y1 * a) * stride + round_half_even(x1 * a - y1 * b)];
            val |= (t0 < t1) << k;
        }
        desc[i] = (uint8_t)val;
    }
}

// steering pair as libm gives it (used by tests that pin the CUDA sincos port)
void orc_steer(float angle_deg, float* a, float* b)
{
    const float factorPI = (float)(3.14159265358979323846 / 180.f);
    const float ang = angle_deg * factorPI;
    *a = cosf(ang); *b = sinf(ang);
}

// ---------------------------------------------------------------------------
// Whole extractor, all-C port (the CPU arm that needs no cv2).
// ref: src/ORBextractor.cc:1245-1389 (operator()), :1481-1506 (ComputePyramid),
// :867-1052 (ComputeKeyPointsOctTree).  Keypoint record = 7 floats:
// x, y, size, angle, response, octave, class_id(-1).
// Returns number of keypoints (<= cap) and *mono_index.
// ---------------------------------------------------------------------------
int orc_orb_extract(const uint8_t* gray, int w0, int h0, int stride0,
                    int nfeatures, float scaleFactor, int nlevels, int ini_th, int min_th,
                    int lap0, int lap1, float* kps /*cap*7*/, uint8_t* desc /*cap*32*/, int cap,
                    int* mono_index, int* n_candidates)
{
    std::vector<float> sc(nlevels), isc(nlevels), s2(nlevels), is2(nlevels);
    std::vector<int> quota(nlevels);
    int umax[16];
    orc_orb_tables(nfeatures, scaleFactor, nlevels, sc.data(), isc.data(), s2.data(), is2.data(), quota.data(), umax);

    std::vector<std::vector<uint8_t>> pyr(nlevels);
    std::vector<int> lw(nlevels), lh(nlevels);
    for (int l = 0; l < nlevels; ++l) {
        orc_level_size(w0, h0, isc[l], &lw[l], &lh[l]);
        pyr[l].resize((size_t)lw[l] * lh[l]);
        if (l == 0) for (int y = 0; y < h0; ++y) std::memcpy(&pyr[0][(size_t)y * w0], gray + (size_t)y * stride0, w0);
        else orc_resize_linear_u8(pyr[l - 1].data(), lw[l - 1], lh[l - 1], lw[l - 1], pyr[l].data(), lw[l], lh[l], lw[l]);
    }

    struct KP { float x, y, size, angle, resp; int octave; };
    std::vector<std::vector<KP>> all(nlevels);
    int ncand_total = 0;
    const int ccap = 1 << 20;
    std::vector<int> cx(ccap), cy(ccap), cr(ccap), sel(ccap);
    std::vector<float> fx, fy, fr;
    for (int l = 0; l < nlevels; ++l) {
        int nc = orc_fast_cells(pyr[l].data(), lw[l], lh[l], lw[l], ini_th, min_th, cx.data(), cy.data(), cr.data(), ccap);
        if (nc > ccap) nc = ccap;
        ncand_total += nc;
        fx.assign(cx.begin(), cx.begin() + nc); fy.assign(cy.begin(), cy.begin() + nc); fr.assign(cr.begin(), cr.begin() + nc);
        const int minB = kEdge - 3;
        int m = orc_distribute_octree(nc, fx.data(), fy.data(), fr.data(), minB, lw[l] - kEdge + 3, minB, lh[l] - kEdge + 3,
                                      quota[l], sel.data());
        const int scaledPatch = (int)(kPatch * sc[l]);
        for (int k = 0; k < m; ++k) {
            KP kp;
            kp.x = fx[sel[k]] + minB; kp.y = fy[sel[k]] + minB;
            kp.size = (float)scaledPatch; kp.resp = fr[sel[k]]; kp.octave = l;
            kp.angle = orc_ic_angle(pyr[l].data(), lw[l], kp.x, kp.y, umax);
            all[l].push_back(kp);
        }
    }
    if (n_candidates) *n_candidates = ncand_total;

    int total = 0;
    for (int l = 0; l < nlevels; ++l) total += (int)all[l].size();
    if (total > cap) return -total;
    int mono = 0, stereo = total - 1;
    std::vector<uint8_t> blur;
    for (int l = 0; l < nlevels; ++l) {
        if (all[l].empty()) continue;
        blur.resize((size_t)lw[l] * lh[l]);
        orc_gauss7_u8(pyr[l].data(), lw[l], lh[l], lw[l], blur.data(), lw[l]);
        for (KP& kp : all[l]) {
            uint8_t d[32];
            orc_orb_descriptor(blur.data(), lw[l], kp.x, kp.y, kp.angle, d);
            if (l != 0) { kp.x *= sc[l]; kp.y *= sc[l]; }
            int slot = (kp.x >= lap0 && kp.x <= lap1) ? stereo-- : mono++;
            float* o = kps + (size_t)slot * 7;
            o[0] = kp.x; o[1] = kp.y; o[2] = kp.size; o[3] = kp.angle; o[4] = kp.resp; o[5] = (float)kp.octave; o[6] = -1.f;
            std::memcpy(desc + (size_t)slot * 32, d, 32);
        }
    }
    *mono_index = mono;
    return total;
}

// ---------------------------------------------------------------------------
// Step after extraction: Frame::UndistortKeyPoints (src/Frame.cc:1507-1553) =
// cv::undistortPoints(mat, mat, K, distCoef, Mat(), K) on the N x 2 float matrix of keypoint coordinates.
// OpenCV's routine (cvUndistortPointsInternal, criteria = 5 iterations, no tilt): everything in double,
// results rounded to float.  Pinned: tests/test_oracle_orb.py compares with the real cv2.undistortPoints.
// K = fx, fy, cx, cy and dist = k1, k2, p1, p2[, k3[, k4, k5, k6[, s1..s4]]] as the FLOAT values PLVS keeps in
// its CV_32F matrices.  dist[0] == 0 means "no distortion": the keypoints are copied (Frame.cc:1510-1514).
// ---------------------------------------------------------------------------
void orc_undistort_points(const float* xy, int n, const float* K, const float* dist, int ndist, float* out)
{
    if (ndist <= 0 || dist[0] == 0.0f) { std::memcpy(out, xy, sizeof(float) * 2 * (size_t)n); return; }
    const double fx = K[0], fy = K[1], cx = K[2], cy = K[3];
    double k[14] = {0};
    for (int i = 0; i < ndist && i < 14; ++i) k[i] = dist[i];
    const double ifx = 1. / fx, ify = 1. / fy;
    for (int i = 0; i < n; ++i) {
        const double u = xy[2 * i], v = xy[2 * i + 1];
        double x = (u - cx) * ifx, y = (v - cy) * ify;
        const double x0 = x, y0 = y;
        for (int j = 0; j < 5; ++j) {
            const double r2 = x * x + y * y;
            const double icdist = (1 + ((k[7] * r2 + k[6]) * r2 + k[5]) * r2) / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
            if (icdist < 0) { x = (u - cx) * ifx; y = (v - cy) * ify; break; }
            const double deltaX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x) + k[8] * r2 + k[9] * r2 * r2;
            const double deltaY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y + k[10] * r2 + k[11] * r2 * r2;
            x = (x0 - deltaX) * icdist;
            y = (y0 - deltaY) * icdist;
        }
        const double xx = fx * x + 0 * y + cx, yy = 0 * x + fy * y + cy, ww = 1. / (0 * x + 0 * y + 1);
        out[2 * i] = (float)(xx * ww); out[2 * i + 1] = (float)(yy * ww);
    }
}

}  // extern "C"
