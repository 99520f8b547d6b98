// TEST INFRASTRUCTURE ONLY -- a minimal stand-in for the parts of the OpenCV C++ API that the reference's
// src/ORBextractor.cc uses (OpenCV's C++ headers are not installed in this image; only the Python cv2 module is).
//
// It lets oracle/ref_build.py compile src/ORBextractor.cc UNMODIFIED, where it lies under /root/reference, into
// oracle/_ref/liborb_ref.so.  The reference's own code then provides everything that is PLVS/ORB-SLAM logic: scale tables
// and feature quotas, pyramid construction, the per-cell FAST loop with its threshold fallback, DistributeOctTree (with
// the real std::list / std::sort), IC_Angle, the steered rBRIEF sampling, the lapping-area assembly.  What this header
// supplies is only cv::Mat bookkeeping (8-bit single-channel, ROI views, clone/copyTo) and the OpenCV *primitives*,
// forwarded to the C restatements in oracle/orb_oracle.cpp, each of which tests/test_oracle_orb.py checks bit-exactly
// against the real OpenCV (cv2) of this image:
//   cv::resize(INTER_LINEAR)      -> orc_resize_linear_u8        cv::GaussianBlur(7x7, 2, REFLECT_101) -> orc_gauss7_u8
//   cv::FAST(img, kps, th, true)  -> orc_fast_rect               cv::fastAtan2 -> orc_fast_atan2
//   cvRound = lrint (SSE cvtss2si, round-half-even), cvFloor/cvCeil = floor/ceil, copyMakeBorder = reflect-101 fill.
#ifndef PLVS_B200_CV_STANDIN
#define PLVS_B200_CV_STANDIN
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <vector>

typedef unsigned char uchar;
#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5      // named by DBoW2's FORB::toMat32F (never called here; this Mat holds bytes)
#define CV_PI 3.1415926535897932384626433832795

extern "C" {
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
void orc_gauss7_u8(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
int orc_fast_rect(const uint8_t* img, int stride, int x0, int y0, int x1, int y1, int th, int* xs, int* ys, int* resp, int cap);
float orc_fast_atan2(float y, float x);
}

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { return (int)std::floor(v); }
inline int cvCeil(double v) { return (int)std::ceil(v); }

namespace cv {

template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <typename U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
};
// OpenCV types.hpp: `a.x = saturate_cast<_Tp>(a.x * b)` with b float (the overload ORBextractor.cc:1364 selects)
template <typename T> inline Point_<T>& operator*=(Point_<T>& a, float b) { a.x = (T)(a.x * b); a.y = (T)(a.y * b); return a; }
template <typename T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <typename T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
typedef Point_<int> Point2i;
typedef Point2i Point;
typedef Point_<float> Point2f;

struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int a, int b, int w, int h) : x(a), y(b), width(w), height(h) {} };
struct Scalar { double v[4]; Scalar() : v{0, 0, 0, 0} {} Scalar(double a) : v{a, 0, 0, 0} {} };
struct Range { int start, end; Range(int a, int b) : start(a), end(b) {} };

struct KeyPoint {           // same layout as cv::KeyPoint (28 bytes)
    Point2f pt; float size, angle, response; int octave, class_id;
    KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };
enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

class _OutputArray;

class Mat {
public:
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    size_t step = 0;
    std::shared_ptr<std::vector<uchar>> buf;

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int /*type*/, void* p, size_t st = 0) : rows(r), cols(c), data((uchar*)p), step(st ? st : (size_t)c) {}

    void create(int r, int c, int /*type*/)
    {
        if (data && r == rows && c == cols) return;          // cv::Mat::create keeps a matrix of the right size
        buf = std::make_shared<std::vector<uchar>>((size_t)r * c);
        rows = r; cols = c; step = (size_t)c; data = buf->data();
    }
    void create(Size s, int type) { create(s.height, s.width, type); }
    void release() { buf.reset(); data = nullptr; rows = cols = 0; step = 0; }
    static Mat zeros(int r, int c, int type) { Mat m(r, c, type); if (m.data) std::memset(m.data, 0, (size_t)r * c); return m; }

    bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
    int type() const { return CV_8UC1; }
    int channels() const { return 1; }
    Size size() const { return Size(cols, rows); }
    size_t step1() const { return step; }
    bool isContinuous() const { return step == (size_t)cols || rows == 1; }

    // signed offsets: the reference takes windows that reach into the border frame around a pyramid level (src/Frame.cc:1886-1916)
    Mat operator()(const Rect& r) const { Mat m; m.rows = r.height; m.cols = r.width; m.step = step; m.data = data + (std::ptrdiff_t)r.y * (std::ptrdiff_t)step + r.x; m.buf = buf; return m; }
    Mat rowRange(int a, int b) const { return (*this)(Rect(0, a, cols, b - a)); }
    Mat colRange(int a, int b) const { return (*this)(Rect(a, 0, b - a, rows)); }
    Mat row(int i) const { return (*this)(Rect(0, i, cols, 1)); }
    Mat clone() const { Mat m(rows, cols, CV_8UC1); for (int y = 0; y < rows; ++y) std::memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, cols); return m; }
    void copyTo(const _OutputArray& dst) const;

    uchar* ptr(int i = 0) { return data + (size_t)i * step; }
    const uchar* ptr(int i = 0) const { return data + (size_t)i * step; }
    template <typename T> T* ptr(int i = 0) { return (T*)(data + (size_t)i * step); }
    template <typename T> const T* ptr(int i = 0) const { return (const T*)(data + (size_t)i * step); }
    template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    // single index: element i of a row or column vector (the query masks of the line matcher are n x 1)
    template <typename T> T& at(int i) { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    template <typename T> const T& at(int i) const { return rows == 1 ? at<T>(0, i) : at<T>(i, 0); }
    // append the rows of m (8-bit, same width): BinaryDescriptorMatcher::add, not on any path the harnesses call
    void push_back(const Mat& m)
    {
        if (m.empty()) return;
        Mat out(rows + m.rows, m.cols, CV_8UC1);
        for (int y = 0; y < rows; ++y) std::memcpy(out.data + (size_t)y * out.step, data + (size_t)y * step, (size_t)cols);
        for (int y = 0; y < m.rows; ++y) std::memcpy(out.data + (size_t)(rows + y) * out.step, m.data + (size_t)y * m.step, (size_t)m.cols);
        *this = out;
    }
};

// bookkeeping types Thirdparty/line_descriptor's matcher needs besides Mat (oracle/ref_build.py::build_linematch)
class Algorithm { public: virtual ~Algorithm() {} };
struct DMatch {
    int queryIdx, trainIdx, imgIdx; float distance;
    DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.4028235e38f) {}
    DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};
template <class T> class Ptr : public std::shared_ptr<T> {
public:
    Ptr() {}
    explicit Ptr(T* p) : std::shared_ptr<T>(p) {}
    Ptr(const std::shared_ptr<T>& p) : std::shared_ptr<T>(p) {}
    void release() { this->reset(); }
    bool empty() const { return !*this; }
};
template <class T, class... A> inline Ptr<T> makePtr(A&&... a) { return Ptr<T>(new T(std::forward<A>(a)...)); }

// cv::InputArray / cv::OutputArray are `const _InputArray&` / `const _OutputArray&` in OpenCV too
class _InputArray {
public:
    Mat m; bool none = true;
    _InputArray() {}
    _InputArray(const Mat& a) : m(a), none(false) {}
    Mat getMat() const { return m; }
    bool empty() const { return none || m.empty(); }
};
class _OutputArray {
public:
    Mat* target = nullptr; mutable Mat view;
    _OutputArray() {}
    _OutputArray(Mat& a) : target(&a) {}
    _OutputArray(const Mat& a) : view(a) {}                 // a temporary header (e.g. descriptors.row(i)): shares the pixels
    Mat& ref() const { return target ? *target : view; }
    Mat getMat() const { return ref(); }
    void create(int r, int c, int type) const { ref().create(r, c, type); }
    void release() const { ref().release(); }
    bool empty() const { return ref().empty(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline _InputArray noArray() { return _InputArray(); }
// cv::norm(a, b, NORM_L1) on 8-bit images: sum of absolute differences (exact in integers; OpenCV returns it as double)
inline double norm(const _InputArray& a_, const _InputArray& b_, int normType)
{
    assert(normType == NORM_L1); (void)normType;
    const Mat a = a_.getMat(), b = b_.getMat();
    assert(a.rows == b.rows && a.cols == b.cols);
    long long s = 0;
    for (int y = 0; y < a.rows; ++y) {
        const uchar* pa = a.data + (std::ptrdiff_t)y * (std::ptrdiff_t)a.step; const uchar* pb = b.data + (std::ptrdiff_t)y * (std::ptrdiff_t)b.step;
        for (int x = 0; x < a.cols; ++x) s += pa[x] > pb[x] ? pa[x] - pb[x] : pb[x] - pa[x];
    }
    return (double)s;
}

inline void Mat::copyTo(const _OutputArray& dst) const
{
    Mat& d = dst.ref();
    d.create(rows, cols, CV_8UC1);
    for (int y = 0; y < rows; ++y) std::memmove(d.data + (size_t)y * d.step, data + (size_t)y * step, cols);
}

inline float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

inline void resize(InputArray src_, OutputArray dst_, Size sz, double = 0, double = 0, int interpolation = INTER_LINEAR)
{
    assert(interpolation == INTER_LINEAR); (void)interpolation;
    const Mat src = src_.getMat();
    dst_.create(sz.height, sz.width, CV_8UC1);
    Mat dst = dst_.getMat();
    orc_resize_linear_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, dst.cols, dst.rows, (int)dst.step);
}

inline int reflect101(int p, int n) { while (p < 0 || p >= n) { if (p < 0) p = -p; else p = 2 * n - 2 - p; } return p; }

inline void copyMakeBorder(InputArray src_, OutputArray dst_, int top, int bottom, int left, int right, int borderType, const Scalar& = Scalar())
{
    assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101); (void)borderType;
    const Mat src = src_.getMat();
    dst_.create(src.rows + top + bottom, src.cols + left + right, CV_8UC1);
    Mat dst = dst_.getMat();
    std::vector<uchar> tmp;                                  // the source may be the interior of the destination
    const uchar* s = src.data; size_t sstep = src.step;
    if (src.data >= dst.data && src.data < dst.data + (size_t)dst.rows * dst.step) {
        tmp.resize((size_t)src.rows * src.cols);
        for (int y = 0; y < src.rows; ++y) std::memcpy(&tmp[(size_t)y * src.cols], src.data + (size_t)y * src.step, src.cols);
        s = tmp.data(); sstep = (size_t)src.cols;
    }
    for (int y = 0; y < dst.rows; ++y) {
        const uchar* srow = s + (size_t)reflect101(y - top, src.rows) * sstep;
        uchar* drow = dst.data + (size_t)y * dst.step;
        for (int x = 0; x < dst.cols; ++x) drow[x] = srow[reflect101(x - left, src.cols)];
    }
}

inline void GaussianBlur(InputArray src_, OutputArray dst_, Size ksize, double sx, double sy = 0, int borderType = BORDER_DEFAULT)
{
    assert(ksize.width == 7 && ksize.height == 7 && sx == 2 && sy == 2 && borderType == BORDER_REFLECT_101); (void)ksize; (void)sx; (void)sy; (void)borderType;
    const Mat src = src_.getMat().clone();                   // in-place calls are allowed
    dst_.create(src.rows, src.cols, CV_8UC1);
    Mat dst = dst_.getMat();
    orc_gauss7_u8(src.data, src.cols, src.rows, (int)src.step, dst.data, (int)dst.step);
}

inline void FAST(InputArray img_, std::vector<KeyPoint>& kps, int threshold, bool nonmax = true)
{
    assert(nonmax); (void)nonmax;
    const Mat img = img_.getMat();
    kps.clear();
    const int cap = std::max(1, img.rows * img.cols);
    std::vector<int> xs(cap), ys(cap), rs(cap);
    const int n = orc_fast_rect(img.data, (int)img.step, 0, 0, img.cols, img.rows, threshold, xs.data(), ys.data(), rs.data(), cap);
    for (int i = 0; i < n; ++i) kps.push_back(KeyPoint((float)xs[i], (float)ys[i], 7.f, -1, (float)rs[i]));
}

struct KeyPointsFilter {       // only referenced by ComputeKeyPointsOld, which operator() does not call (ORBextractor.cc:1269)
    static void retainBest(std::vector<KeyPoint>& kps, int n)
    {
        if (n >= 0 && (int)kps.size() > n) {
            std::stable_sort(kps.begin(), kps.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
            kps.resize(n);
        }
    }
};

// cv::FileStorage / cv::FileNode: DBoW2's TemplatedVocabulary.h names them in its YAML save/load members.  The oracle builds read vocabularies
// with loadFromTextFile and never call those members; the stubs exist so that the header parses and links, and abort if reached.
[[noreturn]] inline void filestorage_unavailable() { std::fprintf(stderr, "cv stand-in: cv::FileStorage is not available\n"); std::abort(); }
class FileNode {
public:
    enum { NONE = 0, SEQ = 4, MAP = 5 };
    FileNode operator[](const char*) const { filestorage_unavailable(); }
    FileNode operator[](const std::string&) const { filestorage_unavailable(); }
    FileNode operator[](int) const { filestorage_unavailable(); }
    int type() const { filestorage_unavailable(); }
    size_t size() const { filestorage_unavailable(); }
    operator int() const { filestorage_unavailable(); }
    operator float() const { filestorage_unavailable(); }
    operator double() const { filestorage_unavailable(); }
    operator std::string() const { filestorage_unavailable(); }
};
class FileStorage {
public:
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string&, int) {}
    bool isOpened() const { return false; }
    void release() {}
    FileNode operator[](const char*) const { filestorage_unavailable(); }
    FileNode operator[](const std::string&) const { filestorage_unavailable(); }
};
template <class T> FileStorage& operator<<(FileStorage&, const T&) { filestorage_unavailable(); }

}  // namespace cv
#endif
