// Minimal stand-ins for the OpenCV / Eigen types the shim touches, so that shim/plvs_shim.hpp can be
// compiled and linked in an image without OpenCV C++ headers or Eigen (tests/test_shim_compile.py).
// NOT a replacement for those libraries: only the members the shim uses exist.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0

namespace cv {
struct KeyPoint { float pt_x, pt_y, size, angle, response; int octave, class_id; };
class Mat {
public:
    int rows = 0, cols = 0; size_t step = 0; unsigned char* data = nullptr;
    Mat() {}
    Mat(int r, int c, int /*type*/) { create(r, c, 0); }
    void create(int r, int c, int /*type*/)
    {
        if (r == rows && c == cols && data) return;
        rows = r; cols = c; step = (size_t)c;
        buf_ = std::shared_ptr<unsigned char>(new unsigned char[(size_t)r * c + 1], std::default_delete<unsigned char[]>());
        data = buf_.get();
    }
    bool empty() const { return rows == 0 || cols == 0 || !data; }
    void release() { rows = cols = 0; data = nullptr; buf_.reset(); }
    Mat getMat() const { return *this; }
    Mat rowRange(int a, int b) const { Mat m = *this; m.rows = b - a; m.data = data + (size_t)a * step; return m; }
    void copyTo(Mat& dst) const { dst.create(rows, cols, 0); for (int r = 0; r < rows; ++r) std::memcpy(dst.data + r * dst.step, data + r * step, cols); }
private:
    std::shared_ptr<unsigned char> buf_;
};
typedef const Mat& InputArray;
typedef Mat& OutputArray;
}  // namespace cv

namespace Eigen {
struct Affine3f {
    float R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, t[3] = {0, 0, 0};
    struct Lin { const float (*m)[3]; float operator()(int r, int c) const { return m[r][c]; } };
    struct Tr { const float* v; float operator()(int r) const { return v[r]; } };
    Lin linear() const { return Lin{R}; }
    Tr translation() const { return Tr{t}; }
};
}  // namespace Eigen
