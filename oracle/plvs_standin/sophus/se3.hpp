// TEST INFRASTRUCTURE ONLY -- minimal Sophus::SE3f / Sim3f so that the reference's src/ORBmatcher.cc compiles
// (Sophus needs the real Eigen, which is not installed).  See plvs_types.hpp: the parity harness feeds identity poses and
// an identity "camera", so none of this arithmetic takes part in a compared result; rotation is kept as a plain matrix.
#ifndef PLVS_B200_SOPHUS_STANDIN
#define PLVS_B200_SOPHUS_STANDIN
#include <Eigen/Core>
#include <Eigen/Geometry>
namespace Sophus {
template <typename T> struct SE3 {
    Eigen::Matrix<T, 3, 3> R; Eigen::Matrix<T, 3, 1> t;
    SE3() : R(Eigen::Matrix<T, 3, 3>::Identity()), t(Eigen::Matrix<T, 3, 1>::Zero()) {}
    SE3(const Eigen::Matrix<T, 3, 3>& r, const Eigen::Matrix<T, 3, 1>& tt) : R(r), t(tt) {}
    Eigen::Matrix<T, 3, 3> rotationMatrix() const { return R; }
    Eigen::Matrix<T, 3, 1> translation() const { return t; }
    SE3 inverse() const { const Eigen::Matrix<T, 3, 3> Rt = R.transpose(); return SE3(Rt, -(Rt * t)); }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const { return R * p + t; }
    SE3 operator*(const SE3& o) const { return SE3(R * o.R, R * o.t + t); }
};
typedef SE3<float> SE3f;
template <typename T> struct Sim3 {
    Eigen::Matrix<T, 3, 3> R; Eigen::Matrix<T, 3, 1> t; T s;
    Sim3() : R(Eigen::Matrix<T, 3, 3>::Identity()), t(Eigen::Matrix<T, 3, 1>::Zero()), s(1) {}
    Eigen::Matrix<T, 3, 3> rotationMatrix() const { return R; }
    Eigen::Matrix<T, 3, 1> translation() const { return t; }
    T scale() const { return s; }
    Sim3 inverse() const { Sim3 o; o.R = R.transpose(); o.s = T(1) / s; o.t = -((o.R * t) * o.s); return o; }
    Eigen::Matrix<T, 3, 1> operator*(const Eigen::Matrix<T, 3, 1>& p) const { return (R * p) * s + t; }
};
typedef Sim3<float> Sim3f;
}  // namespace Sophus
#endif
