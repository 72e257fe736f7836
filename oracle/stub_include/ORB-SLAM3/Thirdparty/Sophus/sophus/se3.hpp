// TEST INFRASTRUCTURE — stand-in for Sophus::SE3 (the real header needs Eigen, absent from this image); see
// ../../../../Eigen/Core for why this exists.  Pose algebra aborts if it is ever called.
#pragma once
#include <Eigen/Core>
namespace Sophus {
template <typename T>
struct SE3 {
    Eigen::Matrix<T, 3, 3> R_ = Eigen::Matrix<T, 3, 3>::Identity();
    Eigen::Matrix<T, 3, 1> t_;
    SE3() = default;
    SE3(const Eigen::Matrix<T, 3, 3>& R, const Eigen::Matrix<T, 3, 1>& t) : R_(R), t_(t) {}
    SE3(const Eigen::Quaternion<T>&, const Eigen::Matrix<T, 3, 1>& t) : t_(t) {}
    Eigen::Matrix<T, 3, 3> rotationMatrix() const { return R_; }
    Eigen::Matrix<T, 3, 1>& translation() { return t_; }
    const Eigen::Matrix<T, 3, 1>& translation() const { return t_; }
    Eigen::Quaternion<T> unit_quaternion() const { Eigen::stub_abort(); }
    Eigen::Matrix<T, 4, 4> matrix() const { Eigen::stub_abort(); }
    SE3 inverse() const { Eigen::stub_abort(); }
    template <typename U> SE3<U> cast() const { Eigen::stub_abort(); }
    SE3 operator*(const SE3&) const { Eigen::stub_abort(); }
};
using SE3f = SE3<float>; using SE3d = SE3<double>;
}  // namespace Sophus
