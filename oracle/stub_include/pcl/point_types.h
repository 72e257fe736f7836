// TEST INFRASTRUCTURE — stand-in for the PCL names the reference's gaussian_model.h mentions (PCL is not in this
// image); see ../Eigen/Core.  gaussian_model.cpp itself never calls PCL.
#pragma once
#include <memory>
#include <vector>
namespace pcl {
struct PointXYZRGB { float x = 0, y = 0, z = 0; unsigned char r = 0, g = 0, b = 0; };
struct PointXYZ { float x = 0, y = 0, z = 0; };
template <class P> struct PointCloud { std::vector<P> points; using Ptr = std::shared_ptr<PointCloud<P>>; using ConstPtr = std::shared_ptr<const PointCloud<P>>; };
}  // namespace pcl
