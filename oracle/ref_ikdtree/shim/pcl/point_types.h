/*
 * Stand-in for <pcl/point_types.h> so that the reference's ikd_Tree.{h,cpp} compile UNCHANGED
 * from /root/reference (PCL and Eigen are not installed in this image).  TEST INFRASTRUCTURE.
 * Only what ikd_Tree.h / ikd_Tree.cpp name is provided: Eigen::aligned_allocator, the three pcl
 * point types of the explicit instantiations (ikd_Tree.cpp:1724-1726) and LIMO-Velo's 32-byte
 * Point (include/Headers/Objects.hpp:20-28: float x,y,z; double time; float intensity, range).
 */
#pragma once
#include <cmath>
#include <memory>
#include <vector>

namespace Eigen {
template <class T>
using aligned_allocator = std::allocator<T>;
}

namespace pcl {
struct PointXYZ { float x, y, z; };
struct PointXYZI { float x, y, z, intensity; };
struct PointXYZINormal { float x, y, z, intensity, normal_x, normal_y, normal_z, curvature; };
}  // namespace pcl

struct Point {
    float x, y, z;
    double time;
    float intensity, range;
};
