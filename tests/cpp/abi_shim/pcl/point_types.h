// TEST INFRASTRUCTURE ONLY -- layouts of the three PCL 1.10 point types the reference instantiates FastGICP for
// (FG/src/fast_gicp/gicp/fast_gicp.cpp:4-6).  PCL is not installed in this image.
#pragma once
namespace pcl {
struct alignas(16) PointXYZ {
  float x, y, z, pad_;
};
struct alignas(16) PointXYZI {
  float x, y, z, pad0_;
  float intensity, pad1_[3];
};
struct alignas(16) PointNormal {
  float x, y, z, pad0_;
  float normal_x, normal_y, normal_z, pad1_;
  float curvature, pad2_[3];
};
}  // namespace pcl
