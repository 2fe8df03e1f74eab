// libfast_gicp.so of the B200 drop-in: the explicit instantiations the reference ships
// (Mapping/src/fast_gicp/src/fast_gicp/gicp/fast_gicp.cpp:4-6 and lsq_registration.cpp:4-6), over the C-ABI.
#include <limits>

#include "fast_gicp_b200/pcl_dropin_impl.hpp"

template class fast_gicp::LsqRegistration<pcl::PointXYZ, pcl::PointXYZ>;
template class fast_gicp::LsqRegistration<pcl::PointXYZI, pcl::PointXYZI>;
template class fast_gicp::LsqRegistration<pcl::PointNormal, pcl::PointNormal>;
template class fast_gicp::FastGICP<pcl::PointXYZ, pcl::PointXYZ>;
template class fast_gicp::FastGICP<pcl::PointXYZI, pcl::PointXYZI>;
template class fast_gicp::FastGICP<pcl::PointNormal, pcl::PointNormal>;
