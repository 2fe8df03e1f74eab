// Drives the drop-in fast_gicp::FastGICP<pcl::PointXYZI, pcl::PointXYZI> exactly as GlobalManager::ICPCheck does
// (Mapping/src/global_manager/src/global_manager.cpp:2435-2443 construction, :2018-2021 use): through a
// pcl::Registration<...>::Ptr, setInputSource / setInputTarget / align(out, guess) / hasConverged / getFinalTransformation.
// Links against mr_slam_b200/libfast_gicp.so (explicit instantiations) -- this TU does NOT include the implementation header.
#include <fast_gicp/gicp/fast_gicp.hpp>

#include <cmath>
#include <cstdio>

using PointTI = pcl::PointXYZI;

static pcl::Registration<PointTI, PointTI>::Ptr select_registration_method() {
  fast_gicp::FastGICP<PointTI, PointTI>::Ptr gicp(new fast_gicp::FastGICP<PointTI, PointTI>());
  gicp->setNumThreads(8);
  gicp->setTransformationEpsilon(1e-3);
  gicp->setMaximumIterations(50);
  gicp->setMaxCorrespondenceDistance(100.0);
  gicp->setCorrespondenceRandomness(15);
  return gicp;
}

int main() {
  if (mrslam_init(0) != MRSLAM_OK) {
    std::printf("NO_GPU %s\n", mrslam_last_error());
    return 0;
  }
  pcl::PointCloud<PointTI>::Ptr tgt(new pcl::PointCloud<PointTI>()), src(new pcl::PointCloud<PointTI>()), out(new pcl::PointCloud<PointTI>());
  unsigned s = 4242;
  auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (float)(s >> 8) / 16777216.0f; };
  for (int i = 0; i < 6000; i++) {
    float x = rnd() * 20 - 10, y = rnd() * 20 - 10, z = 0.3f * std::sin(0.5f * x) + 0.2f * std::cos(0.7f * y);
    PointTI a{}, b{};
    a.x = x; a.y = y; a.z = z; a.intensity = (float)i;
    b.x = x - 0.10f; b.y = y + 0.05f; b.z = z - 0.02f; b.intensity = (float)i;   // target = source + (0.10, -0.05, 0.02)
    tgt->push_back(a);
    src->push_back(b);
  }
  pcl::Registration<PointTI, PointTI>::Ptr icp = select_registration_method();
  icp->setInputSource(src);
  icp->setInputTarget(tgt);
  Eigen::Matrix4f guess = Eigen::Matrix4f::Identity();
  guess(0, 3) = 0.05f;
  icp->align(*out, guess);
  Eigen::Matrix4f T = icp->getFinalTransformation();
  auto* fg = dynamic_cast<fast_gicp::FastGICP<PointTI, PointTI>*>(icp.get());
  const double fit = fg->getFitnessScore(1.0);
  Eigen::Matrix<double, 6, 6> H = fg->getFinalHessian(), H2;
  Eigen::Matrix<double, 6, 1> b6;
  const double cost = fg->evaluateCost(T, &H2, &b6);
  const auto& covs = fg->getSourceCovariances();
  std::printf("converged %d t = %.4f %.4f %.4f fitness %.6f cost %.4f H00 %.3f covs %zu out %zu\n", (int)icp->hasConverged(), T(0, 3), T(1, 3), T(2, 3), fit,
              cost, H(0, 0), covs.size(), out->size());
  bool ok = icp->hasConverged() && std::fabs(T(0, 3) - 0.10f) < 0.02f && std::fabs(T(1, 3) + 0.05f) < 0.02f && std::fabs(T(2, 3) - 0.02f) < 0.02f;
  ok = ok && out->size() == src->size() && (*out)[7].intensity == 7.0f && std::fabs((*out)[7].x - ((*src)[7].x + T(0, 3))) < 0.02f;
  ok = ok && fit >= 0.0 && fit < 0.05 && covs.size() == src->size() && H(0, 0) > 0.0 && std::fabs(H2(0, 0)) > 0.0 && cost >= 0.0;
  ok = ok && std::fabs(covs[0](0, 0) - covs[0](0, 0)) == 0.0 && covs[0](3, 3) == 0.0;
  // swap / same-pointer early-out / covariance hand-over (fast_gicp_impl.hpp:50-100)
  fg->setInputSource(src);                      // same pointer: no-op
  fast_gicp::CovarianceList keep = covs;
  fg->swapSourceAndTarget();
  ok = ok && fg->getTargetCovariances().size() == keep.size();
  pcl::PointCloud<PointTI> out2;
  fg->align(out2);
  Eigen::Matrix4f Tb = fg->getFinalTransformation();
  ok = ok && fg->hasConverged() && std::fabs(Tb(0, 3) + 0.10f) < 0.02f;
  fg->clearSource();
  fg->clearTarget();
  // the other two instantiations exist in the library
  fast_gicp::FastGICP<pcl::PointXYZ, pcl::PointXYZ> g1;
  fast_gicp::FastGICP<pcl::PointNormal, pcl::PointNormal> g2;
  g1.setRegularizationMethod(fast_gicp::RegularizationMethod::FROBENIUS);
  g2.setLSQType(fast_gicp::LSQ_OPTIMIZER_TYPE::GaussNewton);
  std::printf(ok ? "DROPIN_OK\n" : "DROPIN_FAIL\n");
  return ok ? 0 : 1;
}
